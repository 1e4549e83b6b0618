// Evaluation of a Linear-ReLU-Linear network on the tensor cores: one CTA per 128-sample
// tile, a 3-stage bulk-copy pipeline feeding tcgen05.mma (fp32 accumulate in TMEM), and an
// epilogue that finishes the network (bias + ReLU + second layer + argmax) and reduces the tile to a
// confusion matrix -- only C*C integers leave the SM.  Reference: gossipy/model/handler.py:282-334
// (full-test-set forward on the device, predictions shipped to scikit-learn on the host).
//
//   z1[128 x NP] = Xtile[128 x IN] . W1[NP x IN]^T        NP = hidden units padded to 16, K in stages of 32
//
// Default (X3): fp32-equivalent products, a.b ~= a_hi.b_hi + a_hi.b_lo + a_lo.b_hi (kind::tf32 truncates its
// operands, so the raw tiles ARE the hi operands).  The lo images (x - trunc(x), exact) are produced in shared
// memory by the six otherwise idle warps while the tensor core works on the hi product of the same stage: no
// second copy of the test set, no extra HBM traffic.  Plain tf32 (GlobalSettings().allow_tf32) skips them.
//
// Both operands are read from PRE-TILED images (K-major core matrices, one contiguous block per
// pipeline stage, so a stage is two cp.async.bulk copies): the test set is tiled once and cached
// (it never changes), W1 is tiled by a ~3 us pre-pass per evaluation (it changes every round).
// Warp 0 = producer (elected lane), warp 1 = MMA issuer (elected lane), all 8 warps = epilogue.
#include "tc_common.cuh"
#include "kernels.h"
#include <algorithm>

namespace gb {

constexpr int EV_TM = 128;             // samples per tile (MMA M)
constexpr int EV_KS = 32;              // features per pipeline stage
constexpr int EV_STAGES = 3;
constexpr int EV_THREADS_TC = 256;
constexpr int EV_NP_MAX = 128;
constexpr int EV_A_BYTES = EV_TM * EV_KS * 4;                 // 16384
constexpr int EV_CH = EV_KS / 4;                              // 16-byte chunks per row and stage
constexpr int EV_SPLIT_WARPS = 6;                             // warps 2..7 write the lo images

// rows x IN (row-major) -> [tile][stage][row group][EV_CH chunks][8 rows][4 floats], zero padded
__global__ void __launch_bounds__(256)
eval_pretile_kernel(const float* __restrict__ src, int rows, int IN, int rows_per_tile, int nstage,
                    float* __restrict__ dst, int64_t total_chunks) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total_chunks;
         q += (int64_t)gridDim.x * blockDim.x) {
        const int r8 = (int)(q & 7);
        const int c = (int)((q >> 3) % EV_CH);
        int64_t rest = (q >> 3) / EV_CH;
        const int groups = rows_per_tile >> 3;
        const int g = (int)(rest % groups); rest /= groups;
        const int s = (int)(rest % nstage);
        const int64_t t = rest / nstage;
        const int64_t row = t * rows_per_tile + g * 8 + r8;
        const int col = s * EV_KS + c * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < rows) {
            const float* p = src + row * IN + col;
            if (col + 3 < IN) v = *reinterpret_cast<const float4*>(p);
            else {
                if (col < IN) v.x = p[0];
                if (col + 1 < IN) v.y = p[1];
                if (col + 2 < IN) v.z = p[2];
            }
        }
        reinterpret_cast<float4*>(dst)[q] = v;
    }
}

struct EvSmem {
    static constexpr int a = 0;                                        // [3][16384]
    static constexpr int b = a + EV_STAGES * EV_A_BYTES;               // [3][NP_MAX*32*4 = 16384]
    static constexpr int alo = b + EV_STAGES * EV_NP_MAX * EV_KS * 4;  // lo images of the same stages
    static constexpr int blo = alo + EV_STAGES * EV_A_BYTES;
    static constexpr int w2 = blo + EV_STAGES * EV_NP_MAX * EV_KS * 4; // [128][12]
    static constexpr int b1 = w2 + EV_NP_MAX * 12 * 4;
    static constexpr int part = b1 + EV_NP_MAX * 4;                    // [128][12] partial logits of the second half
    static constexpr int cm = part + EV_TM * 12 * 4;                   // [16*16] int
    static constexpr int mbar = cm + 256 * 4;                          // full[3], empty[3], split[3], done
    static constexpr int tslot = mbar + 96;
    static constexpr int total = tslot + 16;
};

GB_DEVICE void ev_bulk(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(mbar)) : "memory");
}

template <bool X3>
__global__ void __launch_bounds__(EV_THREADS_TC, 1)
mlp1_eval_tc_kernel(const float* __restrict__ row, const float* __restrict__ xt /* pre-tiled test set */,
                    const float* __restrict__ w1t /* pre-tiled W1 */, const int64_t* __restrict__ y, int n,
                    int IN, int H, int OUT, int NP, int nstage, int n_classes, int* __restrict__ cm_out,
                    float* __restrict__ score1) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int quad = warp & 3, half = warp >> 2;
    float* w2s = reinterpret_cast<float*>(smem + EvSmem::w2);
    float* b1s = reinterpret_cast<float*>(smem + EvSmem::b1);
    float* part = reinterpret_cast<float*>(smem + EvSmem::part);
    int* cms = reinterpret_cast<int*>(smem + EvSmem::cm);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + EvSmem::mbar);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + EvSmem::tslot);
    const float* b1g = row + (size_t)H * IN;
    const float* W2g = b1g + H;
    const float* b2g = W2g + (size_t)OUT * H;
    const uint32_t b_bytes = (uint32_t)NP * EV_KS * 4u;

    if (warp == 2) tmem_alloc<128>(tslot);
    if (tid == 0) {
        for (int i = 0; i < 2 * EV_STAGES; ++i) mbar_init(&mbar[i], 1);
        for (int i = 0; i < EV_STAGES; ++i) mbar_init(&mbar[2 * EV_STAGES + i], EV_SPLIT_WARPS);
        mbar_init(&mbar[3 * EV_STAGES], 1);
        mbar_fence_init();
    }
    for (int i = tid; i < EV_NP_MAX * 12; i += EV_THREADS_TC) {
        const int j = i / 12, o = i % 12;
        w2s[i] = (j < H && o < OUT) ? W2g[(size_t)o * H + j] : 0.f;
    }
    if (tid < EV_NP_MAX) b1s[tid] = tid < H ? b1g[tid] : 0.f;
    if (tid < 256) cms[tid] = 0;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tslot, 0);
    const int tile = blockIdx.x;
    const float* a_src = xt + (size_t)tile * nstage * (EV_TM * EV_KS);
    uint64_t* full = mbar, *empty = mbar + EV_STAGES, *split = mbar + 2 * EV_STAGES, *done = mbar + 3 * EV_STAGES;

    if (warp == 0) {                    // ---- producer: two bulk copies per stage ----
        if (elect_one()) {
            for (int s = 0; s < nstage; ++s) {
                const int st = s % EV_STAGES;
                if (s >= EV_STAGES) mbar_wait(&empty[st], (uint32_t)(((s / EV_STAGES) - 1) & 1));
                mbar_expect_tx(&full[st], EV_A_BYTES + b_bytes);
                ev_bulk(smem + EvSmem::a + st * EV_A_BYTES, a_src + (size_t)s * (EV_TM * EV_KS), EV_A_BYTES, &full[st]);
                ev_bulk(smem + EvSmem::b + st * (EV_NP_MAX * EV_KS * 4), w1t + (size_t)s * NP * EV_KS, b_bytes, &full[st]);
            }
        }
        __syncwarp();
    } else if (warp == 1) {             // ---- MMA issuer ----
        if (elect_one()) {
            const uint32_t idesc = make_idesc(kFmtTF32, kFmtTF32, EV_TM, NP, false, false);
            for (int s = 0; s < nstage; ++s) {
                const int st = s % EV_STAGES;
                mbar_wait(&full[st], (uint32_t)((s / EV_STAGES) & 1));
                tc_fence_after();
                const uint64_t ad = make_sdesc(smem_u32(smem + EvSmem::a + st * EV_A_BYTES), 128u, EV_CH * 128u);
                const uint64_t bd = make_sdesc(smem_u32(smem + EvSmem::b + st * (EV_NP_MAX * EV_KS * 4)), 128u, EV_CH * 128u);
#pragma unroll
                for (int k = 0; k < EV_KS / 8; ++k)
                    mma_tf32_ss(tmem, ad + (uint64_t)(k * 16), bd + (uint64_t)(k * 16), idesc, (s | k) != 0);
                if (X3) {
                    mbar_wait(&split[st], (uint32_t)((s / EV_STAGES) & 1));      // lo images of this stage written
                    tc_fence_after();
                    const uint64_t al = make_sdesc(smem_u32(smem + EvSmem::alo + st * EV_A_BYTES), 128u, EV_CH * 128u);
                    const uint64_t bl = make_sdesc(smem_u32(smem + EvSmem::blo + st * (EV_NP_MAX * EV_KS * 4)), 128u, EV_CH * 128u);
#pragma unroll
                    for (int k = 0; k < EV_KS / 8; ++k)
                        mma_tf32_ss(tmem, ad + (uint64_t)(k * 16), bl + (uint64_t)(k * 16), idesc, true);
#pragma unroll
                    for (int k = 0; k < EV_KS / 8; ++k)
                        mma_tf32_ss(tmem, al + (uint64_t)(k * 16), bd + (uint64_t)(k * 16), idesc, true);
                }
                mma_commit(&empty[st]);                 // stage is free again when these MMAs retire
            }
            mma_commit(done);
        }
        __syncwarp();
    }
    else if (X3) {                      // ---- warps 2..7: lo = x - trunc(x) of both operand tiles of every stage ----
        const int t2 = tid - 64, n2 = EV_SPLIT_WARPS * 32;
        const int a4 = EV_A_BYTES / 16, b4 = (int)(b_bytes / 16u);
        for (int s = 0; s < nstage; ++s) {
            const int st = s % EV_STAGES;
            mbar_wait(&full[st], (uint32_t)((s / EV_STAGES) & 1));
            const float4* as = reinterpret_cast<const float4*>(smem + EvSmem::a + st * EV_A_BYTES);
            const float4* bs = reinterpret_cast<const float4*>(smem + EvSmem::b + st * (EV_NP_MAX * EV_KS * 4));
            float4* ad = reinterpret_cast<float4*>(smem + EvSmem::alo + st * EV_A_BYTES);
            float4* bd = reinterpret_cast<float4*>(smem + EvSmem::blo + st * (EV_NP_MAX * EV_KS * 4));
            for (int i = t2; i < a4 + b4; i += n2) {
                const float4 v = i < a4 ? as[i] : bs[i - a4];
                const float4 lo = make_float4(v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u),
                                              v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u),
                                              v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u),
                                              v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u));
                if (i < a4) ad[i] = lo; else bd[i - a4] = lo;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(&split[st])) : "memory");
        }
    }
    // ---- epilogue: thread (sample = 32*quad + lane, half) finishes its half of the hidden units ----
    mbar_wait(done, 0u);
    tc_fence_after();
    const int sidx = quad * 32 + lane;
    const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);
    float logit[10];
#pragma unroll
    for (int o = 0; o < 10; ++o) logit[o] = 0.f;
    const int ngroups = NP >> 4;
    const int g0 = half ? (ngroups + 1) / 2 : 0, g1 = half ? ngroups : (ngroups + 1) / 2;
    for (int g = g0; g < g1; ++g) {
        float z[16];
        tmem_ld16(tlane + g * 16, z);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int j = g * 16 + i;
            const float hv = fmaxf(z[i] + b1s[j], 0.f);
            const float4 wa = *reinterpret_cast<const float4*>(w2s + j * 12);
            const float4 wb = *reinterpret_cast<const float4*>(w2s + j * 12 + 4);
            const float4 wc = *reinterpret_cast<const float4*>(w2s + j * 12 + 8);
            logit[0] = fmaf(hv, wa.x, logit[0]); logit[1] = fmaf(hv, wa.y, logit[1]);
            logit[2] = fmaf(hv, wa.z, logit[2]); logit[3] = fmaf(hv, wa.w, logit[3]);
            logit[4] = fmaf(hv, wb.x, logit[4]); logit[5] = fmaf(hv, wb.y, logit[5]);
            logit[6] = fmaf(hv, wb.z, logit[6]); logit[7] = fmaf(hv, wb.w, logit[7]);
            logit[8] = fmaf(hv, wc.x, logit[8]); logit[9] = fmaf(hv, wc.y, logit[9]);
        }
    }
    if (half == 1) {
#pragma unroll
        for (int o = 0; o < 10; ++o) part[sidx * 12 + o] = logit[o];
    }
    __syncthreads();
    if (half == 0) {
        const int gs = tile * EV_TM + sidx;
        if (gs < n) {
            int best = 0; float bv = -3.0e38f;
#pragma unroll
            for (int o = 0; o < 10; ++o) {
                const float v = logit[o] + part[sidx * 12 + o] + (o < OUT ? b2g[o] : 0.f);
                if (o < OUT && v > bv) { bv = v; best = o; }
                if (score1 != nullptr && o == (OUT > 1 ? 1 : 0)) score1[gs] = v;       // class-1 logit (AUC of 2-output nets)
            }
            const int t = (int)y[gs];
            if (t >= 0 && t < n_classes && best < n_classes) atomicAdd(&cms[t * 16 + best], 1);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 256) {
        const int t = tid >> 4, pcl = tid & 15;
        const int v = cms[tid];
        if (v != 0 && t < n_classes && pcl < n_classes) atomicAdd(&cm_out[t * n_classes + pcl], v);
    }
    if (warp == 2) tmem_dealloc<128>(tmem);
}

// ---- host side -------------------------------------------------------------------------------------
int64_t mlp1_eval_pretile_floats(int n, int IN) {
    const int64_t ntile = (n + EV_TM - 1) / EV_TM;
    const int64_t nstage = (IN + EV_KS - 1) / EV_KS;
    return ntile * nstage * EV_TM * EV_KS;
}

void launch_mlp1_eval_pretile(const float* X, int n, int IN, float* out, cudaStream_t stream) {
    const int nstage = (IN + EV_KS - 1) / EV_KS;
    const int64_t chunks = mlp1_eval_pretile_floats(n, IN) / 4;
    const int blocks = (int)std::min<int64_t>((chunks + 255) / 256, (int64_t)sm_count() * 16);
    eval_pretile_kernel<<<blocks, 256, 0, stream>>>(X, n, IN, EV_TM, nstage, out, chunks);
}

// W1 pre-tile scratch: one buffer per (device, stream) -- evaluations of different nodes run on
// different streams concurrently, successive evaluations of a stream are ordered by the stream
struct W1Slot { float* ptr; cudaStream_t stream; int dev; };
static W1Slot g_w1slots[256] = {};

static bool g_eval_tf32 = false;        // GlobalSettings().allow_tf32: plain tf32 products in the evaluation kernel
void set_eval_tf32(bool on) { g_eval_tf32 = on; }

bool launch_mlp1_eval_tc(const float* row, const float* xt, const int64_t* y, int n, int IN, int H, int OUT,
                         int n_classes, int* cm, float* score1, cudaStream_t stream) {
    if (H > EV_NP_MAX || OUT > 10 || n_classes > 16 || n <= 0 || IN % 4 != 0) return false;
    const int NP = (H + 15) & ~15;
    const int nstage = (IN + EV_KS - 1) / EV_KS;
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t w1_floats = (size_t)nstage * EV_NP_MAX * EV_KS;
    static size_t slot_floats = 0;
    float* slot = nullptr;
    int free_i = -1;
    for (int i = 0; i < 256; ++i) {
        if (g_w1slots[i].ptr != nullptr && g_w1slots[i].stream == stream && g_w1slots[i].dev == dev) { slot = g_w1slots[i].ptr; break; }
        if (g_w1slots[i].ptr == nullptr && free_i < 0) free_i = i;
    }
    if (slot != nullptr && w1_floats > slot_floats) return false;       // shape grew: let the simt kernel handle it
    if (slot == nullptr) {
        if (free_i < 0) return false;
        const size_t want = std::max(w1_floats, (size_t)(1024 / EV_KS) * EV_NP_MAX * EV_KS);   // room for IN <= 1024
        if (cudaMalloc(&slot, want * 4) != cudaSuccess) { cudaGetLastError(); return false; }
        slot_floats = std::max(slot_floats, want);
        if (w1_floats > want) return false;
        g_w1slots[free_i] = W1Slot{slot, stream, dev};
    }
    const int64_t chunks = (int64_t)nstage * NP * EV_KS / 4;
    eval_pretile_kernel<<<(int)((chunks + 255) / 256), 256, 0, stream>>>(row, H, IN, NP, nstage, slot, chunks);
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(mlp1_eval_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, EvSmem::total) != cudaSuccess ||
            cudaFuncSetAttribute(mlp1_eval_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, EvSmem::total) != cudaSuccess)
            return false;
        configured = true;
    }
    const int ntile = (n + EV_TM - 1) / EV_TM;
    if (g_eval_tf32)
        mlp1_eval_tc_kernel<false><<<ntile, EV_THREADS_TC, EvSmem::total, stream>>>(row, xt, slot, y, n, IN, H, OUT, NP, nstage, n_classes, cm, score1);
    else
        mlp1_eval_tc_kernel<true><<<ntile, EV_THREADS_TC, EvSmem::total, stream>>>(row, xt, slot, y, n, IN, H, OUT, NP, nstage, n_classes, cm, score1);
    return cudaGetLastError() == cudaSuccess;
}

void preload_eval_tc() {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, eval_pretile_kernel);
    cudaFuncGetAttributes(&a, mlp1_eval_tc_kernel<true>);
    cudaFuncGetAttributes(&a, mlp1_eval_tc_kernel<false>);
}

}  // namespace gb
