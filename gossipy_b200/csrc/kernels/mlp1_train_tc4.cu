// tcgen05 / TMEM training kernel, fourth generation ("tc4"): fp32-EQUIVALENT local update of a 1-hidden-layer
// ReLU MLP on a cluster of NC = 8 (or 4, plain tf32 only) CTAs.  Reference semantics: gossipy/model/handler.py:235-258.
//
// Why a new kernel: kind::tf32 MMAs TRUNCATE their fp32 operands to 10 mantissa bits and the tensor core adds into
// its accumulator with round-toward-zero (both measured: benchmarks/probe_tc3.py, benchmarks/check_tc4.py), which is
// below the reference's precision: accumulating every step's update straight into the fp32 master weights shrinks
// |W| by ~4e-7 per step.  Here (X3 = true)
//  * every tensor-core product is error compensated (3xTF32):  a.b ~= a_hi.b_hi + a_hi.b_lo + a_lo.b_hi  with
//    x_hi = x & 0xffffe000, x_lo = x - x_hi (exact) and fp32 accumulation: relative error ~2^-21 instead of 2^-11;
//  * the update is accumulated into a ZEROED TMEM tile G and added to the master in registers (round to nearest)
//    by the pass that re-splits W into hi / lo anyway; the three forward chains use separate accumulators.
// Per SGD step and CTA (FP = input features of this CTA's slice, <= 128):
//   fwd   D1a = W . Xhi^T,  D1b = W . Xlo^T,  D1c = Wlo . Xhi^T          (M128 N32 K=FP, A = fp32 master in TMEM)
//   upd   G   = A2hi . Xhi + A2hi . Xlo + A2lo . Xhi                      (M128 N=FP K32)
//   pass  W += G ; Wlo = W - trunc(W)      TMEM -> registers -> TMEM, overlapped with D1a / D1b of the next step
// The input features are split over the NC CTAs (K split of the forward GEMM, N split of the update).  The partial
// z1 sums are combined with a reduce-scatter / all-gather over distributed shared memory (st.async, completion on
// the receiver's mbarrier):
//   RS: CTA c receives every CTA's partial sums of ITS 32/NC samples, adds them in a fixed order, applies bias + ReLU
//       and -- owning complete hidden activations of those samples -- computes their logits, softmax and dz2;
//       then dh = dz2 . W2, the ReLU mask and dz1 of those samples;
//   AG: dz1 and dz2 of the owned samples go to all CTAs, which write the operand images of the update MMA;
//   the hidden activations of the owned samples additionally go, sliced by hidden unit, to the CTA that owns that
//   slice of W2: CTA c forms gW2[:, 16c .. 16c+15] over all 32 samples, applies the SGD step and all-gathers its
//   slice of the new W2 (4-byte st.async, off the critical path) -- every CTA holds identical b1 / W2 / b2 replicas.
// The second layer (32 x 100 x 10) therefore runs DISTRIBUTED on the CUDA cores in exact fp32, ~15 kFLOP per CTA and
// step: as tf32 MMAs it is 19 M64/M128 K8 instructions of ~45 cycles each, 57 with error compensation.
// K3 (SC = true): PartitionedTMH divides the gradient of every parameter by the age of its partition
// (reference model/handler.py:497-520).  Because the update reaches the master through registers anyway (W += G),
// the per-element factor 1/age[part(j, k)] is applied there -- partition ids of a thread's 56 weights are packed
// into 7 registers at kernel start -- and to the b1 / W2 / b2 steps; the tensor-core work is unchanged.
// Momentum (MOM = true): torch.optim.SGD with momentum / dampening / nesterov / weight decay.  The momentum buffer of
// W1 lives in a fourth TMEM tile V next to W, Wlo and G (4 x 112 columns + 64 accumulator columns = 512), and the
// W += G pass becomes  d = G + wd*W;  V = mu*V + (1-tau)*d;  W -= lr*(nesterov ? d + mu*V : V)  in registers; the
// buffers of b1 / W2 / b2 sit with their owners.  State is loaded from / stored to the handler's momentum row.
// Warp roles: warps 0-7 compute, warp 8 issues all MMAs and bulk copies and never touches data.
// Operand tiles (hi and lo images of X in both K-major layouts) are written ahead of time by mlp1_stage4_kernel.
#include "tc_common.cuh"
#include "kernels.h"

namespace gb {

constexpr int T4_CTHREADS = 256;          // compute threads (warps 0-7)
constexpr int T4_THREADS = 288;           // + warp 8
constexpr int T4_ISSUER = 8;
constexpr int T4_B = 32;
constexpr int T4_HP = 128;
constexpr int T4_OUTV = 10;               // classes handled (padding classes: weights 0, bias -3e38)
constexpr int T4_DZP = 12;                // floats per dz2 row in shared memory (3 x float4)
constexpr int T4_WCB = 64, T4_WLD = 65;   // TMEM fill / write-back scratch: 64-column blocks, odd pitch
constexpr int T4_NPROF = 14;   // phase counters (compute thread 0, warp 4 lane 0 and the issuer lane)

template <int NC, bool X3, bool MOM = false> struct T4Cfg {
    static_assert(!X3 || NC == 8, "the error-compensated path needs W, Wlo and G in TMEM: 3 x FP <= 384 columns");
    static_assert(!MOM || X3, "momentum rides on the W += G pass of the error-compensated form");
    static constexpr int NIMG = X3 ? 2 : 1;
    static constexpr int FP_MAX = NC == 4 ? 240 : (MOM ? 112 : 128);   // feature columns per CTA (multiple of 16)
    static constexpr int S = T4_B / NC;                         // samples owned per CTA
    static constexpr int GPO = S / 4;                           // float4 sample groups per owner
    static constexpr int t_w1 = 0, t_wlo = FP_MAX, t_g = 2 * FP_MAX, t_v = 3 * FP_MAX;
    static constexpr int t_d1 = MOM ? 4 * FP_MAX : X3 ? 3 * FP_MAX : FP_MAX;
    static constexpr int d1_cols = MOM ? 64 : X3 ? 96 : 32;     // X3: three accumulators; MOM: chain c shares chain a's
    // shared memory (byte offsets; base rounded up to 1024 B)
    static constexpr int tile_max = T4_B * FP_MAX * 4;
    static constexpr int xf = 0;
    static constexpr int xt = xf + NIMG * tile_max;
    static constexpr int a2 = xt + NIMG * tile_max;              // update A operand [128 x 32] K-major, hi (+ lo)
    static constexpr int rs = a2 + NIMG * T4_HP * T4_B * 4;      // [NC src][GPO][128][4]
    static constexpr int ag = rs + 8 * T4_HP * 16;               // [8 sample groups][128][4]
    static constexpr int JS = T4_HP / NC;                        // hidden units per W2 slice
    static constexpr int hsl = ag + 8 * T4_HP * 16;              // h of my W2 slice: [8 sample groups][JS][4]
    static constexpr int w2s = hsl + 8 * JS * 16;                // W2 [10][128]
    static constexpr int gb1p = w2s + T4_OUTV * T4_HP * 4;       // [2 halves][128]
    static constexpr int dzb = gb1p + 2 * T4_HP * 4;             // dz2 [32][12] (all samples, via the all-gather)
    static constexpr int dzo = dzb + T4_B * T4_DZP * 4;          // dz2 of my own samples [S][12]
    static constexpr int red = dzo + 8 * T4_DZP * 4;             // [8 warps][40]
    static constexpr int b2 = red + 8 * 40 * 4;
    static constexpr int inva = b2 + 16 * 4;                     // [2][16] 1/age of every partition, this / next step
    static constexpr int ys = inva + 2 * 16 * 4;                 // [2][32] int
    static constexpr int mbar = ys + 2 * T4_B * 4;               // 11 x uint64 (16 reserved)
    static constexpr int tslot = mbar + 128;
    static constexpr int total = tslot + 16;
    static_assert(t_d1 + d1_cols <= 512, "TMEM budget");
    static_assert(total + 1024 <= 227 * 1024, "shared memory budget");
    static_assert(T4_HP * T4_WLD * 4 <= hsl - a2, "fill / write-back scratch must fit in a2 + rs + ag");
    static_assert(NC * GPO == 8, "8 float4 sample groups");
};

GB_DEVICE void bulk_g2s4(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(mbar)) : "memory");
}
GB_DEVICE void mbar_arrive(uint64_t* mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(mbar)) : "memory");
}
// wait for st.async traffic from other CTAs of the cluster (acquire at cluster scope).  Measured alternatives that
// did not change the step time: CTA-scope acquire (no CCTL.IVALL), a suspend-time hint, one polling lane per warp
// (4x slower) -- the shared-memory port is busy with ~32 KB of DSMEM traffic per step at ~20 B/cycle instead.
GB_DEVICE void mbar_wait_dsm(uint64_t* mbar, uint32_t parity) { mbar_wait_cluster(mbar, parity); }
GB_DEVICE void bar_compute() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
GB_DEVICE void st_async_f32(uint32_t remote_addr, float v, uint32_t remote_mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];"
                 :: "r"(remote_addr), "f"(v), "r"(remote_mbar) : "memory");
}
GB_DEVICE float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// sum 40 per-lane values over the warp with 45 shuffles: afterwards every lane whose bits 2..4 are g holds the
// totals of the original values 5g .. 5g+4 in v[0..4]
GB_DEVICE void warp_reduce40(float (&v)[40], int lane) {
    {
        const bool up = (lane & 16) != 0;
#pragma unroll
        for (int i = 0; i < 20; ++i) {
            const float keep = up ? v[i + 20] : v[i], send = up ? v[i] : v[i + 20];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    {
        const bool up = (lane & 8) != 0;
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            const float keep = up ? v[i + 10] : v[i], send = up ? v[i] : v[i + 10];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    {
        const bool up = (lane & 4) != 0;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const float keep = up ? v[i + 5] : v[i], send = up ? v[i] : v[i + 5];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        v[i] += __shfl_xor_sync(0xffffffffu, v[i], 2);
        v[i] += __shfl_xor_sync(0xffffffffu, v[i], 1);
    }
}

template <int NC, bool X3, bool SC, bool MOM>
__global__ void __launch_bounds__(T4_THREADS, 1)
mlp1_train_tc4_kernel(const TrainParams p, const int FPC, const int FP, const int total_steps,
                      const float* __restrict__ stage, const int* __restrict__ stage_ys) {
    using C = T4Cfg<NC, X3, MOM>;
    static_assert(!(MOM && SC), "momentum and partition scaling are not combined");
    constexpr int NIMG = C::NIMG, GPO = C::GPO, S = C::S;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // round up on the SHARED-window address: going through uintptr_t would lose the address space
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int quad = warp & 3, half = (warp >> 2) & 1;
    const int j = quad * 32 + lane;                        // hidden unit = TMEM lane (compute warps)
    const uint32_t rank = gb_cluster_ctarank();
    const int IN = p.IN, H = p.H, OUT = p.OUT, n = p.n, B = p.B;
    const int f0 = (int)rank * FPC;
    const int fcnt = max(0, min(FPC, IN - f0));
    const int nchunk = FP >> 2;

    float* xf = reinterpret_cast<float*>(smem + C::xf);
    float* xt = reinterpret_cast<float*>(smem + C::xt);
    float* a2 = reinterpret_cast<float*>(smem + C::a2);
    float* rsb = reinterpret_cast<float*>(smem + C::rs);
    float* agb = reinterpret_cast<float*>(smem + C::ag);
    float* hsl = reinterpret_cast<float*>(smem + C::hsl);
    float* w2s = reinterpret_cast<float*>(smem + C::w2s);
    float* gb1p = reinterpret_cast<float*>(smem + C::gb1p);
    float* dzb = reinterpret_cast<float*>(smem + C::dzb);
    float* dzo = reinterpret_cast<float*>(smem + C::dzo);
    float* red = reinterpret_cast<float*>(smem + C::red);
    float* b2s = reinterpret_cast<float*>(smem + C::b2);
    float* inva = reinterpret_cast<float*>(smem + C::inva);
    int* ysm = reinterpret_cast<int*>(smem + C::ys);
    // 0 xf landed, 1 xt landed, 2 forward done, 3 update done, 4 W2 all-gather, 5 RS, 6 AG, 7 a2 ready (256), 8 W/Wlo ready (256),
    // 9 h slices of my W2 columns
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + C::mbar);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + C::tslot);

    float* b1g = p.row + (size_t)H * IN;
    float* W2g = b1g + H;
    float* b2g = W2g + (size_t)OUT * H;
    const uint32_t tile_bytes = (uint32_t)T4_B * (uint32_t)FP * 4u;
    const size_t tile_floats = (size_t)T4_B * FP;
    const size_t blk = 2 * NIMG * tile_floats;             // staged floats per (step, rank): xf images, then xt images
    const float* my_stage = stage + (size_t)rank * blk;    // + s * NC * blk

    // ---- fused MERGE_UPDATE pre-pass: row = w_self*row + w_peer*peer (peer possibly in another GPU's HBM) -------
    if (p.peer != nullptr) {
        if (p.sync.ready != nullptr) {
            if (tid == 0) gb_wait_flag(p.sync.ready, p.sync.gen, p.sync.fault);
            __syncthreads();
        }
        const int64_t P = (int64_t)H * IN + H + (int64_t)OUT * H + OUT;
        const int64_t n4 = ((P + 31) & ~(int64_t)31) >> 2;              // rows are padded to 32 floats
        float4* own4 = reinterpret_cast<float4*>(p.row);
        const float4* peer4 = reinterpret_cast<const float4*>(p.peer);
        constexpr int U = 4;
        const int64_t stride = (int64_t)NC * T4_THREADS;
        int64_t i = (int64_t)rank * T4_THREADS + tid;
        for (; i + (U - 1) * stride < n4; i += U * stride) {
            float4 q[U], o[U];
#pragma unroll
            for (int u = 0; u < U; ++u) q[u] = gb_ld_stream(peer4 + i + u * stride);
#pragma unroll
            for (int u = 0; u < U; ++u) o[u] = own4[i + u * stride];
#pragma unroll
            for (int u = 0; u < U; ++u)
                own4[i + u * stride] = make_float4(p.w_self * o[u].x + p.w_peer * q[u].x, p.w_self * o[u].y + p.w_peer * q[u].y,
                                                   p.w_self * o[u].z + p.w_peer * q[u].z, p.w_self * o[u].w + p.w_peer * q[u].w);
        }
        for (; i < n4; i += stride) {
            const float4 q = gb_ld_stream(peer4 + i), o = own4[i];
            own4[i] = make_float4(p.w_self * o.x + p.w_peer * q.x, p.w_self * o.y + p.w_peer * q.y,
                                  p.w_self * o.z + p.w_peer * q.z, p.w_self * o.w + p.w_peer * q.w);
        }
        __threadfence();
        gb_cluster_sync();                                       // the merged row is visible to all CTAs
        if (p.sync.done != nullptr && rank == 0 && tid == 0) gb_red_release_sys_add(p.sync.done, 1u);
    }

    // ---- one-time set-up ---------------------------------------------------------------------------------------
    const size_t off_b1 = (size_t)H * IN, off_w2 = off_b1 + H, off_b2 = off_w2 + (size_t)OUT * H;
    if (warp == 0) tmem_alloc<512>(tslot);
    if (tid == 0) {
        for (int i = 0; i < 7; ++i) mbar_init(&mbar[i], 1);
        mbar_init(&mbar[7], T4Cfg<NC, X3>::GPO == 1 ? 4 : 8);      // one arrival per writing warp
        mbar_init(&mbar[8], 8);
        mbar_init(&mbar[9], 1);
        mbar_init(&mbar[10], 8);                                     // first 64 columns of W / Wlo rewritten (pipelined forward)
        mbar_fence_init();
    }
    float b1r = (warp < T4_ISSUER && j < H) ? p.row[off_b1 + j] : 0.f;       // both threads of hidden unit j carry b1[j]
    for (int i = tid; i < T4_OUTV * T4_HP; i += T4_THREADS) {
        const int o = i / T4_HP, jj = i % T4_HP;
        w2s[i] = (o < OUT && jj < H) ? p.row[off_w2 + (size_t)o * H + jj] : 0.f;
    }
    if (tid < 16) b2s[tid] = (tid < OUT) ? p.row[off_b2 + tid] : -3.0e38f;   // padding classes: probability 0
    if (SC && tid < 16) {        // ages are incremented before the step (reference handler.py:506)
        inva[tid] = tid < p.n_parts ? 1.f / (float)(p.age_of(tid) + 1) : 1.f;
        inva[16 + tid] = 1.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tslot;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);

    // first operand tiles and labels: issue now, they land while the weights are loaded
    if (tid == T4_ISSUER * 32) {
        mbar_expect_tx(&mbar[0], NIMG * tile_bytes);
        bulk_g2s4(xf, my_stage, NIMG * tile_bytes, &mbar[0]);
        mbar_expect_tx(&mbar[1], NIMG * tile_bytes);
        bulk_g2s4(xt, my_stage + NIMG * tile_floats, NIMG * tile_bytes, &mbar[1]);
    }
    if (tid < T4_B) ysm[tid] = stage_ys[tid];

    // master weights -> TMEM (W and, for 3xTF32, Wlo).  64-column blocks move through a [128][65] scratch
    // (a2 + rs + ag, unused until the cluster barrier below): coalesced 128-bit global loads -> conflict-free
    // column reads -> tcgen05.st.
    {
        float* wbuf = a2;
        for (int cb = 0; cb * T4_WCB < FP; ++cb) {
            const int c0 = cb * T4_WCB;
            for (int idx = tid; idx < H * (T4_WCB / 4); idx += T4_THREADS) {
                const int rr = idx >> 4, c = c0 + ((idx & 15) << 2);
                float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c < fcnt) w = *reinterpret_cast<const float4*>(p.row + (size_t)rr * IN + f0 + c);
                float* d = wbuf + rr * T4_WLD + (c - c0);
                d[0] = w.x; d[1] = w.y; d[2] = w.z; d[3] = w.w;
            }
            __syncthreads();
            if (warp < T4_ISSUER) {
#pragma unroll
                for (int gl = 0; gl < 2; ++gl) {
                    const int g = cb * 4 + half * 2 + gl;                     // warp-uniform
                    if (g * 16 < FP) {
                        float v[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = (j < H) ? wbuf[j * T4_WLD + (half * 2 + gl) * 16 + i] : 0.f;
                        tmem_st16(tlane + C::t_w1 + g * 16, v);
                        if (X3) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = v[i] - tf32_hi(v[i]);
                            tmem_st16(tlane + C::t_wlo + g * 16, v);
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (MOM) {                // momentum buffer of W1 -> TMEM tile V (zeros when there is no state yet)
            for (int cb = 0; cb * T4_WCB < FP; ++cb) {
                const int c0 = cb * T4_WCB;
                for (int idx = tid; idx < H * (T4_WCB / 4); idx += T4_THREADS) {
                    const int rr = idx >> 4, c = c0 + ((idx & 15) << 2);
                    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (c < fcnt && !p.mom_first) w = *reinterpret_cast<const float4*>(p.mom + (size_t)rr * IN + f0 + c);
                    float* d = wbuf + rr * T4_WLD + (c - c0);
                    d[0] = w.x; d[1] = w.y; d[2] = w.z; d[3] = w.w;
                }
                __syncthreads();
                if (warp < T4_ISSUER) {
#pragma unroll
                    for (int gl = 0; gl < 2; ++gl) {
                        const int g = cb * 4 + half * 2 + gl;
                        if (g * 16 < FP) {
                            float v[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = (j < H) ? wbuf[j * T4_WLD + (half * 2 + gl) * 16 + i] : 0.f;
                            tmem_st16(tlane + C::t_v + g * 16, v);
                        }
                    }
                }
                __syncthreads();
            }
        }
        tmem_st_wait();
    }
    tc_fence_before();
    gb_cluster_sync();            // peers are running: their shared memory may be written from here on
    tc_fence_after();

    const float decay = MOM ? 1.f : 1.f - p.lr * p.wd;     // momentum: weight decay enters the buffer (d = g + wd*w), no lazy scale
    const int spe = (n + B - 1) / B;
    const bool profiling = p.dbg != nullptr && p.lr != 0.f;
    unsigned prof[T4_NPROF];
#pragma unroll
    for (int i = 0; i < T4_NPROF; ++i) prof[i] = 0u;
    unsigned tprev = 0u;
#define T4_STAMP(i) do { if (profiling && lane == 0 && (warp == 0 || warp == 4 || warp == T4_ISSUER)) { const unsigned t_ = (unsigned)clock(); prof[i] += t_ - tprev; tprev = t_; } } while (0)

    const uint32_t rs_bytes = 8u * (uint32_t)H * 16u;
    const uint32_t ag_bytes = rs_bytes + (uint32_t)T4_B * T4_DZP * 4u;

    if (warp == T4_ISSUER) {
        // =========================== MMA / bulk-copy issuer ======================================================
        const uint32_t idesc_fwd = make_idesc(kFmtTF32, kFmtTF32, 128, T4_B, false, false);
        const uint32_t idesc_upd = make_idesc(kFmtTF32, kFmtTF32, 128, FP, false, false);
        const uint32_t x_sbo = (uint32_t)nchunk * 128u;
        const uint32_t d1 = tmem_u + C::t_d1, w1 = tmem_u + C::t_w1, wlo = tmem_u + C::t_wlo;
        const uint32_t gacc = tmem_u + (X3 ? C::t_g : C::t_w1);
        const int ksteps = (FPC + 7) >> 3;                    // columns beyond FPC are zero padding
        const uint32_t idesc_fwd2 = make_idesc(kFmtTF32, kFmtTF32, 128, 2 * T4_B, false, false);
        // forward chains of step s that only need the master weights: D1a = W . Xhi^T (and D1b = W . Xlo^T)
        auto fwd_ab = [&](int s) {
            mbar_wait(&mbar[0], (uint32_t)(s & 1));               // X tile(s) of step s have landed
            tc_fence_after();
            if (elect_one()) {
                mbar_expect_tx(&mbar[5], rs_bytes);               // this step's exchanges (4 and 9 are armed by their consumers)
                mbar_expect_tx(&mbar[6], ag_bytes);
                // the lo image follows the hi image at 4 row groups x SBO, so [Xhi; Xlo] is ONE 64-row K-major operand:
                // D1a | D1b = W . [Xhi; Xlo]^T with half the instructions (N = 64; issue cost ~32 cycles per MMA)
                const uint64_t bhi = make_sdesc(smem_u32(xf), 128u, x_sbo);
#pragma unroll 4
                for (int k = 0; k < ksteps; ++k)                   // +256 B per K step = +16 in the address field
                    mma_tf32_ts(d1, w1 + (uint32_t)k * 8u, bhi + (uint64_t)(k * 16), X3 ? idesc_fwd2 : idesc_fwd, k > 0);
                if (!X3) mma_commit(&mbar[2]);
            }
            __syncwarp();
        };
        fwd_ab(0);
        for (int s = 0; s < total_steps; ++s) {
            const uint32_t ph = (uint32_t)(s & 1);
            if (profiling && lane == 0) tprev = (unsigned)clock();
            if (X3 && s == 0) {                                   // (steps s > 0: the whole forward was issued at the end of step s-1)
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t bhi = make_sdesc(smem_u32(xf), 128u, x_sbo);
#pragma unroll 4
                    for (int k = 0; k < ksteps; ++k)
                        mma_tf32_ts(d1 + (MOM ? 0u : 64u), wlo + (uint32_t)k * 8u, bhi + (uint64_t)(k * 16), idesc_fwd, MOM ? true : k > 0);
                    mma_commit(&mbar[2]);
                }
                __syncwarp();
            }
            T4_STAMP(0);
            mbar_wait(&mbar[2], ph);                              // forward retired -> X buffer is free
            if (s + 1 < total_steps) {
                if (elect_one()) {
                    mbar_expect_tx(&mbar[0], NIMG * tile_bytes);
                    bulk_g2s4(xf, my_stage + (size_t)(s + 1) * NC * blk, NIMG * tile_bytes, &mbar[0]);
                }
                __syncwarp();
            }
            T4_STAMP(1);
            mbar_wait(&mbar[7], ph);                              // dz1 operand images written by the compute warps
            T4_STAMP(2);
            mbar_wait(&mbar[1], ph);                              // X^T tile(s) landed
            tc_fence_after();
            T4_STAMP(6);
            if (elect_one()) {
                const uint64_t ahi = make_sdesc(smem_u32(a2), 128u, 1024u);
                const uint64_t bhi = make_sdesc(smem_u32(xt), 128u, 1024u);
#pragma unroll
                for (int k = 0; k < T4_B / 8; ++k)                 // X3: G starts from zero; tf32: straight into the master
                    mma_tf32_ss(gacc, ahi + (uint64_t)(k * 16), bhi + (uint64_t)(k * 16), idesc_upd, X3 ? k > 0 : true);
                if (X3) {
                    const uint64_t alo = make_sdesc(smem_u32(a2) + T4_HP * T4_B * 4, 128u, 1024u);
                    const uint64_t blo = make_sdesc(smem_u32(xt) + tile_bytes, 128u, 1024u);
#pragma unroll
                    for (int k = 0; k < T4_B / 8; ++k)
                        mma_tf32_ss(gacc, ahi + (uint64_t)(k * 16), blo + (uint64_t)(k * 16), idesc_upd, true);
#pragma unroll
                    for (int k = 0; k < T4_B / 8; ++k)
                        mma_tf32_ss(gacc, alo + (uint64_t)(k * 16), bhi + (uint64_t)(k * 16), idesc_upd, true);
                }
                mma_commit(&mbar[3]);
            }
            __syncwarp();
            T4_STAMP(3);
            if (X3) {
                // the next forward reads the master weights, which the compute warps rewrite (W += G) once G is
                // complete: D1a / D1b of step s+1 are issued when that pass has finished, together with D1c
                mbar_wait(&mbar[3], ph);
                if (s + 1 < total_steps) {
                    if (elect_one()) {
                        mbar_expect_tx(&mbar[1], NIMG * tile_bytes);
                        bulk_g2s4(xt, my_stage + (size_t)(s + 1) * NC * blk + NIMG * tile_floats, NIMG * tile_bytes, &mbar[1]);
                    }
                    __syncwarp();
                    // forward of step s+1, pipelined with the W += G pass: the compute warps rewrite the first 64 columns of
                    // W / Wlo first (mbar[10]) and the rest afterwards (mbar[8]); the MMAs of a K range are issued as soon as
                    // its columns are in place, so the tensor pipe works on the first half while the pass finishes the second
                    mbar_wait(&mbar[0], (uint32_t)((s + 1) & 1));  // X tile(s) of step s+1 have landed
                    // (the momentum variant adds its third chain onto the first chain's accumulator: interleaving the two K
                    // ranges would change the order of that sum, so it keeps the unpipelined order: kh = 0)
                    const int kh = MOM ? 0 : (ksteps < 8 ? ksteps : 8);
                    mbar_wait(&mbar[10], ph);
                    tc_fence_after();
                    T4_STAMP(4);
                    if (elect_one()) {
                        mbar_expect_tx(&mbar[5], rs_bytes);
                        mbar_expect_tx(&mbar[6], ag_bytes);
                        const uint64_t bhi = make_sdesc(smem_u32(xf), 128u, x_sbo);
#pragma unroll 4
                        for (int k = 0; k < kh; ++k)
                            mma_tf32_ts(d1, w1 + (uint32_t)k * 8u, bhi + (uint64_t)(k * 16), idesc_fwd2, k > 0);
#pragma unroll 4
                        for (int k = 0; k < kh; ++k)
                            mma_tf32_ts(d1 + (MOM ? 0u : 64u), wlo + (uint32_t)k * 8u, bhi + (uint64_t)(k * 16), idesc_fwd, MOM ? true : k > 0);
                    }
                    __syncwarp();
                    mbar_wait(&mbar[8], ph);                      // all of W, Wlo of step s+1 written
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t bhi = make_sdesc(smem_u32(xf), 128u, x_sbo);
                        for (int k = kh; k < ksteps; ++k)
                            mma_tf32_ts(d1, w1 + (uint32_t)k * 8u, bhi + (uint64_t)(k * 16), idesc_fwd2, k > 0);
                        for (int k = kh; k < ksteps; ++k)
                            mma_tf32_ts(d1 + (MOM ? 0u : 64u), wlo + (uint32_t)k * 8u, bhi + (uint64_t)(k * 16), idesc_fwd, MOM ? true : k > 0);
                        mma_commit(&mbar[2]);
                    }
                    __syncwarp();
                }
                T4_STAMP(5);
            } else {
                if (s + 1 < total_steps) fwd_ab(s + 1);           // queued behind the update in the tensor pipe
                T4_STAMP(4);
                mbar_wait(&mbar[3], ph);                          // update retired -> X^T buffer is free
                if (s + 1 < total_steps) {
                    if (elect_one()) {
                        mbar_expect_tx(&mbar[1], NIMG * tile_bytes);
                        bulk_g2s4(xt, my_stage + (size_t)(s + 1) * NC * blk + NIMG * tile_floats, NIMG * tile_bytes, &mbar[1]);
                    }
                    __syncwarp();
                }
                T4_STAMP(5);
            }
        }
    } else {
        // =========================== compute warps ===============================================================
        constexpr int JS = C::JS;
        float sscale = 1.f;                                  // W_true = sscale * W_tmem (lazy weight decay)
        const bool reducer = half < GPO;                     // owns (hidden unit j, samples rank*S + 4*half .. +3)
        // loop-invariant DSMEM addresses
        uint32_t rs_dst[4], rs_bar[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int sg = 4 * half + q;                          // float4 sample group 0..7
            const uint32_t owner = (uint32_t)(sg / GPO);
            rs_dst[q] = gb_map_shared(rsb + (((size_t)rank * GPO + (sg % GPO)) * T4_HP + j) * 4, owner);
            rs_bar[q] = gb_map_shared(&mbar[5], owner);
        }
        const int my_g = (int)rank * GPO + (reducer ? half : 0);  // global float4 sample group of my owned samples
        uint32_t ag_dst[NC], ag_bar[NC];
#pragma unroll
        for (int d = 0; d < NC; ++d) {
            ag_dst[d] = gb_map_shared(agb + ((size_t)my_g * T4_HP + j) * 4, (uint32_t)d);
            ag_bar[d] = gb_map_shared(&mbar[6], (uint32_t)d);
        }
        const uint32_t hs_dst = gb_map_shared(hsl + ((size_t)my_g * JS + (j % JS)) * 4, (uint32_t)(j / JS));
        const uint32_t hs_bar = gb_map_shared(&mbar[9], (uint32_t)(j / JS));
        // softmax lanes (warp 0): owned sample si -> destination CTA sd
        const int sm_si = lane / NC, sm_d = lane % NC;
        const uint32_t dz_dst = gb_map_shared(dzb + (size_t)((int)rank * S + sm_si) * T4_DZP, (uint32_t)sm_d);
        const uint32_t dz_bar = gb_map_shared(&mbar[6], (uint32_t)sm_d);
        // my slice of W2: hidden units rank*JS .. rank*JS + JS-1
        const int js_valid = max(0, min(JS, H - (int)rank * JS));
        const uint32_t hs_bytes = (uint32_t)js_valid * 8u * 16u;
        const uint32_t w2_bytes = (uint32_t)T4_OUTV * (uint32_t)((H + 3) >> 2) * 16u;   // every group of 4 hidden units, all classes
        if (tid == 0) {
            mbar_expect_tx(&mbar[9], hs_bytes);
            mbar_expect_tx(&mbar[4], w2_bytes);
        }
        // momentum buffers of the small parameters: b1 (every thread of hidden unit j), my W2 group, b2 (warp 5)
        float vb1 = 0.f, vb2 = 0.f;
        float4 vw2 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (MOM && !p.mom_first) {
            if (j < H) vb1 = p.mom[off_b1 + j];
            if (warp == 5 && lane < OUT) vb2 = p.mom[off_b2 + lane];
            if (tid < ((js_valid + 3) >> 2) * T4_OUTV) {
                const int jq = tid / T4_OUTV, o = tid - jq * T4_OUTV, jg = (int)rank * JS + 4 * jq;
                if (o < OUT) {
                    const float* mp = p.mom + off_w2 + (size_t)o * H + jg;
                    vw2 = make_float4(jg < H ? mp[0] : 0.f, jg + 1 < H ? mp[1] : 0.f, jg + 2 < H ? mp[2] : 0.f, jg + 3 < H ? mp[3] : 0.f);
                }
            }
        }
        // one scalar SGD step with torch.optim.SGD semantics: returns the new parameter, updates the buffer
        auto sgd_mom = [&](float w, float g, float& v, bool first) -> float {
            const float d = fmaf(p.wd, w, g);
            v = first ? d : fmaf(p.momentum, v, (1.f - p.dampening) * d);
            return w - p.lr * (p.nesterov ? fmaf(p.momentum, v, d) : v);
        };
        // K3: partition ids (4 bits each) of the master-weight entries this thread adds G to, and of its small parameters
        uint32_t pidw[SC ? 8 : 1];
        uint32_t pid_b1 = 0u, pid_b2 = 0u, pid_w2 = 0u;
        if (SC) {
#pragma unroll
            for (int q = 0; q < 8; ++q) pidw[q] = 0u;
            const int ngrp = FP >> 4;
            int gi = 0;
            for (int g = half; g < ngrp; g += 2, ++gi) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int col = g * 16 + i;
                    uint32_t pid = 0u;
                    if (j < H && col < fcnt) pid = (uint32_t)p.part_id[(size_t)j * IN + f0 + col] & 15u;
                    if (gi < 4) pidw[2 * gi + (i >> 3)] |= pid << (4 * (i & 7));
                }
            }
            if (j < H) pid_b1 = (uint32_t)p.part_id[off_b1 + j] & 15u;
            if (warp == 5 && lane < OUT) pid_b2 = (uint32_t)p.part_id[off_b2 + lane] & 15u;
            if (tid < ((js_valid + 3) >> 2) * T4_OUTV) {
                const int jq = tid / T4_OUTV, o = tid - jq * T4_OUTV;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int jg = (int)rank * JS + 4 * jq + u;
                    if (jg < H && o < OUT) pid_w2 |= ((uint32_t)p.part_id[off_w2 + (size_t)o * H + jg] & 15u) << (4 * u);
                }
            }
        }

        for (int s = 0; s < total_steps; ++s) {
            const int par = s & 1;
            const uint32_t ph = (uint32_t)(s & 1);
            const int pos = p.epochs > 0 ? (s % spe) * B : 0;
            const int bcur = min(B, n - pos);
            if (profiling && lane == 0) tprev = (unsigned)clock();

            // (1) partial z1 of my feature slice -> registers; reduce-scatter over the cluster
            mbar_wait(&mbar[2], ph);
            tc_fence_after();
            T4_STAMP(0);
            {
                float acc[16];
                tmem_ld16(tlane + C::t_d1 + 16 * half, acc);
                if (X3 && MOM) {
                    float accb[16];
                    tmem_ld16(tlane + C::t_d1 + 32 + 16 * half, accb);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc[i] += accb[i];
                } else if (X3) {
                    float accb[16], accc[16];
                    tmem_ld16(tlane + C::t_d1 + 32 + 16 * half, accb);
                    tmem_ld16(tlane + C::t_d1 + 64 + 16 * half, accc);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc[i] += accb[i] + accc[i];
                } else {
                    tmem_ld_wait();
                }
                if (j < H) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        st_async_v4(rs_dst[q], make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]), rs_bar[q]);
                }
            }
            T4_STAMP(1);
            if (s > 0) {                                         // this step's W2 (all-gathered slices of step s-1)
                mbar_wait_dsm(&mbar[4], (uint32_t)((s - 1) & 1));
                if (tid == 0) mbar_expect_tx(&mbar[4], w2_bytes);
            }
            mbar_wait_dsm(&mbar[5], ph);                     // every CTA's partial sums of my samples landed
            T4_STAMP(2);
            // (2) owner work: h of my S samples (all hidden units), their logits / softmax / dz2 / dz1
            float4 h4 = make_float4(0.f, 0.f, 0.f, 0.f);
            float w2r[T4_OUTV];
            if (reducer) {
                float v[40];
                if (j < H) {
                    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int src = 0; src < NC; ++src) {            // fixed order: bit-identical on every run
                        const float4 r4 = *reinterpret_cast<const float4*>(rsb + (((size_t)src * GPO + half) * T4_HP + j) * 4);
                        z.x += r4.x; z.y += r4.y; z.z += r4.z; z.w += r4.w;
                    }
                    h4 = make_float4(fmaxf(fmaf(sscale, z.x, b1r), 0.f), fmaxf(fmaf(sscale, z.y, b1r), 0.f),
                                     fmaxf(fmaf(sscale, z.z, b1r), 0.f), fmaxf(fmaf(sscale, z.w, b1r), 0.f));
                    st_async_v4(hs_dst, h4, hs_bar);               // -> the CTA that owns column j of W2
                }
                const float hv[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
                for (int o = 0; o < T4_OUTV; ++o) {
                    w2r[o] = w2s[o * T4_HP + j];
#pragma unroll
                    for (int bl = 0; bl < 4; ++bl) v[bl * T4_OUTV + o] = hv[bl] * w2r[o];
                }
                warp_reduce40(v, lane);
                if ((lane & 3) == 0) {
                    const int g = lane >> 2;
#pragma unroll
                    for (int i = 0; i < 5; ++i) red[warp * 40 + 5 * g + i] = v[i];
                }
            }
            T4_STAMP(3);
            bar_compute();
            if (warp == 0) {
                const int hh = sm_si >> 2, bl = sm_si & 3;
                const int b = (int)rank * S + sm_si;
                float z[T4_OUTV], dzv[T4_DZP];
#pragma unroll
                for (int o = 0; o < T4_DZP; ++o) dzv[o] = 0.f;
                if (b < bcur) {
#pragma unroll
                    for (int o = 0; o < T4_OUTV; ++o) {
                        const float* rp = red + (hh * 4) * 40 + bl * T4_OUTV + o;
                        z[o] = (rp[0] + rp[40]) + (rp[80] + rp[120]) + b2s[o];
                    }
                    float m = z[0];
#pragma unroll
                    for (int o = 1; o < T4_OUTV; ++o) m = fmaxf(m, z[o]);
                    float sum = 0.f;
#pragma unroll
                    for (int o = 0; o < T4_OUTV; ++o) { z[o] = __expf(z[o] - m); sum += z[o]; }
                    const float inv = 1.f / sum, invb = 1.f / (float)bcur;
                    const int yy = ysm[par * T4_B + b];
#pragma unroll
                    for (int o = 0; o < T4_OUTV; ++o) dzv[o] = (z[o] * inv - (o == yy ? 1.f : 0.f)) * invb;
                }
                if (sm_d == 0) {                                 // local copy for the dh of my own samples
#pragma unroll
                    for (int q = 0; q < 3; ++q)
                        *reinterpret_cast<float4*>(dzo + sm_si * T4_DZP + 4 * q) = make_float4(dzv[4 * q], dzv[4 * q + 1], dzv[4 * q + 2], dzv[4 * q + 3]);
                }
#pragma unroll
                for (int q = 0; q < 3; ++q)
                    st_async_v4(dz_dst + (uint32_t)(q * 16), make_float4(dzv[4 * q], dzv[4 * q + 1], dzv[4 * q + 2], dzv[4 * q + 3]), dz_bar);
            }
            T4_STAMP(4);
            bar_compute();
            if (reducer && j < H) {                              // dh = dz2 . W2[:, j], ReLU mask -> dz1 of my samples, to all CTAs
                const float hv[4] = {h4.x, h4.y, h4.z, h4.w};
                float dz1[4];
#pragma unroll
                for (int bl = 0; bl < 4; ++bl) {
                    const float4* dr = reinterpret_cast<const float4*>(dzo + (4 * half + bl) * T4_DZP);   // warp-wide broadcast
                    const float4 d0 = dr[0], d1 = dr[1], d2 = dr[2];
                    float dh = d0.x * w2r[0];
                    dh = fmaf(d0.y, w2r[1], dh); dh = fmaf(d0.z, w2r[2], dh); dh = fmaf(d0.w, w2r[3], dh);
                    dh = fmaf(d1.x, w2r[4], dh); dh = fmaf(d1.y, w2r[5], dh); dh = fmaf(d1.z, w2r[6], dh);
                    dh = fmaf(d1.w, w2r[7], dh); dh = fmaf(d2.x, w2r[8], dh); dh = fmaf(d2.y, w2r[9], dh);
                    dz1[bl] = (hv[bl] > 0.f) ? dh : 0.f;
                }
                const float4 dv = make_float4(dz1[0], dz1[1], dz1[2], dz1[3]);
#pragma unroll
                for (int d = 0; d < NC; ++d) st_async_v4(ag_dst[d], dv, ag_bar[d]);
            }
            T4_STAMP(5);
            // (3) all-gather landed: dz1 and dz2 of all 32 samples -> operand images of the update MMA
            mbar_wait_dsm(&mbar[6], ph);
            T4_STAMP(6);
            const float s_next = sscale * decay;
            // With one sample group per owner (NC = 8) only half 0 reduces and sends; half 1 writes the operand images
            // of all 32 samples.
            constexpr bool kSplitRoles = (GPO == 1);
            if (!kSplitRoles || half == 1) {
                const float ascale = MOM ? 1.f : -p.lr / s_next;      // momentum: G is the raw gradient
                float gb1 = 0.f;
                constexpr int NQ = kSplitRoles ? 8 : 4;
                const int q0 = kSplitRoles ? 0 : 4 * half;
                // A2[hid = j][batch] K-major core matrices: ((j/8)*8 + b/4)*128 B + (j%8)*16 B + (b%4)*4 B
                float* arow = a2 + (size_t)(j >> 3) * (8 * 32) + (j & 7) * 4 + q0 * 32;
                float4 d[NQ];
#pragma unroll
                for (int q = 0; q < NQ; ++q)
                    d[q] = (j < H) ? *reinterpret_cast<const float4*>(agb + ((size_t)(q0 + q) * T4_HP + j) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int q = 0; q < NQ; ++q) gb1 += (d[q].x + d[q].y) + (d[q].z + d[q].w);
                gb1p[(kSplitRoles ? 0 : half) * T4_HP + j] = gb1;
                if (kSplitRoles) gb1p[T4_HP + j] = 0.f;
                T4_STAMP(12);
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    float4 a = make_float4(ascale * d[q].x, ascale * d[q].y, ascale * d[q].z, ascale * d[q].w);
                    if (p.dbg != nullptr && !profiling && s == 0 && rank == 0) {     // bring-up: dz1 of the first step
                        float* o = p.dbg + j * T4_B + 4 * (q0 + q);
                        o[0] = d[q].x; o[1] = d[q].y; o[2] = d[q].z; o[3] = d[q].w;
                    }
                    if (X3) {
                        const float4 hi = make_float4(tf32_hi(a.x), tf32_hi(a.y), tf32_hi(a.z), tf32_hi(a.w));
                        *reinterpret_cast<float4*>(arow + T4_HP * T4_B + q * 32) = make_float4(a.x - hi.x, a.y - hi.y, a.z - hi.z, a.w - hi.w);
                        a = hi;
                    }
                    *reinterpret_cast<float4*>(arow + q * 32) = a;
                }
                T4_STAMP(13);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&mbar[7]);            // -> the issuer starts the update MMAs
            }
            T4_STAMP(7);
            // (4) off the critical path: my slice of W2 (all samples), db2, labels of the next step
            if (tid >= 64 && tid < 96 && s + 1 < total_steps)
                ysm[(par ^ 1) * T4_B + (tid - 64)] = stage_ys[(size_t)(s + 1) * T4_B + (tid - 64)];
            if (SC && tid >= 96 && tid < 96 + p.n_parts)          // 1/age of the next step (ages grow by one per step)
                inva[16 * (par ^ 1) + (tid - 96)] = 1.f / (float)(p.age_of(tid - 96) + (int64_t)s + 2);
            if (warp == 5) {                                     // db2[o] = sum_b dz2[b][o]: lane = sample
                const float4* dr = reinterpret_cast<const float4*>(dzb + (size_t)lane * T4_DZP);
                const float4 d0 = dr[0], d1 = dr[1], d2 = dr[2];
                float dv[T4_OUTV] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w, d2.x, d2.y};
#pragma unroll
                for (int o = 0; o < T4_OUTV; ++o) dv[o] = gb_warp_sum(dv[o]);
                if (lane < OUT) {
                    float g = dv[0];
#pragma unroll
                    for (int o = 1; o < T4_OUTV; ++o) g = (lane == o) ? dv[o] : g;
                    if (MOM) b2s[lane] = sgd_mom(b2s[lane], g, vb2, p.mom_first && s == 0);
                    else b2s[lane] = fmaf(SC ? -p.lr * inva[16 * par + pid_b2] : -p.lr, g, b2s[lane] * decay);
                }
            }
            if (js_valid > 0) {
                mbar_wait_dsm(&mbar[9], ph);                 // h of my W2 columns, all 32 samples
                // 4 hidden units x 1 class per thread: the new W2 entries travel as 16-byte st.async (4-byte ones cost
                // 4x the DSMEM transactions and mbarrier updates, which slow every shared-memory access of the cluster)
                const int ngr = (js_valid + 3) >> 2;
                for (int e = tid; e < ngr * T4_OUTV; e += T4_CTHREADS) {
                    const int jq = e / T4_OUTV, o = e - jq * T4_OUTV;
                    float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int g8 = 0; g8 < 8; ++g8) {                // fixed order
                        const float dz0 = dzb[(4 * g8) * T4_DZP + o], dz1v = dzb[(4 * g8 + 1) * T4_DZP + o];
                        const float dz2v = dzb[(4 * g8 + 2) * T4_DZP + o], dz3 = dzb[(4 * g8 + 3) * T4_DZP + o];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float4 hv4 = *reinterpret_cast<const float4*>(hsl + ((size_t)g8 * JS + 4 * jq + u) * 4);
                            g[u] = fmaf(hv4.x, dz0, g[u]); g[u] = fmaf(hv4.y, dz1v, g[u]);
                            g[u] = fmaf(hv4.z, dz2v, g[u]); g[u] = fmaf(hv4.w, dz3, g[u]);
                        }
                    }
                    const int jg = (int)rank * JS + 4 * jq;
                    float* slot = w2s + o * T4_HP + jg;
                    const float4 wo = *reinterpret_cast<const float4*>(slot);
                    float l0 = -p.lr, l1 = -p.lr, l2 = -p.lr, l3 = -p.lr;
                    if (SC) {          // (one W2 group per thread: e == tid, see the packing above)
                        const float* ia = inva + 16 * par;
                        l0 *= ia[pid_w2 & 15u]; l1 *= ia[(pid_w2 >> 4) & 15u]; l2 *= ia[(pid_w2 >> 8) & 15u]; l3 *= ia[(pid_w2 >> 12) & 15u];
                    }
                    float4 wn;
                    if (MOM) {
                        const bool first = p.mom_first && s == 0;
                        wn = make_float4(jg < H ? sgd_mom(wo.x, g[0], vw2.x, first) : 0.f, jg + 1 < H ? sgd_mom(wo.y, g[1], vw2.y, first) : 0.f,
                                         jg + 2 < H ? sgd_mom(wo.z, g[2], vw2.z, first) : 0.f, jg + 3 < H ? sgd_mom(wo.w, g[3], vw2.w, first) : 0.f);
                    } else {
                        wn = make_float4(jg < H ? fmaf(l0, g[0], wo.x * decay) : 0.f,
                                         jg + 1 < H ? fmaf(l1, g[1], wo.y * decay) : 0.f,
                                         jg + 2 < H ? fmaf(l2, g[2], wo.z * decay) : 0.f,
                                         jg + 3 < H ? fmaf(l3, g[3], wo.w * decay) : 0.f);
                    }
#pragma unroll
                    for (int d = 0; d < NC; ++d)
                        st_async_v4(gb_map_shared(slot, (uint32_t)d), wn, gb_map_shared(&mbar[4], (uint32_t)d));
                }
            }
            T4_STAMP(8);
            // (5) update retired: W += G (round to nearest) and re-split into hi / lo for the next forward pass
            mbar_wait(&mbar[3], ph);
            tc_fence_after();
            T4_STAMP(9);
            if (X3) {
                const int ngrp = FP >> 4;                        // 16-column groups; this thread takes g = half, half+2, ...
                const bool more = s + 1 < total_steps;
                for (int g = half; g < ngrp; g += 4) {               // two groups per round: four loads in flight
                    if (g >= 4 && g < 8) {                           // round 1 rewrote columns 0..63: the issuer may start
                        tmem_st_wait();                              // the forward MMAs of that K range (pipelined forward)
                        tc_fence_before();
                        __syncwarp();
                        if (more && lane == 0) mbar_arrive(&mbar[10]);
                    }
                    const bool two = g + 2 < ngrp;
                    float wv[16], gv[16], wv2[16], gv2[16];
                    tmem_ld16(tlane + C::t_w1 + g * 16, wv);
                    tmem_ld16(tlane + C::t_g + g * 16, gv);
                    if (two) {
                        tmem_ld16(tlane + C::t_w1 + (g + 2) * 16, wv2);
                        tmem_ld16(tlane + C::t_g + (g + 2) * 16, gv2);
                    }
                    tmem_ld_wait();
                    if (MOM) {
                        float vv[16];
                        tmem_ld16(tlane + C::t_v + g * 16, vv);
                        tmem_ld_wait();
                        const bool first = p.mom_first && s == 0;
#pragma unroll
                        for (int i = 0; i < 16; ++i) wv[i] = sgd_mom(wv[i], gv[i], vv[i], first);
                        tmem_st16(tlane + C::t_v + g * 16, vv);
                    } else if (SC) {
                        const float* ia = inva + 16 * par;
                        const int gi = (g - half) >> 1;
                        const uint32_t pa = gi == 0 ? pidw[0] : gi == 2 ? pidw[4] : 0u, pb = gi == 0 ? pidw[1] : gi == 2 ? pidw[5] : 0u;
#pragma unroll
                        for (int i = 0; i < 16; ++i) wv[i] = fmaf(gv[i], ia[((i < 8 ? pa : pb) >> (4 * (i & 7))) & 15u], wv[i]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) wv[i] += gv[i];
                    }
                    tmem_st16(tlane + C::t_w1 + g * 16, wv);
                    if (more) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) gv[i] = wv[i] - tf32_hi(wv[i]);
                        tmem_st16(tlane + C::t_wlo + g * 16, gv);
                    }
                    if (two) {
                        if (MOM) {
                            float vv[16];
                            tmem_ld16(tlane + C::t_v + (g + 2) * 16, vv);
                            tmem_ld_wait();
                            const bool first = p.mom_first && s == 0;
#pragma unroll
                            for (int i = 0; i < 16; ++i) wv2[i] = sgd_mom(wv2[i], gv2[i], vv[i], first);
                            tmem_st16(tlane + C::t_v + (g + 2) * 16, vv);
                        } else if (SC) {
                            const float* ia = inva + 16 * par;
                            const int gi = ((g - half) >> 1) + 1;
                            const uint32_t pa = gi == 1 ? pidw[2] : gi == 3 ? pidw[6] : 0u, pb = gi == 1 ? pidw[3] : gi == 3 ? pidw[7] : 0u;
#pragma unroll
                            for (int i = 0; i < 16; ++i) wv2[i] = fmaf(gv2[i], ia[((i < 8 ? pa : pb) >> (4 * (i & 7))) & 15u], wv2[i]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i) wv2[i] += gv2[i];
                        }
                        tmem_st16(tlane + C::t_w1 + (g + 2) * 16, wv2);
                        if (more) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) gv2[i] = wv2[i] - tf32_hi(wv2[i]);
                            tmem_st16(tlane + C::t_wlo + (g + 2) * 16, gv2);
                        }
                    }
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (more && lane == 0) {
                    if (half + 4 >= ngrp) mbar_arrive(&mbar[10]);     // this warp had a single round (FP <= 64 or the last odd group)
                    mbar_arrive(&mbar[8]);
                }
            }
            T4_STAMP(10);
            bar_compute();                                       // end of step: gb1p complete, buffers of this step consumed
            T4_STAMP(11);
            if (tid == 0) mbar_expect_tx(&mbar[9], hs_bytes);    // next step's h slices
            if (MOM) b1r = sgd_mom(b1r, gb1p[j] + gb1p[T4_HP + j], vb1, p.mom_first && s == 0);
            else b1r = fmaf(SC ? -p.lr * inva[16 * par + pid_b1] : -p.lr, gb1p[j] + gb1p[T4_HP + j], b1r * decay);
            sscale = s_next;
        }
        // the last step's W2 all-gather
        mbar_wait_dsm(&mbar[4], (uint32_t)((total_steps - 1) & 1));
        if (tid == 0) red[0] = sscale;
        if (rank == 0 && j < H && half == 0) b1g[j] = b1r;
        if (MOM) {                                           // momentum buffers of the small parameters -> the handler's row
            if (rank == 0 && j < H && half == 0) p.mom[off_b1 + j] = vb1;
            if (rank == 0 && warp == 5 && lane < OUT) p.mom[off_b2 + lane] = vb2;
            if (tid < ((js_valid + 3) >> 2) * T4_OUTV) {
                const int jq = tid / T4_OUTV, o = tid - jq * T4_OUTV, jg = (int)rank * JS + 4 * jq;
                if (o < OUT) {
                    float* mp = p.mom + off_w2 + (size_t)o * H + jg;
                    if (jg < H) mp[0] = vw2.x;
                    if (jg + 1 < H) mp[1] = vw2.y;
                    if (jg + 2 < H) mp[2] = vw2.z;
                    if (jg + 3 < H) mp[3] = vw2.w;
                }
            }
        }
    }

    // ---- every MMA has retired (all compute threads waited for the last update): write everything back -----------
    __syncthreads();
    tc_fence_after();
    {
        const float sscale = red[0];
        float* wbuf = a2;
        for (int cb = 0; cb * T4_WCB < FP; ++cb) {
            const int c0 = cb * T4_WCB;
            if (warp < T4_ISSUER) {
#pragma unroll
                for (int gl = 0; gl < 2; ++gl) {
                    const int g = cb * 4 + half * 2 + gl;
                    if (g * 16 < FP) {
                        float v[16];
                        tmem_ld16(tlane + C::t_w1 + g * 16, v);
                        tmem_ld_wait();
                        if (j < H) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) wbuf[j * T4_WLD + (half * 2 + gl) * 16 + i] = sscale * v[i];
                        }
                    }
                }
            }
            __syncthreads();
            for (int idx = tid; idx < H * (T4_WCB / 4); idx += T4_THREADS) {
                const int rr = idx >> 4, c = c0 + ((idx & 15) << 2);
                if (c < fcnt) {
                    const float* sp = wbuf + rr * T4_WLD + (c - c0);
                    *reinterpret_cast<float4*>(p.row + (size_t)rr * IN + f0 + c) = make_float4(sp[0], sp[1], sp[2], sp[3]);
                }
            }
            __syncthreads();
        }
    }
    if (MOM) {                                               // TMEM tile V -> the momentum row (same path as W)
        float* wbuf = a2;
        for (int cb = 0; cb * T4_WCB < FP; ++cb) {
            const int c0 = cb * T4_WCB;
            if (warp < T4_ISSUER) {
#pragma unroll
                for (int gl = 0; gl < 2; ++gl) {
                    const int g = cb * 4 + half * 2 + gl;
                    if (g * 16 < FP) {
                        float v[16];
                        tmem_ld16(tlane + C::t_v + g * 16, v);
                        tmem_ld_wait();
                        if (j < H) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) wbuf[j * T4_WLD + (half * 2 + gl) * 16 + i] = v[i];
                        }
                    }
                }
            }
            __syncthreads();
            for (int idx = tid; idx < H * (T4_WCB / 4); idx += T4_THREADS) {
                const int rr = idx >> 4, c = c0 + ((idx & 15) << 2);
                if (c < fcnt) {
                    const float* sp = wbuf + rr * T4_WLD + (c - c0);
                    *reinterpret_cast<float4*>(p.mom + (size_t)rr * IN + f0 + c) = make_float4(sp[0], sp[1], sp[2], sp[3]);
                }
            }
            __syncthreads();
        }
    }
    if (rank == 0) {
        for (int i = tid; i < OUT * H; i += T4_THREADS) W2g[i] = w2s[(i / H) * T4_HP + (i % H)];
        if (tid < OUT) b2g[tid] = b2s[tid];
    }
    if (profiling && lane == 0 && (warp == 0 || warp == 4 || warp == T4_ISSUER)) {
        for (int i = 0; i < T4_NPROF; ++i)
            p.dbg[rank * 64 + (warp == 0 ? 0 : warp == 4 ? 16 : 32) + i] = (float)((double)prof[i] / (double)total_steps);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
    gb_cluster_sync();            // no CTA may exit while peers can still write into its shared memory
#undef T4_STAMP
}

// ---- device-side loader: shuffled mini-batches as UMMA operand images -----------------------------------------
// One CTA per (step, cluster rank).  Output block of (s, r):  XF images then XT images, each image
// T4_B x FP floats; with 3xTF32 every operand has a hi image (x & 0xffffe000) and a lo image (x - hi).
//   XF  forward operand X_s[:, slice r]   (N = batch rows, K = features)  [4 batch groups][FP/4 chunks][8 rows][16 B]
//   XT  update operand  X_s[:, slice r]^T (N = features,   K = batch)     [FP/8 feature groups][8 batch chunks][8 rows][16 B]
template <bool X3>
__global__ void __launch_bounds__(256)
mlp1_stage4_kernel(const float* __restrict__ X, const int64_t* __restrict__ y, int n, int IN, int B, int epochs,
                   uint64_t key, int NC, int FPC, int FP, float* __restrict__ out, int* __restrict__ ys) {
    constexpr int NIMG = X3 ? 2 : 1;
    extern __shared__ __align__(16) float tile[];       // [32][FP + 1]
    __shared__ int ids[T4_B];
    const int s = blockIdx.x, r = blockIdx.y, tid = threadIdx.x;
    const int spe = (n + B - 1) / B;
    const int e = epochs > 0 ? s / spe : 0;
    const int pos = epochs > 0 ? (s % spe) * B : 0;
    const int bcur = min(B, n - pos);
    const int f0 = r * FPC;
    const int fcnt = max(0, min(FPC, IN - f0));
    const int ld = FP + 1;
    if (tid < T4_B) {
        int id = -1;
        if (tid < bcur) {
            GbPerm perm; perm.init((uint32_t)n, gb_mix64(key ^ (uint64_t)e));
            id = (int)perm((uint32_t)(pos + tid));
        }
        ids[tid] = id;
        if (r == 0) ys[(size_t)s * T4_B + tid] = id >= 0 ? (int)y[id] : -1;
    }
    __syncthreads();
    {   // gather: 128-bit loads, four in flight per thread (random rows: a chain of DRAM round trips otherwise)
        const int nch = FP >> 2, total = T4_B * nch;
        constexpr int U = 4;
        for (int i0 = tid; i0 < total; i0 += U * 256) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * 256;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < total) {
                    const int b = i / nch, c = (i - b * nch) << 2;
                    const int id = ids[b];
                    if (id >= 0 && c < fcnt) v[u] = __ldg(reinterpret_cast<const float4*>(X + (size_t)id * IN + f0 + c));
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * 256;
                if (i < total) {
                    const int b = i / nch, c = (i - b * nch) << 2;
                    float* d = tile + b * ld + c;
                    d[0] = v[u].x; d[1] = v[u].y; d[2] = v[u].z; d[3] = v[u].w;
                }
            }
        }
    }
    __syncthreads();
    const size_t tile_floats = (size_t)T4_B * FP;
    float* blk = out + ((size_t)s * NC + r) * (2 * NIMG * tile_floats);
    auto put = [&](float4* hi, float4* lo, int q, float a, float b, float c, float d) {
        if (X3) {
            const float ah = tf32_hi(a), bh = tf32_hi(b), ch = tf32_hi(c), dh = tf32_hi(d);
            hi[q] = make_float4(ah, bh, ch, dh);
            lo[q] = make_float4(a - ah, b - bh, c - ch, d - dh);
        } else {
            hi[q] = make_float4(a, b, c, d);
        }
    };
    {   // XF: chunk q = ((g * nchunk + c) * 8 + r8) holds X[g*8 + r8][4c .. 4c+3]
        const int nchunk = FP >> 2;
        float4* hi = reinterpret_cast<float4*>(blk);
        float4* lo = reinterpret_cast<float4*>(blk + tile_floats);
        for (int q = tid; q < T4_B * nchunk; q += 256) {
            const int r8 = q & 7, gc = q >> 3, c = gc % nchunk, g = gc / nchunk;
            const float* src = tile + (g * 8 + r8) * ld + 4 * c;
            put(hi, lo, q, src[0], src[1], src[2], src[3]);
        }
    }
    {   // XT: chunk q = ((fg * 8 + bc) * 8 + fr) holds X[4bc .. 4bc+3][fg*8 + fr]
        float4* hi = reinterpret_cast<float4*>(blk + NIMG * tile_floats);
        float4* lo = reinterpret_cast<float4*>(blk + (NIMG + 1) * tile_floats);
        const int nq = (FP >> 3) * 64;
        for (int q = tid; q < nq; q += 256) {
            const int fr = q & 7, bc = (q >> 3) & 7, fg = q >> 6;
            const float* src = tile + (bc * 4) * ld + fg * 8 + fr;
            put(hi, lo, q, src[0], src[ld], src[2 * ld], src[3 * ld]);
        }
    }
}

static void tc4_geometry(int IN, int NC, int* FPC, int* FP) {
    *FPC = (((IN + NC - 1) / NC) + 3) & ~3;
    *FP = (*FPC + 15) & ~15;
}

size_t mlp1_stage4_bytes(int n, int IN, int B, int epochs, int NC, bool x3, int* FPC_out, int* FP_out, int* steps_out) {
    int FPC, FP;
    tc4_geometry(IN, NC, &FPC, &FP);
    const int spe = (n + B - 1) / B;
    const int steps = epochs > 0 ? epochs * spe : 1;
    if (FPC_out) *FPC_out = FPC;
    if (FP_out) *FP_out = FP;
    if (steps_out) *steps_out = steps;
    const size_t tile = (size_t)T4_B * FP * 4;
    return (size_t)steps * ((size_t)NC * 2 * (x3 ? 2 : 1) * tile + T4_B * 4);
}

// staging buffers: one per (device, stream), sized by reserve_train_staging() (init_nodes) or grown on first use
struct StageBuf4 { void* ptr; size_t bytes; cudaStream_t stream; int dev; };
static StageBuf4 g_stage4[1024] = {};

static void* stage_buffer_for4(cudaStream_t stream, size_t bytes, bool may_alloc) {
    int dev = 0;
    cudaGetDevice(&dev);
    int free_slot = -1;
    for (int i = 0; i < 1024; ++i) {
        StageBuf4& sb = g_stage4[i];
        if (sb.ptr != nullptr && sb.stream == stream && sb.dev == dev) {
            if (sb.bytes >= bytes) return sb.ptr;
            if (!may_alloc) return nullptr;
            cudaStreamSynchronize(stream);
            cudaFree(sb.ptr);
            sb.ptr = nullptr;
            free_slot = i;
            break;
        }
        if (sb.ptr == nullptr && free_slot < 0) free_slot = i;
    }
    if (free_slot < 0 || !may_alloc) return nullptr;
    void* ptr = nullptr;
    if (cudaMalloc(&ptr, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    g_stage4[free_slot] = StageBuf4{ptr, bytes, stream, dev};
    return ptr;
}

template <int NC, bool X3, bool SC, bool MOM = false>
static bool tc4_launch(const TrainParams& p, cudaStream_t stream) {
    using C = T4Cfg<NC, X3, MOM>;
    if (MOM && (p.mom == nullptr || p.momentum <= 0.f)) return false;
    if (p.H > T4_HP || p.OUT > T4_OUTV || p.B > T4_B || p.IN % 4 != 0) return false;
    int FPC, FP, steps;
    const size_t bytes = mlp1_stage4_bytes(p.n, p.IN, p.B, p.epochs, NC, X3, &FPC, &FP, &steps);
    if (FP > C::FP_MAX || (NC - 1) * FPC >= p.IN) return false;               // every CTA needs a non-empty slice
    if ((double)steps * (double)p.lr * (double)p.wd > 20.0) return false;    // lazy decay scale would underflow
    if (bytes > ((size_t)1 << 32)) return false;
    void* staging = stage_buffer_for4(stream, bytes, true);
    if (staging == nullptr) return false;
    float* out = static_cast<float*>(staging);
    const size_t tile_floats = (size_t)T4_B * FP;
    int* ys = reinterpret_cast<int*>(out + (size_t)steps * NC * 2 * C::NIMG * tile_floats);
    {
        const size_t smem = (size_t)T4_B * (FP + 1) * 4;
        static size_t configured = 0;
        if (smem > configured) {
            if (cudaFuncSetAttribute(mlp1_stage4_kernel<X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
                return false;
            configured = smem;
        }
        if (p.stage_mode != 2)
            mlp1_stage4_kernel<X3><<<dim3(steps, NC), 256, smem, stream>>>(p.X, p.y, p.n, p.IN, p.B, p.epochs, p.key, NC, FPC, FP, out, ys);
    }
    if (SC && (p.n_parts > 16 || !p.scaled())) return false;
    if (p.stage_mode == 1) return cudaGetLastError() == cudaSuccess;
    auto kern = mlp1_train_tc4_kernel<NC, X3, SC, MOM>;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::total + 1024) != cudaSuccess) return false;
        configured = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(NC); cfg.blockDim = dim3(T4_THREADS); cfg.dynamicSmemBytes = C::total + 1024; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NC; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, p, FPC, FP, steps, (const float*)out, (const int*)ys) == cudaSuccess;
}

// nc in {4, 8}; x3 = error-compensated (fp32-equivalent) products, otherwise plain tf32
bool mlp1_train_tc4(const TrainParams& p, int nc, bool x3, cudaStream_t stream) {
    if (nc != 8) return false;      // (the kernel is written for NC in {4, 8}; only the 8-CTA form is validated and shipped)
    const bool scaled = p.scaled();
    if (p.momentum != 0.f) return x3 && !scaled && tc4_launch<8, true, false, true>(p, stream);
    if (scaled) return x3 && tc4_launch<8, true, true>(p, stream);      // K3 rides on the W += G pass of the X3 form
    return x3 ? tc4_launch<8, true, false>(p, stream) : tc4_launch<8, false, false>(p, stream);
}

// allocate the staging buffer of `stream` ahead of time (init_nodes): no cudaMalloc -- an implicit device-wide
// barrier -- may happen on the launch path while kernels spin on other GPUs' flags
bool reserve_train_staging(int n, int IN, int B, int epochs, int nc, bool x3, cudaStream_t stream) {
    const size_t bytes = mlp1_stage4_bytes(n, IN, B, epochs, nc, x3, nullptr, nullptr, nullptr);
    return stage_buffer_for4(stream, bytes, true) != nullptr;
}

// force-load this file's kernels (CUDA loads functions lazily; loading one while another kernel spins
// on a cross-GPU flag could deadlock, so the extension loads everything up front)
void preload_train_tc4() {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, mlp1_train_tc4_kernel<8, true, false, false>);
    cudaFuncGetAttributes(&a, mlp1_train_tc4_kernel<8, true, true, false>);
    cudaFuncGetAttributes(&a, mlp1_train_tc4_kernel<8, true, false, true>);
    cudaFuncGetAttributes(&a, mlp1_train_tc4_kernel<8, false, false, false>);
    cudaFuncGetAttributes(&a, mlp1_stage4_kernel<true>);
    cudaFuncGetAttributes(&a, mlp1_stage4_kernel<false>);
}

}  // namespace gb
