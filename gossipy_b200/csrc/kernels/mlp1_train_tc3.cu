// tcgen05 / TMEM training kernel, CTA-pair variant ("tc3", plain tf32 products: opt-in, see mlp1_train_tc4.cu for the
// fp32-equivalent default): the whole SGD step incl. the SECOND LAYER on
// the tensor core as well.  Per SGD step the CTA pair issues five GEMMs:
//
//   fwd   z1^T[128 x 32]  = W1(TMEM, fp32 master) . X^T                       (M128 N32  K392, TS mode)
//   d2    z2  [64  x 16]  = h[64(b) x 128(j)] . W2[16(o) x 128(j)]^T           (M64  N16  K128, SW128 operands)
//   dh    dh^T[128 x 32]  = W2^T[128(j) x 16(o)] . dz2[32(b) x 16(o)]^T        (M128 N32  K16)
//   gw2   gW2^T[128 x 16] = h^T[128(j) x 32(b)] . dz2^T[16(o) x 32(b)]^T       (M128 N16  K32)
//   upd   W1[128 x 392]  += (-lr/s dz1)^T[128 x 32] . X                        (M128 N392 K32, into the master)
//
// so the CUDA cores only do element-wise work: bias + ReLU, the 32 softmaxes, the ReLU mask, and the
// 1 110 second-layer parameter updates.  Every thread keeps the tile rows it already owns: thread
// (j, half) reads row j of dh^T / gW2^T straight out of TMEM.  Operand images in shared memory:
// h (b-major, 128-byte swizzle so the 16 per-thread stores are conflict free), h^T, W2 (K = j, swizzled),
// W2^T (K = o), dz2 and dz2^T (no-swizzle core matrices).  Everything else (loader, bulk copies,
// st.async exchange, lazy weight decay, fused MERGE_UPDATE): see the sections below.
// Reference semantics: gossipy/model/handler.py:235-258.  tf32 products, fp32 accumulation/master.
#include "tc_common.cuh"
#include "kernels.h"

namespace gb {

constexpr int T3_THREADS = 256;
constexpr int T3_B = 32;          // mini-batch tile
constexpr int T3_HP = 128;        // hidden units padded to the MMA M
constexpr int T3_OUTP = 16;
constexpr int T3_FP_MAX = 400;    // feature columns per CTA (TMEM: 400 + 32 accumulator columns <= 512)
constexpr int T3_TMEM_COLS = 512;
constexpr int T3_MMA_WARP = 4;    // issues the forward / update MMAs and their bulk copies (warp 0 issues the small second-layer GEMMs)
constexpr int T3_WCB = 64;        // weight columns moved per pass of the TMEM fill / write-back
constexpr int T3_WLD = 65;        // row pitch of the [128][64] scratch (odd: conflict-free column access)

// shared-memory images of the second-layer operands (float offsets)
GB_DEVICE int sw128_off(int rows, int r, int k) {    // K-major, 128-byte swizzle: element (row r, k)
    return ((k >> 5) * (rows >> 3) + (r >> 3)) * 256 + (r & 7) * 32 + ((((k & 31) >> 2) ^ (r & 7)) << 2) + (k & 3);
}
GB_DEVICE int kmaj_off(int kchunks, int r, int k) {  // K-major core matrices, no swizzle: element (row r, k)
    return ((r >> 3) * kchunks + (k >> 2)) * 32 + (r & 7) * 4 + (k & 3);
}

struct T3Smem {   // byte offsets inside dynamic shared memory (base rounded up to 1024 B by hand)
    static constexpr int xf = 0;
    static constexpr int tile_bytes = T3_B * T3_FP_MAX * 4;          // 51200
    static constexpr int xt = xf + tile_bytes;
    static constexpr int a2 = xt + tile_bytes;                       // dz1^T operand of the update MMA [128 x 32]
    static constexpr int zpart = a2 + T3_HP * T3_B * 4;              // [2][8][128][4] peer partial z1
    static constexpr int hA = zpart + 2 * T3_HP * T3_B * 4;          // h  [64(b) x 128(j)], SW128      (1024-aligned)
    static constexpr int w2k = hA + 64 * T3_HP * 4;                  // W2 [16(o) x 128(j)], SW128      (1024-aligned)
    static constexpr int hT = w2k + 16 * T3_HP * 4;                  // h^T  [128(j) x 32(b)]
    static constexpr int w2t = hT + T3_HP * T3_B * 4;                // W2^T [128(j) x 16(o)]
    static constexpr int dz = w2t + T3_HP * 16 * 4;                  // dz2   [32(b) x 16(o)]
    static constexpr int dzT = dz + T3_B * 16 * 4;                   // dz2^T [16(o) x 32(b)]
    static constexpr int gb1p = dzT + 16 * T3_B * 4;                 // [2][128] per-half partials of db1
    static constexpr int b2 = gb1p + 2 * T3_HP * 4;
    static constexpr int ys = b2 + 16 * 4;                           // [2][32] int labels
    static constexpr int mbar = ys + 2 * T3_B * 4;                   // 10 x uint64
    static constexpr int tslot = mbar + 96;
    static constexpr int total = tslot + 16;
};
static_assert(T3Smem::hA % 1024 == 0 && T3Smem::w2k % 1024 == 0, "swizzled operands need 1024-byte alignment");
static_assert(T3Smem::total + 1024 <= 227 * 1024, "shared memory budget");
static_assert(T3_HP * T3_WLD * 4 <= T3Smem::hA - T3Smem::a2, "fill / write-back scratch must fit in a2 + zpart");
static_assert(T3_WCB == 64, "index arithmetic below assumes 16 float4 per scratch row and 4 column groups per pass");


GB_DEVICE void bulk_g2s3(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(mbar)) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T3_THREADS, 1)
mlp1_train_tc3_kernel(const TrainParams p, const int FPC, const int FP, const int total_steps,
                      const float* __restrict__ stage_xf, const float* __restrict__ stage_xt,
                      const int* __restrict__ stage_ys) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // round the base up to 1024 B on the SHARED-window address and keep `smem` derived from `smem_raw`:
    // going through uintptr_t loses the address space and turns every LDS/STS below into a generic LD/ST
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int quad = warp & 3, half = warp >> 2;
    const int j = quad * 32 + lane;                        // hidden unit = TMEM lane of this thread
    const uint32_t rank = gb_cluster_ctarank(), peer_cta = rank ^ 1u;
    const int IN = p.IN, H = p.H, OUT = p.OUT, n = p.n, B = p.B;
    const int f0 = (int)rank * FPC;
    const int fcnt = max(0, min(FPC, IN - f0));
    const int nchunk = FP >> 2;

    float* xf = reinterpret_cast<float*>(smem + T3Smem::xf);
    float* xt = reinterpret_cast<float*>(smem + T3Smem::xt);
    float* a2 = reinterpret_cast<float*>(smem + T3Smem::a2);
    float* zpart = reinterpret_cast<float*>(smem + T3Smem::zpart);
    float* hA = reinterpret_cast<float*>(smem + T3Smem::hA);
    float* w2k = reinterpret_cast<float*>(smem + T3Smem::w2k);
    float* hT = reinterpret_cast<float*>(smem + T3Smem::hT);
    float* w2t = reinterpret_cast<float*>(smem + T3Smem::w2t);
    float* dzs = reinterpret_cast<float*>(smem + T3Smem::dz);
    float* dzT = reinterpret_cast<float*>(smem + T3Smem::dzT);
    float* gb1p = reinterpret_cast<float*>(smem + T3Smem::gb1p);
    float* b2s = reinterpret_cast<float*>(smem + T3Smem::b2);
    int* ysm = reinterpret_cast<int*>(smem + T3Smem::ys);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + T3Smem::mbar);   // 0 xf, 1 xt, 2 fwd, 3 upd, 4 drain, 5/6 exchange, 7 d2, 8 dh, 9 gw2
    uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + T3Smem::tslot);

    float* b1g = p.row + (size_t)H * IN;
    float* W2g = b1g + H;
    float* b2g = W2g + (size_t)OUT * H;
    const uint32_t tile_bytes = (uint32_t)T3_B * (uint32_t)FP * 4u;
    const size_t tile_floats = (size_t)T3_B * FP;
    const float* my_xf = stage_xf + (size_t)rank * tile_floats;          // + s * 2 * tile_floats
    const float* my_xt = stage_xt + (size_t)rank * tile_floats;

    // ---- one-time set-up -------------------------------------------------------------------------
    // Fused MERGE_UPDATE.  With a peer row the starting point is w_self*row + w_peer*peer: the CTA pair
    // first streams the peer's row -- possibly out of ANOTHER GPU's HBM, after spinning on its `ready`
    // flag -- with coalesced 128-bit loads (NVLink moves 32-byte sectors: the per-thread 4-byte pattern of
    // the TMEM fill below would fetch every sector eight times), folds it into the own row in place,
    // acknowledges the read on the owner's `done` counter, and only then loads the weights on chip.
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
        const int64_t stride = 2 * T3_THREADS;
        int64_t i = (int64_t)rank * T3_THREADS + tid;
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
        gb_cluster_sync();                                       // the merged row is visible to both CTAs
        if (p.sync.done != nullptr && rank == 0 && tid == 0) gb_red_release_sys_add(p.sync.done, 1u);
    }
    auto ldp = [&](size_t off) -> float { return p.row[off]; };
    const size_t off_b1 = (size_t)H * IN, off_w2 = off_b1 + H, off_b2 = off_w2 + (size_t)OUT * H;
    if (warp == 0) tmem_alloc<T3_TMEM_COLS>(tslot);
    if (tid == 0) {
        for (int i = 0; i < 10; ++i) mbar_init(&mbar[i], 1);
        mbar_fence_init();
    }
    for (int i = tid; i < 64 * T3_HP; i += T3_THREADS) hA[i] = 0.f;          // rows 32..63 stay zero (MMA M = 64)
    for (int i = tid; i < T3_HP * T3_B; i += T3_THREADS) hT[i] = 0.f;
    for (int i = tid; i < 2 * T3_B * 16; i += T3_THREADS) dzs[i] = 0.f;      // dz2 and dz2^T (contiguous)
    for (int i = tid; i < T3_OUTP * T3_HP; i += T3_THREADS) {
        const int o = i / T3_HP, jj = i % T3_HP;
        const float v = (o < OUT && jj < H) ? ldp(off_w2 + (size_t)o * H + jj) : 0.f;
        w2k[sw128_off(16, o, jj)] = v;
        w2t[kmaj_off(4, jj, o)] = v;
    }
    float b1r = (j < H) ? ldp(off_b1 + j) : 0.f;           // both threads of hidden unit j carry its bias
    if (tid < T3_OUTP) b2s[tid] = (tid < OUT) ? ldp(off_b2 + tid) : -3.0e38f;   // padding classes: logit -inf, probability 0
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tslot;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);          // warp-uniform copy for the MMA issuer
    const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);
    const uint32_t exch_bytes = (uint32_t)H * T3_B * 4u;                // peer partial sums per step (rows < H)
    const uint32_t t_w1 = 0, t_d1 = T3_FP_MAX, t_d2 = T3_FP_MAX + 32, t_dh = T3_FP_MAX + 48, t_gw = T3_FP_MAX + 80;

    // first X / X^T tiles and labels: issue now, they land while the weights are loaded
    if (tid == 0) {
        mbar_expect_tx(&mbar[0], tile_bytes);
        bulk_g2s3(xf, my_xf, tile_bytes, &mbar[0]);
        mbar_expect_tx(&mbar[1], tile_bytes);
        bulk_g2s3(xt, my_xt, tile_bytes, &mbar[1]);
    }
    if (tid < T3_B) ysm[tid] = stage_ys[tid];

    // master weights -> TMEM.  Thread (j, half) owns TMEM lane j, but reading "its" weight row straight from
    // global memory makes every warp-wide load touch 32 different cache lines; instead the CTA moves
    // 64-column blocks through shared memory: coalesced 128-bit loads (16 consecutive threads = 256
    // contiguous bytes of one row) -> [128][65] scratch (a2 + zpart, unused until the cluster barrier
    // below) -> conflict-free column reads -> tcgen05.st.
    {
        float* wbuf = a2;
        for (int cb = 0; cb * T3_WCB < T3_FP_MAX; ++cb) {
            const int c0 = cb * T3_WCB;
            for (int idx = tid; idx < H * (T3_WCB / 4); idx += T3_THREADS) {
                const int rr = idx >> 4, c = c0 + ((idx & 15) << 2);
                float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c < fcnt) w = *reinterpret_cast<const float4*>(p.row + (size_t)rr * IN + f0 + c);
                float* d = wbuf + rr * T3_WLD + (c - c0);
                d[0] = w.x; d[1] = w.y; d[2] = w.z; d[3] = w.w;
            }
            __syncthreads();
#pragma unroll
            for (int gl = 0; gl < 2; ++gl) {
                const int g = cb * 4 + half * 2 + gl;                     // warp-uniform
                if (g < T3_FP_MAX / 16) {
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = (j < H) ? wbuf[j * T3_WLD + (half * 2 + gl) * 16 + i] : 0.f;
                    tmem_st16(tlane + t_w1 + g * 16, v);
                }
            }
            __syncthreads();
        }
        tmem_st_wait();
        for (int i = tid; i < T3_HP * T3_B; i += T3_THREADS) a2[i] = 0.f;
    }
    tc_fence_before();
    gb_cluster_sync();            // peer is running (its smem may be written from here on)
    tc_fence_after();

    const float decay = 1.f - p.lr * p.wd;
    float sscale = 1.f;                                  // W_true = sscale * W_tmem
    const uint32_t idesc_fwd = make_idesc(kFmtTF32, kFmtTF32, 128, T3_B, false, false);
    const uint32_t x_sbo = (uint32_t)nchunk * 128u;
    const int spe = (n + B - 1) / B;
    unsigned prof[17];                                   // fine-grained phase counters (thread 0, profiling runs only)
#pragma unroll
    for (int i = 0; i < 17; ++i) prof[i] = 0u;
    unsigned tprev = 0u;
#define T3_STAMP(i) do { if (profiling && tid == 0) { const unsigned t_ = (unsigned)clock(); prof[i] += t_ - tprev; tprev = t_; } } while (0)
    const bool profiling = p.dbg != nullptr && p.lr == 0.f ? false : (p.dbg != nullptr);

    for (int s = 0; s < total_steps; ++s) {
        const int par = s & 1;
        const uint32_t ph = (uint32_t)(s & 1);
        const int pos = p.epochs > 0 ? (s % spe) * B : 0;
        const int bcur = min(B, n - pos);
        if (profiling && tid == 0) tprev = (unsigned)clock();

        // (A)+(B) forward MMA: D1[128 x 32] = W1(TMEM) . X^T ; queued behind update(s-1).
        // Warp 0 stays converged and one ELECTED lane issues (operands in uniform registers).
        if (warp == T3_MMA_WARP) {
            mbar_wait(&mbar[0], ph);                     // X tile of this step has landed
            tc_fence_after();
            if (elect_one()) {
                mbar_expect_tx(&mbar[5 + par], exch_bytes);   // the peer's partial sums of this step
                const uint64_t bdesc0 = make_sdesc(smem_u32(xf), 128u, x_sbo);
                const uint32_t d1 = tmem_u + t_d1, a0 = tmem_u + t_w1;
                mma_tf32_ts(d1, a0, bdesc0, idesc_fwd, false);
#pragma unroll 7
                for (int k = 1; k < FP / 8; ++k)          // +256 B per K step = +16 in the address field
                    mma_tf32_ts(d1, a0 + (uint32_t)k * 8u, bdesc0 + (uint64_t)(k * 16), idesc_fwd, true);
                mma_commit(&mbar[2]);
            }
            __syncwarp();
            if (s > 0) {                                 // update(s-1) retired -> X^T buffer is free
                mbar_wait(&mbar[3], (uint32_t)((s - 1) & 1));
                if (elect_one()) {
                    mbar_expect_tx(&mbar[1], tile_bytes);
                    bulk_g2s3(xt, my_xt + (size_t)s * 2 * tile_floats, tile_bytes, &mbar[1]);
                }
                __syncwarp();
            }
        }
        T3_STAMP(0);

        // (C) accumulator -> registers, exchange partial sums with the peer CTA
        mbar_wait(&mbar[2], ph);
        tc_fence_after();
        if (warp == T3_MMA_WARP && s + 1 < total_steps) {   // forward MMA retired -> X buffer is free
            if (elect_one()) {
                mbar_expect_tx(&mbar[0], tile_bytes);
                bulk_g2s3(xf, my_xf + (size_t)(s + 1) * 2 * tile_floats, tile_bytes, &mbar[0]);
            }
            __syncwarp();
        }
        float acc[16];
        tmem_ld16(tlane + t_d1 + 16 * half, acc);
        tmem_ld_wait();
        T3_STAMP(1);
        // my partial sums -> the peer's zpart; layout [par][half*4 + q][j][4 samples]: every warp-wide
        // store covers 512 contiguous bytes of the peer's shared memory, each 16-B piece signalling
        // the peer's exchange mbarrier (st.async complete_tx); rows j >= H carry nothing
        float* zslot = zpart + (((size_t)par * 8 + half * 4) * T3_HP + j) * 4;
        if (j < H) {
            const uint32_t remote = gb_map_shared(zslot, peer_cta);
            const uint32_t rbar = gb_map_shared(&mbar[5 + par], peer_cta);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                st_async_v4(remote + (uint32_t)(q * T3_HP * 16), make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]), rbar);
        }
        T3_STAMP(2);
        mbar_wait_cluster(&mbar[5 + par], (uint32_t)((s >> 1) & 1));   // all of the peer's partials landed
        T3_STAMP(3);
        float h[16];
        {
            const float bj = b1r;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 o = (j < H) ? *reinterpret_cast<const float4*>(zslot + (size_t)q * T3_HP * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                h[4 * q] = acc[4 * q] + o.x; h[4 * q + 1] = acc[4 * q + 1] + o.y;
                h[4 * q + 2] = acc[4 * q + 2] + o.z; h[4 * q + 3] = acc[4 * q + 3] + o.w;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float z = fmaf(sscale, h[i], bj);
                h[i] = (j < H) ? fmaxf(z, 0.f) : 0.f;
                hA[sw128_off(64, 16 * half + i, j)] = h[i];             // lanes = consecutive j: conflict free
            }
            float* trow = hT + (size_t)(j >> 3) * (8 * 32) + (j & 7) * 4 + (4 * half) * 32;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                *reinterpret_cast<float4*>(trow + q * 32) = make_float4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
        }
        T3_STAMP(4);
        fence_proxy_async();
        tc_fence_before();
        if (p.dbg != nullptr && !profiling && s == 0 && rank == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) p.dbg[j * T3_B + 16 * half + i] = h[i];
        }
        __syncthreads();
        T3_STAMP(5);

        // (D) logits on the tensor core: z2[64 x 16] = h . W2^T (M64 N16 K128, both operands 128B-swizzled)
        if (warp == 0) {
            tc_fence_after();
            if (elect_one()) {
                const uint32_t idesc_d2 = make_idesc(kFmtTF32, kFmtTF32, 64, 16, false, false);
                const uint32_t ha = smem_u32(hA), wk = smem_u32(w2k);
                const int ksteps = (H + 7) >> 3;        // columns j >= H of both operands are zero
                // K step k covers j = 8k .. 8k+7: swizzle atom k/4 (8 KB apart in h, 2 KB in W2), 32 B inside the atom;
                // fully unrolled so that the descriptors are base + constant
                const uint64_t ad0 = make_sdesc_sw128(ha, 16u, 1024u), bd0 = make_sdesc_sw128(wk, 16u, 1024u);
#pragma unroll
                for (int k = 0; k < T3_HP / 8; ++k)
                    if (k < ksteps)
                        mma_tf32_ss(tmem_u + t_d2, ad0 + (uint64_t)((k >> 2) * 512 + (k & 3) * 2),
                                    bd0 + (uint64_t)((k >> 2) * 128 + (k & 3) * 2), idesc_d2, k > 0);
                mma_commit(&mbar[7]);
            }
            __syncwarp();
        }
        T3_STAMP(6);
        // (E) softmax cross-entropy gradient: M = 64 accumulators put sample b on lane b%16 of quadrant
        // b/16, i.e. lanes 0..15 of warps 0 and 1 own the 32 samples
        if (warp < 2) {
            mbar_wait(&mbar[7], ph);
            tc_fence_after();
            T3_STAMP(7);
            float z[16];
            tmem_ld16(tlane + t_d2, z);
            tmem_ld_wait();
            T3_STAMP(8);
            if (lane < 16) {
                const int b = 16 * warp + lane;
                float dzv[16];
                if (b < bcur) {
                    // straight-line: the bias image carries -3e38 for the padding classes o >= OUT, so they
                    // drop out of the max and get probability exactly 0 without a predicate per class
                    float m = -3.0e38f;
#pragma unroll
                    for (int o = 0; o < 10; ++o) { z[o] += b2s[o]; m = fmaxf(m, z[o]); }
                    float sum = 0.f;
#pragma unroll
                    for (int o = 0; o < 10; ++o) { z[o] = __expf(z[o] - m); sum += z[o]; }
                    const float inv = 1.f / sum, invb = 1.f / (float)bcur;
                    const int yy = ysm[par * T3_B + b];
#pragma unroll
                    for (int o = 0; o < 16; ++o) dzv[o] = (o < 10) ? (z[o] * inv - (o == yy ? 1.f : 0.f)) * invb : 0.f;
                } else {
#pragma unroll
                    for (int o = 0; o < 16; ++o) dzv[o] = 0.f;
                }
                float* drow = dzs + (size_t)(b >> 3) * (4 * 32) + (b & 7) * 4;         // dz2[b][o]: K = o
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<float4*>(drow + q * 32) = make_float4(dzv[4 * q], dzv[4 * q + 1], dzv[4 * q + 2], dzv[4 * q + 3]);
#pragma unroll
                for (int o = 0; o < 16; ++o) dzT[kmaj_off(8, o, b)] = dzv[o];             // dz2^T[o][b]: K = b
            }
            fence_proxy_async();
            T3_STAMP(9);
        }
        if (tid >= 64 && tid < 96 && s + 1 < total_steps)            // labels of the next step
            ysm[(par ^ 1) * T3_B + (tid - 64)] = stage_ys[(size_t)(s + 1) * T3_B + (tid - 64)];
        tc_fence_before();
        __syncthreads();
        T3_STAMP(10);

        // (F) backward GEMMs: dh^T = W2^T . dz2^T (K = 16) and gW2^T = h^T . dz2 (K = 32)
        if (warp == 0) {
            tc_fence_after();
            if (elect_one()) {
                const uint32_t idesc_dh = make_idesc(kFmtTF32, kFmtTF32, 128, 32, false, false);
                const uint32_t idesc_gw = make_idesc(kFmtTF32, kFmtTF32, 128, 16, false, false);
                const uint64_t wt = make_sdesc(smem_u32(w2t), 128u, 512u), dd = make_sdesc(smem_u32(dzs), 128u, 512u);
                const uint64_t ht = make_sdesc(smem_u32(hT), 128u, 1024u), dt = make_sdesc(smem_u32(dzT), 128u, 1024u);
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    mma_tf32_ss(tmem_u + t_dh, wt + (uint64_t)(k * 16), dd + (uint64_t)(k * 16), idesc_dh, k > 0);
                mma_commit(&mbar[8]);                    // dh is on the critical path ...
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    mma_tf32_ss(tmem_u + t_gw, ht + (uint64_t)(k * 16), dt + (uint64_t)(k * 16), idesc_gw, k > 0);
                mma_commit(&mbar[9]);                    // ... gW2 only feeds the off-path second-layer update
            }
            __syncwarp();
        }
        T3_STAMP(11);
        mbar_wait(&mbar[8], ph);
        tc_fence_after();
        T3_STAMP(12);
        const float s_next = sscale * decay;
        float gw2[16];
        float gb1 = 0.f;
        {
            float dh[16];
            tmem_ld16(tlane + t_dh + 16 * half, dh);
            tmem_ld_wait();
            const float ascale = -p.lr / s_next;
            float outv[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float dz1 = (h[i] > 0.f) ? dh[i] : 0.f;
                gb1 += dz1;
                outv[i] = ascale * dz1;
            }
            // A2[hid = j][batch] in K-major core-matrix layout: ((j/8)*8 + b/4)*128 B + (j%8)*16 B + (b%4)*4 B
            float* arow = a2 + (size_t)(j >> 3) * (8 * 32) + (j & 7) * 4 + (4 * half) * 32;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                *reinterpret_cast<float4*>(arow + q * 32) = make_float4(outv[4 * q], outv[4 * q + 1], outv[4 * q + 2], outv[4 * q + 3]);
            gb1p[half * T3_HP + j] = gb1;                // partial bias gradient of my 16 samples
        }
        T3_STAMP(13);
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        T3_STAMP(14);

        // (G) update MMA: W1[128 x FP] += A2[128 x 32] . X^T-tile (both operands K-major, K = batch)
        if (warp == T3_MMA_WARP) {
            mbar_wait(&mbar[1], ph);                     // X^T tile of this step has landed
            tc_fence_after();
            if (elect_one()) {
                const uint64_t adesc0 = make_sdesc(smem_u32(a2), 128u, 1024u);
                const uint64_t bdesc0 = make_sdesc(smem_u32(xt), 128u, 1024u);
                for (int n0 = 0; n0 < FP; n0 += 256) {
                    const int nn = min(256, FP - n0);
                    const uint32_t idesc_upd = make_idesc(kFmtTF32, kFmtTF32, 128, nn, false, false);
                    const uint64_t bn = bdesc0 + (uint64_t)((n0 >> 3) * 64);      // (n0/8) * 1024 B
#pragma unroll
                    for (int k = 0; k < T3_B / 8; ++k)
                        mma_tf32_ss(tmem_u + t_w1 + (uint32_t)n0, adesc0 + (uint64_t)(k * 16), bn + (uint64_t)(k * 16),
                                    idesc_upd, true);
                }
                mma_commit(&mbar[3]);
            }
            __syncwarp();
        }
        T3_STAMP(15);
        // (H) second-layer parameters: thread (j, 0) owns column j of W2 (both operand images)
        mbar_wait(&mbar[9], ph);                         // gW2 retired (every warp: h^T / dz2^T are rewritten next step)
        tc_fence_after();
        if (half == 0) {
            tmem_ld16(tlane + t_gw, gw2);
            tmem_ld_wait();
            if (j < H) {
                float* wrow = w2t + (size_t)(j >> 3) * (4 * 32) + (j & 7) * 4;          // W2^T[j][o], 4 chunks of 4 outputs
                float wn[12];
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const float4 w = *reinterpret_cast<const float4*>(wrow + q * 32);
                    wn[4 * q] = w.x; wn[4 * q + 1] = w.y; wn[4 * q + 2] = w.z; wn[4 * q + 3] = w.w;
                }
#pragma unroll
                for (int o = 0; o < 10; ++o)
                    if (o < OUT) { wn[o] = fmaf(-p.lr, gw2[o], wn[o] * decay); w2k[sw128_off(16, o, j)] = wn[o]; }
#pragma unroll
                for (int q = 0; q < 3; ++q)
                    *reinterpret_cast<float4*>(wrow + q * 32) = make_float4(wn[4 * q], wn[4 * q + 1], wn[4 * q + 2], wn[4 * q + 3]);
            }
        } else if (tid >= 160 && tid < 160 + OUT) {     // (warp 5: warp 4 is busy issuing the update MMAs)
            const int o = tid - 160;
            float gsum = 0.f;
            for (int b = 0; b < T3_B; ++b) gsum += dzs[kmaj_off(4, b, o)];
            b2s[o] = fmaf(-p.lr, gsum, b2s[o] * decay);
        }
        b1r = fmaf(-p.lr, gb1 + gb1p[(half ^ 1) * T3_HP + j], b1r * decay);
        sscale = s_next;
        T3_STAMP(16);
    }

    // ---- drain the tensor pipe and write everything back ------------------------------------------------
    if (warp == T3_MMA_WARP) {
        if (elect_one()) mma_commit(&mbar[4]);
        __syncwarp();
    }
    mbar_wait(&mbar[4], 0u);
    tc_fence_after();
    __syncthreads();
    {   // master weights TMEM -> global through the same [128][65] scratch, written back with coalesced 128-bit stores
        float* wbuf = a2;                                                 // every MMA has retired, the last exchange is consumed
        for (int cb = 0; cb * T3_WCB < T3_FP_MAX; ++cb) {
            const int c0 = cb * T3_WCB;
#pragma unroll
            for (int gl = 0; gl < 2; ++gl) {
                const int g = cb * 4 + half * 2 + gl;
                if (g < T3_FP_MAX / 16) {
                    float v[16];
                    tmem_ld16(tlane + t_w1 + g * 16, v);
                    tmem_ld_wait();
                    if (j < H) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) wbuf[j * T3_WLD + (half * 2 + gl) * 16 + i] = sscale * v[i];
                    }
                }
            }
            __syncthreads();
            for (int idx = tid; idx < H * (T3_WCB / 4); idx += T3_THREADS) {
                const int rr = idx >> 4, c = c0 + ((idx & 15) << 2);
                if (c < fcnt) {
                    const float* sp = wbuf + rr * T3_WLD + (c - c0);
                    *reinterpret_cast<float4*>(p.row + (size_t)rr * IN + f0 + c) = make_float4(sp[0], sp[1], sp[2], sp[3]);
                }
            }
            __syncthreads();
        }
    }
    if (rank == 0) {
        for (int i = tid; i < OUT * H; i += T3_THREADS) { const int o = i / H, jj = i % H; W2g[i] = w2t[kmaj_off(4, jj, o)]; }
        if (half == 0 && j < H) b1g[j] = b1r;
        if (tid < OUT) b2g[tid] = b2s[tid];
    }
    if (profiling && tid == 0) {
        for (int i = 0; i < 17; ++i) p.dbg[rank * 32 + i] = (float)((double)prof[i] / (double)total_steps);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<T3_TMEM_COLS>(tmem);
    gb_cluster_sync();
#undef T3_STAMP
}

// staging buffers: one per (device, stream), grown on demand.  The training kernel that consumes a
// staging buffer is enqueued right behind its loader on the same stream, so successive updates of a
// stream can reuse the buffer; different streams (= different gossip nodes) get different buffers.
struct StageBuf3 { void* ptr; size_t bytes; cudaStream_t stream; int dev; };
static StageBuf3 g_stage3[512] = {};

static void* stage_buffer_for3(cudaStream_t stream, size_t bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    int free_slot = -1;
    for (int i = 0; i < 512; ++i) {
        StageBuf3& sb = g_stage3[i];
        if (sb.ptr != nullptr && sb.stream == stream && sb.dev == dev) {
            if (sb.bytes >= bytes) return sb.ptr;
            cudaStreamSynchronize(stream);
            cudaFree(sb.ptr);
            sb.ptr = nullptr;
            free_slot = i;
            break;
        }
        if (sb.ptr == nullptr && free_slot < 0) free_slot = i;
    }
    if (free_slot < 0) return nullptr;
    void* ptr = nullptr;
    if (cudaMalloc(&ptr, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    g_stage3[free_slot] = StageBuf3{ptr, bytes, stream, dev};
    return ptr;
}

bool mlp1_train_tc3(const TrainParams& p, cudaStream_t stream) {
    // shape envelope; anything else falls back to the first-generation tc kernel / the cluster kernel
    if (p.H > T3_HP || p.OUT > 10 || p.B > T3_B || p.IN % 8 != 0) return false;
    int FPC, FP, steps;
    const size_t bytes = mlp1_stage_bytes(p.n, p.IN, p.B, p.epochs, &FPC, &FP, &steps);
    if (FP > T3_FP_MAX || p.IN - FPC > FPC || p.IN - FPC <= 0) return false;
    if ((double)steps * (double)p.lr * (double)p.wd > 20.0) return false;   // lazy decay scale would underflow
    if (bytes > ((size_t)1 << 31)) return false;
    void* staging = stage_buffer_for3(stream, bytes);
    if (staging == nullptr) return false;
    if (!launch_mlp1_stage(p.X, p.y, p.n, p.IN, p.B, p.epochs, p.key, staging, stream)) return false;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(mlp1_train_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 T3Smem::total + 1024) != cudaSuccess) return false;
        configured = true;
    }
    const size_t tile_floats = (size_t)T3_B * FP;
    const float* xf = static_cast<const float*>(staging);
    const float* xt = xf + (size_t)steps * 2 * tile_floats;
    const int* ys = reinterpret_cast<const int*>(xt + (size_t)steps * 2 * tile_floats);
    mlp1_train_tc3_kernel<<<2, T3_THREADS, T3Smem::total + 1024, stream>>>(p, FPC, FP, steps, xf, xt, ys);
    return cudaGetLastError() == cudaSuccess;
}

// force-load this file's kernels (CUDA loads functions lazily; loading one while another kernel spins
// on a cross-GPU flag could deadlock, so the extension loads everything up front)
void preload_train_tc3() {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, mlp1_train_tc3_kernel);
}

}  // namespace gb
