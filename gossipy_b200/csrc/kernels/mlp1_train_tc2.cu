// tcgen05 / TMEM training kernel, second generation ("tc2"): the default local-update kernel of the
// flagship MLP.  Same idea as mlp1_train_tc.cu -- the fp32 MASTER WEIGHTS of the first layer live in
// Tensor Memory for the whole local epoch, the forward MMA reads them from TMEM (TS mode, tf32) and
// the SGD update is a tcgen05.mma that accumulates INTO them -- with the per-step critical path cut
// to what truly depends on the previous step:
//
//   * operands arrive pre-shuffled and pre-tiled from the device-side loader (mlp1_stage.cu): two
//     contiguous 51 KB bulk copies per step (cp.async.bulk + mbarrier complete_tx), no index math,
//     no per-row gathers and no shared->shared transpose in the training loop;
//   * 256 threads (8 warps): two warps share each TMEM lane quadrant and split the accumulator
//     columns, so the CUDA-core phases (second layer, softmax, backward) run 2x wider;
//   * the X^T tile of step s+1 is requested as soon as the update MMA of step s retires, the X tile
//     as soon as the forward MMA does -- single buffers, copies fully overlapped;
//   * forward MMA of step s+1 is queued right behind the update MMA of step s (the tensor pipe
//     executes in issue order), so the tensor core never waits for a round trip through the CTA.
//
// A CTA pair splits the input features (K of the forward GEMM, N of the update GEMM): each CTA owns
// FPC <= 400 TMEM columns; partial pre-activations are exchanged through distributed shared memory
// (one cluster barrier per step).  Second layer / softmax-CE / backward run redundantly (bit-identical)
// on both CTAs.  Weight decay is a lazy scalar (W_true = s * W_tmem).  With `peer` set the initial
// weights are w_self*row + w_peer*peer (fused MERGE_UPDATE, peer row possibly in another GPU's HBM).
// Reference semantics: gossipy/model/handler.py:235-258.  tf32 products, fp32 accumulation/master.
#include "tc_common.cuh"
#include "kernels.h"

namespace gb {

constexpr int T2_THREADS = 256;
constexpr int T2_B = 32;          // mini-batch tile
constexpr int T2_HP = 128;        // hidden units padded to the MMA M
constexpr int T2_OUTP = 16;
constexpr int T2_FP_MAX = 400;    // feature columns per CTA (TMEM: 400 + 32 accumulator columns <= 512)
constexpr int T2_TMEM_COLS = 512;
constexpr int T2_HS = 132;        // stride of hs[b][j] (sample-major; 132 = 4 mod 32 keeps 128-bit reads conflict free)

struct T2Smem {   // byte offsets inside dynamic shared memory (1024-aligned base)
    static constexpr int xf = 0;
    static constexpr int tile_bytes = T2_B * T2_FP_MAX * 4;          // 51200
    static constexpr int xt = xf + tile_bytes;
    static constexpr int a2 = xt + tile_bytes;                       // [16 hid groups][8 chunks][8][16B]
    static constexpr int zpart = a2 + T2_HP * T2_B * 4;              // [2][128][32] peer partial z1
    static constexpr int hs = zpart + 2 * T2_HP * T2_B * 4;          // [32][132]
    static constexpr int w2 = hs + T2_B * T2_HS * 4;                 // [16][128]
    static constexpr int z2 = w2 + T2_OUTP * T2_HP * 4;              // [32][16] logits -> dL/dz2
    static constexpr int gws = z2 + T2_B * T2_OUTP * 4;              // [128][12] second-half partial grads
    static constexpr int b1 = gws + T2_HP * 12 * 4;
    static constexpr int b2 = b1 + T2_HP * 4;
    static constexpr int ys = b2 + T2_OUTP * 4;                      // [2][32] int labels
    static constexpr int mbar = ys + 2 * T2_B * 4;                   // 8 x uint64
    static constexpr int tslot = mbar + 64;
    static constexpr int total = tslot + 16;
};

GB_DEVICE void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(mbar)) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T2_THREADS, 1)
mlp1_train_tc2_kernel(const TrainParams p, const int FPC, const int FP, const int total_steps,
                      const float* __restrict__ stage_xf, const float* __restrict__ stage_xt,
                      const int* __restrict__ stage_ys) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int quad = warp & 3, half = warp >> 2;
    const int j = quad * 32 + lane;                        // hidden unit = TMEM lane of this thread
    const uint32_t rank = gb_cluster_ctarank(), peer_cta = rank ^ 1u;
    const int IN = p.IN, H = p.H, OUT = p.OUT, n = p.n, B = p.B;
    const int f0 = (int)rank * FPC;
    const int fcnt = max(0, min(FPC, IN - f0));
    const int nchunk = FP >> 2;

    float* xf = reinterpret_cast<float*>(smem + T2Smem::xf);
    float* xt = reinterpret_cast<float*>(smem + T2Smem::xt);
    float* a2 = reinterpret_cast<float*>(smem + T2Smem::a2);
    float* zpart = reinterpret_cast<float*>(smem + T2Smem::zpart);
    float* hs = reinterpret_cast<float*>(smem + T2Smem::hs);
    float* w2s = reinterpret_cast<float*>(smem + T2Smem::w2);
    float* z2s = reinterpret_cast<float*>(smem + T2Smem::z2);
    float* gws = reinterpret_cast<float*>(smem + T2Smem::gws);
    float* b1s = reinterpret_cast<float*>(smem + T2Smem::b1);
    float* b2s = reinterpret_cast<float*>(smem + T2Smem::b2);
    int* ysm = reinterpret_cast<int*>(smem + T2Smem::ys);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + T2Smem::mbar);   // 0 xf, 1 xt, 2 fwd, 3 upd, 4 drain, 5/6 exchange
    uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + T2Smem::tslot);

    float* b1g = p.row + (size_t)H * IN;
    float* W2g = b1g + H;
    float* b2g = W2g + (size_t)OUT * H;
    const uint32_t tile_bytes = (uint32_t)T2_B * (uint32_t)FP * 4u;
    const size_t tile_floats = (size_t)T2_B * FP;
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
        const int64_t stride = 2 * T2_THREADS;
        int64_t i = (int64_t)rank * T2_THREADS + tid;
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
    if (warp == 0) tmem_alloc<T2_TMEM_COLS>(tslot);
    if (tid == 0) {
        for (int i = 0; i < 7; ++i) mbar_init(&mbar[i], 1);
        mbar_fence_init();
    }
    for (int i = tid; i < T2_HP * T2_B; i += T2_THREADS) a2[i] = 0.f;
    for (int i = tid; i < T2_OUTP * T2_HP; i += T2_THREADS) {
        const int o = i / T2_HP, jj = i % T2_HP;
        w2s[i] = (o < OUT && jj < H) ? ldp(off_w2 + (size_t)o * H + jj) : 0.f;
    }
    for (int i = tid; i < T2_B * T2_HS; i += T2_THREADS) hs[i] = 0.f;
    if (tid < T2_HP) b1s[tid] = (tid < H) ? ldp(off_b1 + tid) : 0.f;
    if (tid < T2_OUTP) b2s[tid] = (tid < OUT) ? ldp(off_b2 + tid) : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tslot;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);          // warp-uniform copy for the MMA issuer
    const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);
    const uint32_t exch_bytes = (uint32_t)H * T2_B * 4u;                // peer partial sums per step (rows < H)
    const uint32_t t_w1 = 0, t_d1 = T2_FP_MAX;

    // first X / X^T tiles and labels: issue now, they land while the weights are loaded
    if (tid == 0) {
        mbar_expect_tx(&mbar[0], tile_bytes);
        bulk_g2s(xf, my_xf, tile_bytes, &mbar[0]);
        mbar_expect_tx(&mbar[1], tile_bytes);
        bulk_g2s(xt, my_xt, tile_bytes, &mbar[1]);
    }
    if (tid < T2_B) ysm[tid] = stage_ys[tid];

    // master weights -> TMEM: thread (j, half) fills column groups [half*13, ...) of its lane
    {
        const int ngroups = T2_FP_MAX / 16;                               // 25
        const int g0 = half ? 13 : 0, g1 = half ? ngroups : 13;
        for (int g = g0; g < g1; ++g) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int c = g * 16 + i;
                v[i] = (j < H && c < fcnt) ? ldp((size_t)j * IN + f0 + c) : 0.f;
            }
            tmem_st16(tlane + t_w1 + g * 16, v);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    gb_cluster_sync();            // peer is running (its smem may be written from here on)
    tc_fence_after();

    const float decay = 1.f - p.lr * p.wd;
    float sscale = 1.f;                                  // W_true = sscale * W_tmem
    const uint32_t idesc_fwd = make_idesc(kFmtTF32, kFmtTF32, 128, T2_B, false, false);
    const uint32_t x_sbo = (uint32_t)nchunk * 128u;
    const int spe = (n + B - 1) / B;
    long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const bool profiling = p.dbg != nullptr && p.lr == 0.f ? false : (p.dbg != nullptr);

    for (int s = 0; s < total_steps; ++s) {
        const int par = s & 1;
        const uint32_t ph = (uint32_t)(s & 1);
        const int pos = p.epochs > 0 ? (s % spe) * B : 0;
        const int bcur = min(B, n - pos);
        long long t0 = 0;
        if (profiling && tid == 0) t0 = clock64();

        // (A)+(B) forward MMA: D1[128 x 32] = W1(TMEM) . X^T ; queued behind update(s-1).
        // Warp 0 stays converged and one ELECTED lane issues (operands in uniform registers).
        if (warp == 0) {
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
                    bulk_g2s(xt, my_xt + (size_t)s * 2 * tile_floats, tile_bytes, &mbar[1]);
                }
                __syncwarp();
            }
        }
        if (profiling && tid == 0) { const long long t = clock64(); prof[0] += t - t0; t0 = t; }

        // (C) accumulator -> registers, exchange partial sums with the peer CTA
        mbar_wait(&mbar[2], ph);
        tc_fence_after();
        if (warp == 0 && s + 1 < total_steps) {          // forward MMA retired -> X buffer is free
            if (elect_one()) {
                mbar_expect_tx(&mbar[0], tile_bytes);
                bulk_g2s(xf, my_xf + (size_t)(s + 1) * 2 * tile_floats, tile_bytes, &mbar[0]);
            }
            __syncwarp();
        }
        float acc[16];
        tmem_ld16(tlane + t_d1 + 16 * half, acc);
        tmem_ld_wait();
        if (profiling && tid == 0) { const long long t = clock64(); prof[1] += t - t0; t0 = t; }
        // my partial sums -> the peer's zpart; layout [par][half*4 + q][j][4 samples]: every warp-wide
        // store covers 512 contiguous bytes of the peer's shared memory, each 16-B piece signalling
        // the peer's exchange mbarrier (st.async complete_tx); rows j >= H carry nothing
        float* zslot = zpart + (((size_t)par * 8 + half * 4) * T2_HP + j) * 4;
        if (j < H) {
            const uint32_t remote = gb_map_shared(zslot, peer_cta);
            const uint32_t rbar = gb_map_shared(&mbar[5 + par], peer_cta);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                st_async_v4(remote + (uint32_t)(q * T2_HP * 16), make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]), rbar);
        }
        mbar_wait_cluster(&mbar[5 + par], (uint32_t)((s >> 1) & 1));   // all of the peer's partials landed
        float h[16];
        {
            const float bj = b1s[j];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 o = (j < H) ? *reinterpret_cast<const float4*>(zslot + (size_t)q * T2_HP * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                h[4 * q] = acc[4 * q] + o.x; h[4 * q + 1] = acc[4 * q + 1] + o.y;
                h[4 * q + 2] = acc[4 * q + 2] + o.z; h[4 * q + 3] = acc[4 * q + 3] + o.w;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float z = fmaf(sscale, h[i], bj);
                h[i] = (j < H) ? fmaxf(z, 0.f) : 0.f;
                hs[(16 * half + i) * T2_HS + j] = h[i];
            }
        }
        if (p.dbg != nullptr && !profiling && s == 0 && rank == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) p.dbg[j * T2_B + 16 * half + i] = h[i];
        }
        __syncthreads();
        if (profiling && tid == 0) { const long long t = clock64(); prof[2] += t - t0; t0 = t; }

        // (D)+(E) layer 2 forward and softmax cross-entropy gradient, all 8 warps alike: warp w owns
        // samples 4w..4w+3; lane = (g = hidden group of 16, bl = sample).  Partial dot products over the
        // group, butterfly reduction across the 8 groups (lane bits 2..4), then every lane finishes
        // the softmax of its sample for the outputs o = g and g + 8.
        {
            const int bl = lane & 3, g = lane >> 2, b = 4 * warp + bl;
            float zacc[10];
#pragma unroll
            for (int o = 0; o < 10; ++o) zacc[o] = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int jb = 16 * g + 4 * c;
                const float4 hv = *reinterpret_cast<const float4*>(hs + b * T2_HS + jb);
#pragma unroll
                for (int o = 0; o < 10; ++o) {
                    const float4 wv = *reinterpret_cast<const float4*>(w2s + o * T2_HP + jb);
                    zacc[o] = fmaf(hv.x, wv.x, fmaf(hv.y, wv.y, fmaf(hv.z, wv.z, fmaf(hv.w, wv.w, zacc[o]))));
                }
            }
#pragma unroll
            for (int o = 0; o < 10; ++o) {
                zacc[o] += __shfl_xor_sync(0xffffffffu, zacc[o], 4);
                zacc[o] += __shfl_xor_sync(0xffffffffu, zacc[o], 8);
                zacc[o] += __shfl_xor_sync(0xffffffffu, zacc[o], 16);
            }
            float m = -3.0e38f;
#pragma unroll
            for (int o = 0; o < 10; ++o) { zacc[o] = (o < OUT) ? zacc[o] + b2s[o] : -3.0e38f; m = fmaxf(m, zacc[o]); }
            // my outputs: o1 = g, o2 = g + 8 (only groups 0 and 1 have a second one)
            float z1v = zacc[0], z2v = -3.0e38f;
#pragma unroll
            for (int o = 1; o < 8; ++o) z1v = (g == o) ? zacc[o] : z1v;
            z2v = (g == 0) ? zacc[8] : ((g == 1) ? zacc[9] : z2v);
            const float e1 = (g < OUT) ? __expf(z1v - m) : 0.f;
            const float e2 = (g + 8 < OUT && g < 2) ? __expf(z2v - m) : 0.f;
            float sum = e1 + e2;
            sum += __shfl_xor_sync(0xffffffffu, sum, 4);
            sum += __shfl_xor_sync(0xffffffffu, sum, 8);
            sum += __shfl_xor_sync(0xffffffffu, sum, 16);
            const bool live = b < bcur;
            const float inv = live ? 1.f / sum : 0.f, invb = 1.f / (float)bcur;
            const int yy = ysm[par * T2_B + b];
            float* zr = z2s + b * T2_OUTP;
            zr[g] = (g < OUT && live) ? (e1 * inv - (g == yy ? 1.f : 0.f)) * invb : 0.f;
            zr[g + 8] = (g < 2 && g + 8 < OUT && live) ? (e2 * inv - (g + 8 == yy ? 1.f : 0.f)) * invb : 0.f;
        }
        if (tid >= 32 && tid < 64 && s + 1 < total_steps)           // labels of the next step
            ysm[(par ^ 1) * T2_B + (tid - 32)] = stage_ys[(size_t)(s + 1) * T2_B + (tid - 32)];
        __syncthreads();
        if (profiling && tid == 0) { const long long t = clock64(); prof[3] += t - t0; t0 = t; }

        // (F) backward through layer 2: thread (j, half) handles its 16 samples
        const float s_next = sscale * decay;
        float gw2[10];
        float gb1 = 0.f;
        {
            float w2c[10];
#pragma unroll
            for (int o = 0; o < 10; ++o) { w2c[o] = w2s[o * T2_HP + j]; gw2[o] = 0.f; }
            const float ascale = -p.lr / s_next;
            float outv[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float4* dzr = reinterpret_cast<const float4*>(z2s + (16 * half + i) * T2_OUTP);
                const float4 d0 = dzr[0], d1 = dzr[1], d2 = dzr[2];
                const float dz[10] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w, d2.x, d2.y};
                float dh = 0.f;
#pragma unroll
                for (int o = 0; o < 10; ++o) { dh = fmaf(dz[o], w2c[o], dh); gw2[o] = fmaf(dz[o], h[i], gw2[o]); }
                const float dz1 = (h[i] > 0.f) ? dh : 0.f;
                gb1 += dz1;
                outv[i] = ascale * dz1;
            }
            // A2[hid = j][batch] in K-major core-matrix layout: ((j/8)*8 + b/4)*128 B + (j%8)*16 B + (b%4)*4 B
            float* arow = a2 + (size_t)(j >> 3) * (8 * 32) + (j & 7) * 4 + (4 * half) * 32;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                *reinterpret_cast<float4*>(arow + q * 32) = make_float4(outv[4 * q], outv[4 * q + 1], outv[4 * q + 2], outv[4 * q + 3]);
            if (half == 1) {
                float* g = gws + j * 12;
#pragma unroll
                for (int o = 0; o < 10; ++o) g[o] = gw2[o];
                g[10] = gb1;
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (profiling && tid == 0) { const long long t = clock64(); prof[4] += t - t0; t0 = t; }

        // (G) update MMA: W1[128 x FP] += A2[128 x 32] . X^T-tile (both operands K-major, K = batch)
        if (warp == 0) {
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
                    for (int k = 0; k < T2_B / 8; ++k)
                        mma_tf32_ss(tmem_u + t_w1 + (uint32_t)n0, adesc0 + (uint64_t)(k * 16), bn + (uint64_t)(k * 16),
                                    idesc_upd, true);
                }
                mma_commit(&mbar[3]);
            }
            __syncwarp();
        }
        // (H) second-layer parameters: thread (j, 0) adds the other half's partial gradients
        if (half == 0) {
            const float* g = gws + j * 12;
            if (j < H) {
#pragma unroll
                for (int o = 0; o < 10; ++o)
                    if (o < OUT) w2s[o * T2_HP + j] = fmaf(-p.lr, gw2[o] + g[o], w2s[o * T2_HP + j] * decay);
            }
            b1s[j] = fmaf(-p.lr, gb1 + g[10], b1s[j] * decay);
        } else if (tid >= 128 && tid < 128 + OUT) {
            const int o = tid - 128;
            float gsum = 0.f;
            for (int b = 0; b < T2_B; ++b) gsum += z2s[b * T2_OUTP + o];
            b2s[o] = fmaf(-p.lr, gsum, b2s[o] * decay);
        }
        sscale = s_next;
        if (profiling && tid == 0) { const long long t = clock64(); prof[5] += t - t0; t0 = t; }
    }

    // ---- drain the tensor pipe and write everything back ------------------------------------------------
    if (warp == 0) {
        if (elect_one()) mma_commit(&mbar[4]);
        __syncwarp();
    }
    mbar_wait(&mbar[4], 0u);
    tc_fence_after();
    __syncthreads();
    {
        const int ngroups = T2_FP_MAX / 16;
        const int g0 = half ? 13 : 0, g1 = half ? ngroups : 13;
        for (int g = g0; g < g1; ++g) {
            float v[16];
            tmem_ld16(tlane + t_w1 + g * 16, v);
            tmem_ld_wait();
            if (j < H) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int c = g * 16 + i;
                    if (c < fcnt) p.row[(size_t)j * IN + f0 + c] = sscale * v[i];
                }
            }
        }
    }
    if (rank == 0) {
        for (int i = tid; i < OUT * H; i += T2_THREADS) { const int o = i / H, jj = i % H; W2g[i] = w2s[o * T2_HP + jj]; }
        if (tid < H) b1g[tid] = b1s[tid];
        if (tid < OUT) b2g[tid] = b2s[tid];
    }
    if (profiling && tid == 0 && rank == 0) {
        for (int i = 0; i < 6; ++i) p.dbg[i] = (float)((double)prof[i] / (double)total_steps);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<T2_TMEM_COLS>(tmem);
    gb_cluster_sync();
}

// staging buffers: one per (device, stream), grown on demand.  The training kernel that consumes a
// staging buffer is enqueued right behind its loader on the same stream, so successive updates of a
// stream can reuse the buffer; different streams (= different gossip nodes) get different buffers.
struct StageBuf { void* ptr; size_t bytes; cudaStream_t stream; int dev; };
static StageBuf g_stage[512] = {};

static void* stage_buffer_for(cudaStream_t stream, size_t bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    int free_slot = -1;
    for (int i = 0; i < 512; ++i) {
        StageBuf& sb = g_stage[i];
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
    g_stage[free_slot] = StageBuf{ptr, bytes, stream, dev};
    return ptr;
}

bool mlp1_train_tc2(const TrainParams& p, cudaStream_t stream) {
    // shape envelope; anything else falls back to the first-generation tc kernel / the cluster kernel
    if (p.H > T2_HP || p.OUT > 10 || p.B > T2_B || p.IN % 8 != 0) return false;
    int FPC, FP, steps;
    const size_t bytes = mlp1_stage_bytes(p.n, p.IN, p.B, p.epochs, &FPC, &FP, &steps);
    if (FP > T2_FP_MAX || p.IN - FPC > FPC || p.IN - FPC <= 0) return false;
    if ((double)steps * (double)p.lr * (double)p.wd > 20.0) return false;   // lazy decay scale would underflow
    if (bytes > ((size_t)1 << 31)) return false;
    void* staging = stage_buffer_for(stream, bytes);
    if (staging == nullptr) return false;
    if (!launch_mlp1_stage(p.X, p.y, p.n, p.IN, p.B, p.epochs, p.key, staging, stream)) return false;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(mlp1_train_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 T2Smem::total + 1024) != cudaSuccess) return false;
        configured = true;
    }
    const size_t tile_floats = (size_t)T2_B * FP;
    const float* xf = static_cast<const float*>(staging);
    const float* xt = xf + (size_t)steps * 2 * tile_floats;
    const int* ys = reinterpret_cast<const int*>(xt + (size_t)steps * 2 * tile_floats);
    mlp1_train_tc2_kernel<<<2, T2_THREADS, T2Smem::total + 1024, stream>>>(p, FPC, FP, steps, xf, xt, ys);
    return cudaGetLastError() == cudaSuccess;
}

// force-load this file's kernels (CUDA loads functions lazily; loading one while another kernel spins
// on a cross-GPU flag could deadlock, so the extension loads everything up front)
void preload_train_tc2() {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, mlp1_train_tc2_kernel);
}

}  // namespace gb
