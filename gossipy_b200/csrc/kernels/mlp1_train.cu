// Fused local update of a Linear-ReLU-Linear network: a WHOLE `_update` (all local epochs: keyed
// shuffle, forward, softmax-CE gradient, backward, SGD with weight decay) is ONE kernel launch.
// Reference semantics: gossipy/model/handler.py:235-258 (+ 503-520 for PartitionedTMH).
//
// "cluster" implementation (exact fp32, CUDA cores):
//   * one thread-block cluster of C CTAs (C <= 8) per model; CTA r owns hidden units
//     [r*Hs, (r+1)*Hs) -- their W1 rows live in REGISTERS for the entire launch (lane l of the
//     warp that owns a unit holds columns l, l+32, ...), their W2 columns / biases in shared memory;
//   * the mini-batch is streamed with cp.async into a double-buffered shared-memory tile (the next
//     batch lands while the current one is being processed); sample order comes from the keyed
//     Feistel permutation, so no shuffled copy of the shard is ever materialised;
//   * per step the only cross-CTA traffic is the [B x OUT] partial-logit tile, exchanged through
//     distributed shared memory (st.shared::cluster) followed by ONE cluster barrier;
//   * weights are written back to the HBM row once, at the end.
// The tcgen05 / TMEM implementation of the same op lives in mlp1_train_tc.cu.
#include "common.cuh"
#include "kernels.h"
#include <algorithm>

namespace gb {

constexpr int UPW = 4;        // hidden units per warp
constexpr int NW = 4;         // warps per CTA
constexpr int SLOTS = UPW * NW;  // hidden-unit slots per CTA
constexpr int OUT_MAX = 16;
constexpr int CMAX = 8;       // portable cluster size
constexpr int MAX_PARTS = 16;


template <int KPL, bool SCALED>
__global__ void __launch_bounds__(NW * 32, 1) mlp1_train_cluster_kernel(const TrainParams p) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t rank = gb_cluster_ctarank();
    const int IN = p.IN, H = p.H, OUT = p.OUT, B = p.B;
    const int INP = KPL * 32;
    const int BP = (B + 7) & ~7;
    const int unit0 = rank * p.Hs;                               // first hidden unit of this CTA
    const int nslots = max(0, min(p.Hs, H - unit0));             // valid slots in this CTA

    // ---- shared memory carve-up -------------------------------------------------------------
    float* xs = smem;                                   // [nbuf][BP][INP]
    float* hs = xs + (size_t)p.nbuf * BP * INP;          // [BP][SLOTS]   relu(z1)
    float* dz1s = hs + BP * SLOTS;                       // [BP][SLOTS]   -lr * dL/dz1 (or -dL/dz1)
    float* w2s = dz1s + BP * SLOTS;                      // [OUT_MAX][SLOTS]
    float* b1s = w2s + OUT_MAX * SLOTS;                  // [SLOTS]
    float* b2s = b1s + SLOTS;                            // [OUT_MAX]
    float* z2s = b2s + OUT_MAX;                          // [BP][OUT_MAX]  logits, then dL/dz2
    float* part = z2s + BP * OUT_MAX;                    // [2][CMAX][BP*OUT_MAX] partial logits
    float* coef = part + 2 * CMAX * BP * OUT_MAX;        // [MAX_PARTS] lr/age per partition
    int* idxs = reinterpret_cast<int*>(coef + MAX_PARTS);  // [nbuf][BP] sample ids
    int* ys = idxs + p.nbuf * BP;                        // [nbuf][BP] labels

    float* b1g = p.row + (size_t)H * IN;
    float* W2g = b1g + H;
    float* b2g = W2g + (size_t)OUT * H;

    // ---- load this CTA's parameters ----------------------------------------------------------
    // fused MERGE_UPDATE: with a peer row the starting point is w_self*row + w_peer*peer, the peer
    // row being pulled (possibly over NVLink) while the weights are loaded on chip
    const bool merging = p.peer != nullptr;
    if (merging && p.sync.ready != nullptr) {
        if (tid == 0) gb_wait_flag(p.sync.ready, p.sync.gen, p.sync.fault);
        __syncthreads();
    }
    auto ldp = [&](size_t off) -> float {
        const float own = p.row[off];
        return merging ? p.w_self * own + p.w_peer * gb_ld_stream1(p.peer + off) : own;
    };
    const size_t off_b1 = (size_t)H * IN, off_w2 = off_b1 + H, off_b2 = off_w2 + (size_t)OUT * H;
    float w[UPW][KPL];
    uint32_t pidpack[SCALED ? (UPW * KPL + 3) / 4 : 1];
    if (SCALED) {
#pragma unroll
        for (int q = 0; q < (UPW * KPL + 3) / 4; ++q) pidpack[q] = 0;
    }
#pragma unroll
    for (int u = 0; u < UPW; ++u) {
        const int slot = warp * UPW + u;
        const bool valid = slot < nslots;
#pragma unroll
        for (int i = 0; i < KPL; ++i) {
            const int k = i * 32 + lane;
            const bool ok = valid && k < IN;
            w[u][i] = ok ? ldp((size_t)(unit0 + slot) * IN + k) : 0.f;
            if (SCALED) {
                const uint32_t pid = ok ? (uint32_t)p.part_id[(size_t)(unit0 + slot) * IN + k] : 0u;
                const int e = u * KPL + i;
                pidpack[e >> 2] |= (pid & 0xffu) << (8 * (e & 3));
            }
        }
    }
    for (int i = tid; i < p.nbuf * BP * INP; i += blockDim.x) xs[i] = 0.f;   // zero incl. padding
    for (int i = tid; i < OUT_MAX * SLOTS; i += blockDim.x) {
        const int o = i / SLOTS, s = i % SLOTS;
        w2s[i] = (o < OUT && s < nslots) ? ldp(off_w2 + (size_t)o * H + unit0 + s) : 0.f;
    }
    if (tid < SLOTS) b1s[tid] = (tid < nslots) ? ldp(off_b1 + unit0 + tid) : 0.f;
    if (tid < OUT_MAX) b2s[tid] = (tid < OUT) ? ldp(off_b2 + tid) : 0.f;
    for (int i = tid; i < BP * SLOTS; i += blockDim.x) { hs[i] = 0.f; dz1s[i] = 0.f; }
    __syncthreads();
    gb_cluster_sync();   // every CTA of the cluster is running before anyone writes into its smem
    if (merging && p.sync.done != nullptr && rank == 0 && tid == 0)
        gb_red_release_sys_add(p.sync.done, 1u);   // all CTAs have consumed their peer loads

    const int n = p.n;
    const int spe = (n + B - 1) / B;                          // steps per epoch
    const int total_steps = p.epochs > 0 ? p.epochs * spe : 1;
    const int vec_per_row = IN >> 2;                          // IN % 4 == 0 (checked on the host)

    auto stage = [&](int s, int buf) {                        // issue the loads of step s
        const int e = p.epochs > 0 ? s / spe : 0;
        const int pos = p.epochs > 0 ? (s % spe) * B : 0;
        const int bcur = min(B, n - pos);
        if (tid < bcur) {
            GbPerm perm; perm.init((uint32_t)n, gb_mix64(p.key ^ (uint64_t)e));
            const int id = (int)perm((uint32_t)(pos + tid));
            idxs[buf * BP + tid] = id;
            ys[buf * BP + tid] = (int)p.y[id];
        }
        __syncthreads();
        float* dstb = xs + (size_t)buf * BP * INP;
        for (int c = tid; c < bcur * vec_per_row; c += blockDim.x) {
            const int b = c / vec_per_row, v = c - b * vec_per_row;
            gb_cp_async16(dstb + (size_t)b * INP + 4 * v,
                          p.X + (size_t)idxs[buf * BP + b] * IN + 4 * v);
        }
        gb_cp_async_commit();
    };

    stage(0, 0);
    for (int s = 0; s < total_steps; ++s) {
        const int buf = (p.nbuf == 2) ? (s & 1) : 0;
        const int pos = p.epochs > 0 ? (s % spe) * B : 0;
        const int bcur = min(B, n - pos);
        if (p.nbuf == 2) {
            if (s + 1 < total_steps) { stage(s + 1, buf ^ 1); gb_cp_async_wait<1>(); }
            else gb_cp_async_wait<0>();
        } else {
            if (s > 0) stage(s, 0);
            gb_cp_async_wait<0>();
        }
        __syncthreads();
        const float* xb = xs + (size_t)buf * BP * INP;
        const int* yb = ys + buf * BP;
        const float inv_b = 1.f / (float)bcur;

        if (SCALED && tid < p.n_parts)     // ages are incremented before the step (ref :506)
            coef[tid] = p.lr / (float)(p.age_of(tid) + (int64_t)s + 1);

        // ---- forward, layer 1: z1 = x W1^T + b1 ; h = relu(z1) -------------------------------
        if (warp * UPW < nslots) {
            for (int g = 0; g < BP; g += 8) {
                if (g >= bcur) break;
                float val[32];
#pragma unroll
                for (int v = 0; v < 32; ++v) val[v] = 0.f;
#pragma unroll
                for (int i = 0; i < KPL; ++i) {
                    float xv[8];
#pragma unroll
                    for (int sI = 0; sI < 8; ++sI) xv[sI] = xb[(size_t)(g + sI) * INP + i * 32 + lane];
#pragma unroll
                    for (int sI = 0; sI < 8; ++sI)
#pragma unroll
                        for (int u = 0; u < UPW; ++u) val[sI * UPW + u] = fmaf(w[u][i], xv[sI], val[sI * UPW + u]);
                }
                // transposing butterfly: 32 values x 32 lanes -> lane L holds the full sum of value L
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    const bool upper = (lane & off) != 0;
#pragma unroll
                    for (int j = 0; j < off; ++j) {
                        const float keep = upper ? val[j + off] : val[j];
                        const float send = upper ? val[j] : val[j + off];
                        val[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                    }
                }
                const int sI = lane / UPW, u = lane % UPW, slot = warp * UPW + u, b = g + sI;
                const float z = val[0] + b1s[slot];
                hs[b * SLOTS + slot] = (slot < nslots && b < bcur) ? fmaxf(z, 0.f) : 0.f;
            }
        }
        __syncthreads();

        // ---- forward, layer 2: partial logits over my hidden slice -> every CTA of the cluster --
        const int par = s & 1;
        for (int e = tid; e < bcur * OUT; e += blockDim.x) {
            const int b = e / OUT, o = e - b * OUT;
            float acc = 0.f;
#pragma unroll
            for (int sl = 0; sl < SLOTS; ++sl) acc = fmaf(hs[b * SLOTS + sl], w2s[o * SLOTS + sl], acc);
            float* mine = part + ((size_t)(par * CMAX + rank) * BP + b) * OUT_MAX + o;
            for (int r = 0; r < p.C; ++r) gb_st_cluster(gb_map_shared(mine, r), acc);
        }
        gb_cluster_sync();   // release my stores / acquire everybody else's

        // ---- softmax cross-entropy gradient (every CTA computes the full [B x OUT] tile) ---------
        for (int e = tid; e < bcur * OUT; e += blockDim.x) {
            const int b = e / OUT, o = e - b * OUT;
            float z = b2s[o];
            for (int r = 0; r < p.C; ++r) z += part[((size_t)(par * CMAX + r) * BP + b) * OUT_MAX + o];
            z2s[b * OUT_MAX + o] = z;
        }
        __syncthreads();
        if (tid < bcur) {
            float* zr = z2s + tid * OUT_MAX;
            float m = zr[0];
            for (int o = 1; o < OUT; ++o) m = fmaxf(m, zr[o]);
            float sum = 0.f;
            for (int o = 0; o < OUT; ++o) { const float ex = __expf(zr[o] - m); zr[o] = ex; sum += ex; }
            const float inv = 1.f / sum;
            const int yy = yb[tid];
            for (int o = 0; o < OUT; ++o) zr[o] = (zr[o] * inv - (o == yy ? 1.f : 0.f)) * inv_b;
        }
        __syncthreads();

        // ---- backward through layer 2 ---------------------------------------------------------------
        const float gscale = SCALED ? -1.f : -p.lr;      // dz1s carries -lr (plain) or -1 (scaled)
        for (int e = tid; e < bcur * SLOTS; e += blockDim.x) {
            const int b = e / SLOTS, sl = e - b * SLOTS;
            float dh = 0.f;
            for (int o = 0; o < OUT; ++o) dh = fmaf(z2s[b * OUT_MAX + o], w2s[o * SLOTS + sl], dh);
            dz1s[e] = (hs[e] > 0.f) ? gscale * dh : 0.f;
        }
        __syncthreads();
        const float decay = 1.f - p.lr * p.wd;
        for (int e = tid; e < OUT * SLOTS; e += blockDim.x) {          // W2 slice
            const int o = e / SLOTS, sl = e - o * SLOTS;
            if (sl < nslots) {
                float gacc = 0.f;
                for (int b = 0; b < bcur; ++b) gacc = fmaf(z2s[b * OUT_MAX + o], hs[b * SLOTS + sl], gacc);
                float c = p.lr;
                if (SCALED) c = coef[p.part_id[(size_t)H * IN + H + (size_t)o * H + unit0 + sl]];
                w2s[o * SLOTS + sl] = fmaf(-c, gacc, w2s[o * SLOTS + sl] * decay);
            }
        }
        if (tid < OUT) {                                                // b2 (replicated in every CTA)
            float gacc = 0.f;
            for (int b = 0; b < bcur; ++b) gacc += z2s[b * OUT_MAX + tid];
            float c = p.lr;
            if (SCALED) c = coef[p.part_id[(size_t)H * IN + H + (size_t)OUT * H + tid]];
            b2s[tid] = fmaf(-c, gacc, b2s[tid] * decay);
        }
        if (tid >= 32 && tid < 32 + nslots) {                           // b1 slice
            const int sl = tid - 32;
            float gacc = 0.f;
            for (int b = 0; b < bcur; ++b) gacc += dz1s[b * SLOTS + sl];   // already carries gscale
            float c = 1.f;
            if (SCALED) c = coef[p.part_id[(size_t)H * IN + unit0 + sl]];
            b1s[sl] = fmaf(c, gacc, b1s[sl] * decay);
        }

        // ---- backward into W1 (registers): w = w*decay + sum_b (-lr dz1[b]) x[b] ---------------------
        if (warp * UPW < nslots) {
            if (SCALED) {
#pragma unroll
                for (int u = 0; u < UPW; ++u)
#pragma unroll
                    for (int i = 0; i < KPL; ++i) {
                        const int e = u * KPL + i;
                        const float c = coef[(pidpack[e >> 2] >> (8 * (e & 3))) & 0xffu];
                        w[u][i] *= decay / c;       // w_new = c * (w*decay/c + sum -dz1 x)
                    }
            } else if (p.wd != 0.f) {
#pragma unroll
                for (int u = 0; u < UPW; ++u)
#pragma unroll
                    for (int i = 0; i < KPL; ++i) w[u][i] *= decay;
            }
            for (int b = 0; b < bcur; ++b) {
                const float4 d = *reinterpret_cast<const float4*>(dz1s + b * SLOTS + warp * UPW);
                const float* xr = xb + (size_t)b * INP + lane;
#pragma unroll
                for (int i = 0; i < KPL; ++i) {
                    const float xv = xr[i * 32];
                    w[0][i] = fmaf(d.x, xv, w[0][i]); w[1][i] = fmaf(d.y, xv, w[1][i]);
                    w[2][i] = fmaf(d.z, xv, w[2][i]); w[3][i] = fmaf(d.w, xv, w[3][i]);
                }
            }
            if (SCALED) {
#pragma unroll
                for (int u = 0; u < UPW; ++u)
#pragma unroll
                    for (int i = 0; i < KPL; ++i) {
                        const int e = u * KPL + i;
                        w[u][i] *= coef[(pidpack[e >> 2] >> (8 * (e & 3))) & 0xffu];
                    }
            }
        }
        __syncthreads();
    }

    // ---- write the parameters back to the HBM row --------------------------------------------------
#pragma unroll
    for (int u = 0; u < UPW; ++u) {
        const int slot = warp * UPW + u;
        if (slot < nslots) {
#pragma unroll
            for (int i = 0; i < KPL; ++i) {
                const int k = i * 32 + lane;
                if (k < IN) p.row[(size_t)(unit0 + slot) * IN + k] = w[u][i];
            }
        }
    }
    for (int i = tid; i < OUT * SLOTS; i += blockDim.x) {
        const int o = i / SLOTS, sl = i % SLOTS;
        if (sl < nslots) W2g[(size_t)o * H + unit0 + sl] = w2s[i];
    }
    if (tid < nslots) b1g[unit0 + tid] = b1s[tid];
    if (rank == 0 && tid < OUT) b2g[tid] = b2s[tid];
    gb_cluster_sync();   // no CTA may exit while peers can still write into its shared memory
}

static size_t train_smem_bytes(int nbuf, int BP, int INP) {
    size_t fl = (size_t)nbuf * BP * INP + 2 * (size_t)BP * SLOTS + OUT_MAX * SLOTS + SLOTS + OUT_MAX +
                (size_t)BP * OUT_MAX + 2 * (size_t)CMAX * BP * OUT_MAX + MAX_PARTS;
    return fl * 4 + (size_t)2 * nbuf * BP * 4;
}

template <int KPL>
static bool launch_cluster(const TrainParams& p, bool scaled, cudaStream_t stream) {
    const int BP = (p.B + 7) & ~7, INP = KPL * 32;
    TrainParams q = p;
    q.nbuf = 2;
    size_t smem = train_smem_bytes(2, BP, INP);
    if (smem > 227 * 1024) { q.nbuf = 1; smem = train_smem_bytes(1, BP, INP); }
    if (smem > 227 * 1024) return false;
    auto kern = scaled ? mlp1_train_cluster_kernel<KPL, true> : mlp1_train_cluster_kernel<KPL, false>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(q.C); cfg.blockDim = dim3(NW * 32); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = q.C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, q) == cudaSuccess;
}

// auto = fp32-equivalent: tc4 (4-CTA cluster, 3xTF32 tcgen05) -> cluster kernel (fp32 CUDA cores).  The plain-tf32
// variants (tc3, tc8-tf32) run only when asked for by name (GlobalSettings().allow_tf32 / impl=...).
static TrainImpl g_auto_impl = kTrainAuto;     // what "auto" means for this process (ops.set_train_impl / allow_tf32)
void set_default_train_impl(TrainImpl impl) { g_auto_impl = impl; }

bool launch_mlp1_train(TrainParams p, TrainImpl impl, cudaStream_t stream, const char** why) {
    static const char* kNone = "";
    *why = kNone;
    if (impl == kTrainAuto) impl = g_auto_impl;
    const bool scaled = p.scaled();
    if (scaled && p.n_parts > MAX_PARTS) { *why = "fused partitioned training supports <= 16 partitions"; return false; }
    if (p.stage_mode == 1) {        // loader only: meaningful for the tcgen05 (tc8) kernels; false = nothing was staged
        const bool tc8 = impl == kTrainAuto || impl == kTrainTc8 || (impl == kTrainTc8Tf32 && !scaled && p.momentum == 0.f);
        return tc8 && mlp1_train_tc4(p, 8, impl != kTrainTc8Tf32, stream);
    }
    if (p.momentum != 0.f) {                                       // fused momentum-SGD: the tcgen05 kernel only
        if ((impl == kTrainAuto || impl == kTrainTc8) && mlp1_train_tc4(p, 8, true, stream)) return true;
        *why = "fused momentum-SGD needs the tcgen05 (tc8) kernel: 32 <= in <= 896 (multiple of 4), hidden <= 128, out <= 10, batch <= 32";
        return false;
    }
    if (scaled && (impl == kTrainAuto || impl == kTrainTc8)) {     // K3 on the tensor-core kernel (<= 16 partitions)
        if (mlp1_train_tc4(p, 8, true, stream)) return true;
        if (impl == kTrainTc8) { *why = "tcgen05 (tc8) training kernel does not support this partitioned configuration"; return false; }
    }
    if (impl != kTrainCluster && !scaled) {
        switch (impl) {
        case kTrainTc8: case kTrainTc8Tf32:
            if (mlp1_train_tc4(p, 8, impl == kTrainTc8, stream)) return true;
            *why = "tcgen05 (tc4) training kernel does not support this configuration";
            return false;
        case kTrainTc3:
            if (mlp1_train_tc3(p, stream)) return true;
            *why = "tcgen05 (tc3) training kernel does not support this configuration";
            return false;
        default:
            if (mlp1_train_tc4(p, 8, true, stream)) return true;
            break;
        }
    }
    if (!(p.IN % 4 == 0 && p.IN <= 1024 && p.OUT <= OUT_MAX && p.B <= 64 && p.H <= CMAX * SLOTS)) {
        *why = "mlp1_train(cluster): unsupported shape";
        return false;
    }
    p.C = std::min(CMAX, std::max(1, (p.H + UPW - 1) / UPW));
    p.Hs = (p.H + p.C - 1) / p.C;
    if (p.Hs > SLOTS) { *why = "mlp1_train(cluster): hidden slice too large"; return false; }
    const int kpl = (p.IN + 31) / 32;
    bool ok;
    if (kpl <= 2) ok = launch_cluster<2>(p, scaled, stream);
    else if (kpl <= 8) ok = launch_cluster<8>(p, scaled, stream);
    else if (kpl <= 16) ok = launch_cluster<16>(p, scaled, stream);
    else if (kpl <= 25) ok = launch_cluster<25>(p, scaled, stream);
    else ok = launch_cluster<32>(p, scaled, stream);
    if (!ok) *why = "mlp1_train(cluster): batch tile does not fit in shared memory / launch failed";
    return ok;
}

// force-load this file's kernels (CUDA loads functions lazily; loading one while another kernel spins
// on a cross-GPU flag could deadlock, so the extension loads everything up front)
void preload_train_cluster() {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, mlp1_train_cluster_kernel<2, false>);
    cudaFuncGetAttributes(&a, mlp1_train_cluster_kernel<2, true>);
    cudaFuncGetAttributes(&a, mlp1_train_cluster_kernel<8, false>);
    cudaFuncGetAttributes(&a, mlp1_train_cluster_kernel<8, true>);
    cudaFuncGetAttributes(&a, mlp1_train_cluster_kernel<16, false>);
    cudaFuncGetAttributes(&a, mlp1_train_cluster_kernel<16, true>);
    cudaFuncGetAttributes(&a, mlp1_train_cluster_kernel<25, false>);
    cudaFuncGetAttributes(&a, mlp1_train_cluster_kernel<25, true>);
    cudaFuncGetAttributes(&a, mlp1_train_cluster_kernel<32, false>);
    cudaFuncGetAttributes(&a, mlp1_train_cluster_kernel<32, true>);
}

}  // namespace gb
