// Small-model kernels: logistic regression SGD, AdaLine / Pegasos sequential learners, online
// k-means, matrix-factorisation SGD.  These models are far too small for tensor cores
// (57 -> 2 etc.); each is one CTA (or one warp) with the model in shared memory / registers and
// the whole local update in ONE launch.  Reference: gossipy/model/handler.py:235-258 (with
// nn.py:147-174), :364-368, :416-423, :550-560, :604-615.
#include "common.cuh"
#include "kernels.h"
#include <algorithm>

namespace gb {

// ------------------------------------------------------------------------------------------------
// logistic regression: sigmoid(Wx+b) fed to mean cross-entropy, SGD(+wd); one CTA per model
// ------------------------------------------------------------------------------------------------
constexpr int LR_THREADS = 256;
constexpr int LR_BMAX = 64;     // samples processed per pass (larger batches loop over passes)
constexpr int LR_OMAX = 16;


__global__ void __launch_bounds__(LR_THREADS) logreg_train_kernel(const LogregParams p) {
    extern __shared__ __align__(16) float sm[];
    const int IN = p.IN, OUT = p.OUT, P = OUT * IN + OUT;
    float* W = sm;                       // [OUT][IN] then bias [OUT]
    float* G = W + P;                    // gradient accumulator [P]
    float* xs = G + P;                   // [LR_BMAX][IN]
    float* dz = xs + LR_BMAX * IN;       // [LR_BMAX][LR_OMAX]
    float* coef = dz + LR_BMAX * LR_OMAX;  // [16]
    __shared__ int ids[LR_BMAX];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = LR_THREADS / 32;
    if (p.peer != nullptr) {            // fused MERGE_UPDATE: start from w_self*row + w_peer*peer
        if (p.sync.ready != nullptr && tid == 0)
            gb_wait_flag(p.sync.ready, p.sync.gen, p.sync.fault);
        __syncthreads();
        for (int i = tid; i < P; i += LR_THREADS) W[i] = p.w_self * p.row[i] + p.w_peer * gb_ld_stream1(p.peer + i);
        __syncthreads();
        if (p.sync.done != nullptr && tid == 0) gb_red_release_sys_add(p.sync.done, 1u);
    } else {
        for (int i = tid; i < P; i += LR_THREADS) W[i] = p.row[i];
    }
    __syncthreads();
    const int n = p.n, B = p.B;
    const int spe = (n + B - 1) / B;
    const int total = p.epochs > 0 ? p.epochs * spe : 1;
    const bool scaled = p.scaled();
    for (int s = 0; s < total; ++s) {
        const int e = p.epochs > 0 ? s / spe : 0;
        const int pos = p.epochs > 0 ? (s % spe) * B : 0;
        const int bcur = min(B, n - pos);
        GbPerm perm; perm.init((uint32_t)n, gb_mix64(p.key ^ (uint64_t)e));
        for (int i = tid; i < P; i += LR_THREADS) G[i] = 0.f;
        if (scaled && tid < p.n_parts) coef[tid] = 1.f / (float)(p.age_of(tid) + (int64_t)s + 1);
        for (int c0 = 0; c0 < bcur; c0 += LR_BMAX) {           // passes over the mini-batch
            const int cb = min(LR_BMAX, bcur - c0);
            __syncthreads();
            if (tid < cb) ids[tid] = (int)perm((uint32_t)(pos + c0 + tid));
            __syncthreads();
            for (int i = tid; i < cb * IN; i += LR_THREADS) {
                const int b = i / IN, k = i - b * IN;
                xs[i] = p.X[(size_t)ids[b] * IN + k];
            }
            __syncthreads();
            for (int b = warp; b < cb; b += nwarps) {           // one warp per sample: forward
                float z[LR_OMAX];
                for (int o = 0; o < OUT; ++o) {
                    float acc = 0.f;
                    for (int k = lane; k < IN; k += 32) acc = fmaf(W[o * IN + k], xs[b * IN + k], acc);
                    z[o] = gb_warp_sum(acc) + W[OUT * IN + o];
                }
                if (lane == 0) {
                    float sg[LR_OMAX], m = -1e30f, sum = 0.f;
                    for (int o = 0; o < OUT; ++o) { sg[o] = 1.f / (1.f + __expf(-z[o])); m = fmaxf(m, sg[o]); }
                    float ex[LR_OMAX];
                    for (int o = 0; o < OUT; ++o) { ex[o] = __expf(sg[o] - m); sum += ex[o]; }
                    const int yy = (int)p.y[ids[b]];
                    for (int o = 0; o < OUT; ++o) {
                        const float pr = ex[o] / sum;
                        dz[b * LR_OMAX + o] = (pr - (o == yy ? 1.f : 0.f)) / (float)bcur * sg[o] * (1.f - sg[o]);
                    }
                }
            }
            __syncthreads();
            for (int i = tid; i < P; i += LR_THREADS) {         // gradient of every parameter
                float acc = 0.f;
                if (i < OUT * IN) {
                    const int o = i / IN, k = i - o * IN;
                    for (int b = 0; b < cb; ++b) acc = fmaf(dz[b * LR_OMAX + o], xs[b * IN + k], acc);
                } else {
                    const int o = i - OUT * IN;
                    for (int b = 0; b < cb; ++b) acc += dz[b * LR_OMAX + o];
                }
                G[i] += acc;
            }
        }
        __syncthreads();
        for (int i = tid; i < P; i += LR_THREADS) {
            float g = G[i];
            if (scaled) g *= coef[p.part_id[i]];
            W[i] = W[i] - p.lr * (g + p.wd * W[i]);
        }
        __syncthreads();
    }
    for (int i = tid; i < P; i += LR_THREADS) p.row[i] = W[i];
}

bool launch_logreg_train(LogregParams p, cudaStream_t stream) {
    if (p.OUT > LR_OMAX || p.n <= 0 || p.n_parts > 16) return false;
    const int P = p.OUT * p.IN + p.OUT;
    const size_t smem = ((size_t)2 * P + (size_t)LR_BMAX * p.IN + LR_BMAX * LR_OMAX + 16) * 4;
    if (smem > 200 * 1024) return false;
    static size_t configured = 0;
    if (smem > configured) {
        if (cudaFuncSetAttribute(logreg_train_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return false;
        configured = smem;
    }
    logreg_train_kernel<<<1, LR_THREADS, smem, stream>>>(p);
    return true;
}

__global__ void __launch_bounds__(256)
logreg_scores_kernel(const float* __restrict__ row, const float* __restrict__ X, int n, int IN, int OUT,
                     float* __restrict__ out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int s = warp; s < n; s += nw) {
        for (int o = 0; o < OUT; ++o) {
            float acc = 0.f;
            for (int k = lane; k < IN; k += 32) acc = fmaf(row[o * IN + k], X[(size_t)s * IN + k], acc);
            acc = gb_warp_sum(acc);
            if (lane == 0) out[(size_t)s * OUT + o] = 1.f / (1.f + __expf(-(acc + row[OUT * IN + o])));
        }
    }
}

void launch_logreg_scores(const float* row, const float* X, int n, int IN, int OUT, float* out,
                          cudaStream_t stream) {
    if (n <= 0) return;
    const int blocks = std::max(1, std::min((n + 7) / 8, sm_count() * 8));
    logreg_scores_kernel<<<blocks, 256, 0, stream>>>(row, X, n, IN, OUT, out);
}

// ------------------------------------------------------------------------------------------------
// AdaLine / Pegasos: strictly sequential per-sample updates; one warp, w in registers
// ------------------------------------------------------------------------------------------------
constexpr int SEQ_KPL = 32;   // supports dim <= 1024

__global__ void __launch_bounds__(32)
linear_seq_kernel(float* __restrict__ w, const float* __restrict__ X, const float* __restrict__ y,
                  int n, int dim, int kind, float lr, long long t0) {
    const int lane = threadIdx.x;
    float wr[SEQ_KPL];
#pragma unroll
    for (int i = 0; i < SEQ_KPL; ++i) { const int k = i * 32 + lane; wr[i] = k < dim ? w[k] : 0.f; }
    const int kpl = (dim + 31) / 32;
    long long t = t0;
    for (int s = 0; s < n; ++s) {
        const float* x = X + (size_t)s * dim;
        float xr[SEQ_KPL];
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < SEQ_KPL; ++i) {
            if (i < kpl) { const int k = i * 32 + lane; xr[i] = k < dim ? x[k] : 0.f; acc = fmaf(wr[i], xr[i], acc); }
        }
        const float yhat = gb_warp_sum(acc);
        const float ys = y[s];
        if (kind == 0) {                       // AdaLine: w += lr (y - w.x) x
            const float c = lr * (ys - yhat);
#pragma unroll
            for (int i = 0; i < SEQ_KPL; ++i) if (i < kpl) wr[i] = fmaf(c, xr[i], wr[i]);
        } else {                               // Pegasos: t=++age; eta=1/(t lam); w*=(1-eta lam); hinge step
            t += 1;
            const float eta = 1.f / ((float)t * lr);
            const float sc = 1.f - eta * lr;
            const float c = (yhat * ys - 1.f < 0.f) ? eta * ys : 0.f;
#pragma unroll
            for (int i = 0; i < SEQ_KPL; ++i) if (i < kpl) wr[i] = fmaf(c, xr[i], wr[i] * sc);
        }
    }
#pragma unroll
    for (int i = 0; i < SEQ_KPL; ++i) { const int k = i * 32 + lane; if (k < dim) w[k] = wr[i]; }
}

void launch_linear_seq(float* w, const float* X, const float* y, int n, int dim, int kind, float lr,
                       long long t0, cudaStream_t stream) {
    if (n <= 0) return;
    linear_seq_kernel<<<1, 32, 0, stream>>>(w, X, y, n, dim, kind, lr, t0);
}

// ------------------------------------------------------------------------------------------------
// k-means: assignment (+ the reference's batched EMA update, last sample per centroid wins)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
kmeans_assign_kernel(const float* __restrict__ C, const float* __restrict__ X, int n, int k, int dim,
                     int64_t* __restrict__ out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int s = warp; s < n; s += nw) {
        int best = 0; float bd = 3.4e38f;
        for (int c = 0; c < k; ++c) {
            float acc = 0.f;
            for (int d = lane; d < dim; d += 32) { const float df = X[(size_t)s * dim + d] - C[c * dim + d]; acc = fmaf(df, df, acc); }
            acc = gb_warp_sum(acc);
            if (acc < bd) { bd = acc; best = c; }
        }
        if (lane == 0) out[s] = best;
    }
}

// one CTA: winner[c] = largest sample index assigned to c (matches the "last write wins" of
// `C[idx] = C[idx]*(1-a) + a*x` in the reference for sorted duplicate handling by index order)
__global__ void __launch_bounds__(256)
kmeans_apply_kernel(float* __restrict__ C, const float* __restrict__ X, const int64_t* __restrict__ asg,
                    int n, int k, int dim, float alpha) {
    extern __shared__ int winner[];
    for (int c = threadIdx.x; c < k; c += blockDim.x) winner[c] = -1;
    __syncthreads();
    for (int s = threadIdx.x; s < n; s += blockDim.x) atomicMax(&winner[(int)asg[s]], s);
    __syncthreads();
    for (int e = threadIdx.x; e < k * dim; e += blockDim.x) {
        const int c = e / dim, d = e - c * dim;
        const int s = winner[c];
        if (s >= 0) C[e] = C[e] * (1.f - alpha) + alpha * X[(size_t)s * dim + d];
    }
}

// Matched merge of two sets of k <= 8 centroids (KMeansHandler "hungarian", reference handler.py:617-630 with the
// column assignment actually applied): the optimal assignment minimising the summed Euclidean distance
// is found by enumerating all k! <= 40 320 permutations (factoradic decode, 256 threads, packed atomicMin -- ties go
// to the lexicographically smallest permutation), then C = w_own * C + w_peer * C_peer[perm] in the same kernel.
// No host round trip (the reference's path is cdist -> CPU -> scipy.linear_sum_assignment -> GPU).
__global__ void __launch_bounds__(256)
kmeans_match_merge_kernel(float* __restrict__ C, const float* __restrict__ P, int k, int dim, float w_own, float w_peer,
                          PeerSync sync, int64_t* __restrict__ perm_out) {
    __shared__ float cost[64];
    __shared__ unsigned long long best;
    __shared__ int perm[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (sync.ready != nullptr) { if (tid == 0) gb_wait_flag(sync.ready, sync.gen, sync.fault); __syncthreads(); }
    if (tid == 0) best = ~0ull;
    for (int e = warp; e < k * k; e += 8) {                 // cost[i][j] = || C[i] - P[j] ||
        const int i = e / k, j = e % k;
        float acc = 0.f;
        for (int d = lane; d < dim; d += 32) { const float df = C[i * dim + d] - gb_ld_stream1(P + j * dim + d); acc = fmaf(df, df, acc); }
        acc = gb_warp_sum(acc);
        if (lane == 0) cost[i * 8 + j] = sqrtf(acc);
    }
    __syncthreads();
    int nperm = 1;
    for (int i = 2; i <= k; ++i) nperm *= i;
    for (int pi = tid; pi < nperm; pi += 256) {
        int avail[8], idx = pi, fact = nperm;
#pragma unroll
        for (int i = 0; i < 8; ++i) avail[i] = i;
        float total = 0.f;
        for (int i = 0; i < k; ++i) {                        // factoradic digits, most significant first: lexicographic order
            fact /= (k - i);
            const int dgt = idx / fact; idx -= dgt * fact;
            int col = avail[dgt];
            for (int t = dgt; t < k - i - 1; ++t) avail[t] = avail[t + 1];
            total += cost[i * 8 + col];
        }
        atomicMin(&best, ((unsigned long long)__float_as_uint(total) << 32) | (unsigned)pi);     // costs >= 0: bit order = value order
    }
    __syncthreads();
    if (tid == 0) {
        int avail[8], idx = (int)(best & 0xffffffffull), fact = nperm;
        for (int i = 0; i < 8; ++i) avail[i] = i;
        for (int i = 0; i < k; ++i) {
            fact /= (k - i);
            const int dgt = idx / fact; idx -= dgt * fact;
            perm[i] = avail[dgt];
            for (int t = dgt; t < k - i - 1; ++t) avail[t] = avail[t + 1];
            if (perm_out != nullptr) perm_out[i] = perm[i];
        }
    }
    __syncthreads();
    for (int e = tid; e < k * dim; e += 256) {
        const int i = e / dim, d = e - i * dim;
        C[e] = w_own * C[e] + w_peer * gb_ld_stream1(P + perm[i] * dim + d);
    }
    if (sync.done != nullptr) { __syncthreads(); if (tid == 0) { __threadfence(); gb_red_release_sys_add(sync.done, 1u); } }
}

bool launch_kmeans_match_merge(float* C, const float* P, int k, int dim, float w_own, float w_peer, PeerSync sync,
                               int64_t* perm_out, cudaStream_t stream) {
    if (k < 1 || k > 8 || dim < 1) return false;
    kmeans_match_merge_kernel<<<1, 256, 0, stream>>>(C, P, k, dim, w_own, w_peer, sync, perm_out);
    return true;
}

void launch_kmeans_assign(const float* C, const float* X, int n, int k, int dim, int64_t* out,
                          cudaStream_t stream) {
    if (n <= 0) return;
    const int blocks = std::max(1, std::min((n + 7) / 8, sm_count() * 8));
    kmeans_assign_kernel<<<blocks, 256, 0, stream>>>(C, X, n, k, dim, out);
}

void launch_kmeans_apply(float* C, const float* X, const int64_t* asg, int n, int k, int dim, float alpha,
                         cudaStream_t stream) {
    if (n <= 0) return;
    kmeans_apply_kernel<<<1, 256, k * sizeof(int), stream>>>(C, X, asg, n, k, dim, alpha);
}

// ------------------------------------------------------------------------------------------------
// matrix factorisation: sequential rank-k SGD over one user's ratings; one warp, lanes = factors
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
mf_update_kernel(float* __restrict__ Xu, float* __restrict__ bu, float* __restrict__ Y, float* __restrict__ c,
                 const float* __restrict__ ratings, int m, int k, float reg, float lr) {
    const int lane = threadIdx.x;
    const int kpl = (k + 31) / 32;          // k <= 128
    float x[4], b = bu[0];
    for (int i = 0; i < 4; ++i) { const int f = i * 32 + lane; x[i] = (i < kpl && f < k) ? Xu[f] : 0.f; }
    const float decay = 1.f - reg * lr;
    for (int r = 0; r < m; ++r) {
        const int item = (int)ratings[2 * r];
        const float rating = ratings[2 * r + 1];
        float yv[4], acc = 0.f;
        for (int i = 0; i < 4; ++i) {
            const int f = i * 32 + lane;
            yv[i] = (i < kpl && f < k) ? Y[(size_t)item * k + f] : 0.f;
            acc = fmaf(x[i], yv[i], acc);
        }
        const float ci = c[item];
        const float err = rating - gb_warp_sum(acc) - b - ci;
        for (int i = 0; i < 4; ++i) {
            const int f = i * 32 + lane;
            if (i < kpl && f < k) {
                const float ynew = decay * yv[i] + lr * err * x[i];     // item factor first ...
                x[i] = decay * x[i] + lr * err * ynew;                  // ... user factor uses the NEW one
                Y[(size_t)item * k + f] = ynew;
            }
        }
        b += lr * err;
        if (lane == 0) c[item] = ci + lr * err;
        __syncwarp();
    }
    for (int i = 0; i < 4; ++i) { const int f = i * 32 + lane; if (i < kpl && f < k) Xu[f] = x[i]; }
    if (lane == 0) bu[0] = b;
}

void launch_mf_update(float* Xu, float* bu, float* Y, float* c, const float* ratings, int m, int k,
                      float reg, float lr, cudaStream_t stream) {
    if (m <= 0) return;
    mf_update_kernel<<<1, 32, 0, stream>>>(Xu, bu, Y, c, ratings, m, k, reg, lr);
}

// force-load this file's kernels (CUDA loads functions lazily; loading one while another kernel spins
// on a cross-GPU flag could deadlock, so the extension loads everything up front)
void preload_small() {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, logreg_train_kernel);
    cudaFuncGetAttributes(&a, logreg_scores_kernel);
    cudaFuncGetAttributes(&a, linear_seq_kernel);
    cudaFuncGetAttributes(&a, kmeans_assign_kernel);
    cudaFuncGetAttributes(&a, kmeans_apply_kernel);
    cudaFuncGetAttributes(&a, kmeans_match_merge_kernel);
    cudaFuncGetAttributes(&a, mf_update_kernel);
}

}  // namespace gb
