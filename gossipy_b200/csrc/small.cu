// Small-model kernels: logistic regression SGD, AdaLine / Pegasos sequential learners, online
// k-means, matrix-factorisation SGD.  These models are far too small for tensor cores
// (57 -> 2 etc.); each is one CTA (or one warp) with the model in shared memory / registers and
// the whole local update in ONE launch.  Reference: gossipy/model/handler.py:235-258 (with
// nn.py:147-174), :364-368, :416-423, :550-560, :604-615.
#include "common.cuh"
#include "ops.h"
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

namespace gb {

// ------------------------------------------------------------------------------------------------
// logistic regression: sigmoid(Wx+b) fed to mean cross-entropy, SGD(+wd); one CTA per model
// ------------------------------------------------------------------------------------------------
constexpr int LR_THREADS = 256;
constexpr int LR_BMAX = 64;     // samples processed per pass (larger batches loop over passes)
constexpr int LR_OMAX = 16;

struct LogregParams {
    float* row; const float* X; const int64_t* y; int n, IN, OUT, B, epochs; float lr, wd; uint64_t key;
    const int64_t* part_id; const int64_t* ages; int n_parts;
};

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
    for (int i = tid; i < P; i += LR_THREADS) W[i] = p.row[i];
    __syncthreads();
    const int n = p.n, B = p.B;
    const int spe = (n + B - 1) / B;
    const int total = p.epochs > 0 ? p.epochs * spe : 1;
    const bool scaled = p.part_id != nullptr;
    for (int s = 0; s < total; ++s) {
        const int e = p.epochs > 0 ? s / spe : 0;
        const int pos = p.epochs > 0 ? (s % spe) * B : 0;
        const int bcur = min(B, n - pos);
        GbPerm perm; perm.init((uint32_t)n, gb_mix64(p.key ^ (uint64_t)e));
        for (int i = tid; i < P; i += LR_THREADS) G[i] = 0.f;
        if (scaled && tid < p.n_parts) coef[tid] = 1.f / (float)(p.ages[tid] + (int64_t)s + 1);
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

int64_t logreg_train(at::Tensor row, at::Tensor X, at::Tensor y, std::tuple<int64_t, int64_t> dims,
                     int64_t batch_size, int64_t local_epochs, double lr, double wd, int64_t key,
                     c10::optional<at::Tensor> part_id, c10::optional<at::Tensor> ages) {
    TORCH_CHECK(row.is_cuda() && X.is_cuda() && y.is_cuda() && X.is_contiguous() && y.is_contiguous());
    TORCH_CHECK(X.scalar_type() == at::kFloat && y.scalar_type() == at::kLong && row.scalar_type() == at::kFloat);
    LogregParams p{};
    p.IN = (int)std::get<0>(dims); p.OUT = (int)std::get<1>(dims);
    p.n = (int)X.size(0);
    TORCH_CHECK(X.dim() == 2 && X.size(1) == p.IN && p.OUT <= LR_OMAX && p.n > 0);
    p.B = (int)(batch_size == 0 ? p.n : std::min<int64_t>(batch_size, p.n));
    p.epochs = (int)local_epochs; p.lr = (float)lr; p.wd = (float)wd; p.key = (uint64_t)key;
    p.row = row.data_ptr<float>(); p.X = X.data_ptr<float>(); p.y = y.data_ptr<int64_t>();
    if (part_id.has_value() && ages.has_value()) {
        p.part_id = part_id->data_ptr<int64_t>(); p.ages = ages->data_ptr<int64_t>();
        p.n_parts = (int)ages->numel();
        TORCH_CHECK(p.n_parts <= 16);
    }
    const int P = p.OUT * p.IN + p.OUT;
    const size_t smem = ((size_t)2 * P + (size_t)LR_BMAX * p.IN + LR_BMAX * LR_OMAX + 16) * 4;
    TORCH_CHECK(smem <= 200 * 1024, "logreg_train: model too large for the fused kernel");
    c10::cuda::CUDAGuard guard(row.device());
    C10_CUDA_CHECK(cudaFuncSetAttribute(logreg_train_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    logreg_train_kernel<<<1, LR_THREADS, smem, at::cuda::getCurrentCUDAStream()>>>(p);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    const int spe = (p.n + p.B - 1) / p.B;
    return p.epochs > 0 ? (int64_t)p.epochs * spe : 1;
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

at::Tensor logreg_scores(at::Tensor row, at::Tensor X, std::tuple<int64_t, int64_t> dims) {
    const int IN = (int)std::get<0>(dims), OUT = (int)std::get<1>(dims), n = (int)X.size(0);
    TORCH_CHECK(row.is_cuda() && X.is_cuda() && X.is_contiguous() && X.scalar_type() == at::kFloat);
    c10::cuda::CUDAGuard guard(row.device());
    auto out = at::empty({n, OUT}, X.options());
    const int blocks = std::max(1, std::min((n + 7) / 8, 148 * 8));
    logreg_scores_kernel<<<blocks, 256, 0, at::cuda::getCurrentCUDAStream()>>>(
        row.data_ptr<float>(), X.data_ptr<float>(), n, IN, OUT, out.data_ptr<float>());
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    return out;
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

void linear_seq_update(at::Tensor w, at::Tensor X, at::Tensor y, int64_t kind, double lr, int64_t n_updates) {
    TORCH_CHECK(w.is_cuda() && X.is_cuda() && y.is_cuda());
    auto Xc = X.to(at::kFloat).contiguous();
    auto yc = y.to(at::kFloat).contiguous();
    const int dim = (int)w.numel(), n = (int)Xc.size(0);
    TORCH_CHECK(dim <= 32 * SEQ_KPL && Xc.numel() == (int64_t)n * dim && yc.numel() == n);
    if (n == 0) return;
    c10::cuda::CUDAGuard guard(w.device());
    linear_seq_kernel<<<1, 32, 0, at::cuda::getCurrentCUDAStream()>>>(
        w.data_ptr<float>(), Xc.data_ptr<float>(), yc.data_ptr<float>(), n, dim, (int)kind, (float)lr,
        (long long)n_updates);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
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

at::Tensor kmeans_assign(at::Tensor C, at::Tensor X) {
    TORCH_CHECK(C.is_cuda() && X.is_cuda() && C.dim() == 2);
    auto Xc = X.to(at::kFloat).contiguous();
    auto Cc = C.contiguous();
    const int k = (int)C.size(0), dim = (int)C.size(1), n = (int)Xc.size(0);
    c10::cuda::CUDAGuard guard(C.device());
    auto out = at::empty({n}, Xc.options().dtype(at::kLong));
    if (n == 0) return out;
    const int blocks = std::max(1, std::min((n + 7) / 8, 148 * 8));
    kmeans_assign_kernel<<<blocks, 256, 0, at::cuda::getCurrentCUDAStream()>>>(
        Cc.data_ptr<float>(), Xc.data_ptr<float>(), n, k, dim, out.data_ptr<int64_t>());
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    return out;
}

void kmeans_update(at::Tensor C, at::Tensor X, double alpha) {
    TORCH_CHECK(C.is_cuda() && C.is_contiguous() && C.dim() == 2);
    auto Xc = X.to(at::kFloat).contiguous();
    const int k = (int)C.size(0), dim = (int)C.size(1), n = (int)Xc.size(0);
    if (n == 0) return;
    auto asg = kmeans_assign(C, Xc);
    c10::cuda::CUDAGuard guard(C.device());
    kmeans_apply_kernel<<<1, 256, k * sizeof(int), at::cuda::getCurrentCUDAStream()>>>(
        C.data_ptr<float>(), Xc.data_ptr<float>(), asg.data_ptr<int64_t>(), n, k, dim, (float)alpha);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
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

void mf_update(at::Tensor X, at::Tensor b, at::Tensor Y, at::Tensor c, at::Tensor ratings, double reg, double lr) {
    TORCH_CHECK(Y.is_cuda() && ratings.is_cuda() && Y.dim() == 2);
    auto rc = ratings.to(at::kFloat).contiguous();
    const int m = (int)rc.size(0), k = (int)Y.size(1);
    TORCH_CHECK(k <= 128, "mf_update: rank <= 128 supported");
    if (m == 0) return;
    c10::cuda::CUDAGuard guard(Y.device());
    mf_update_kernel<<<1, 32, 0, at::cuda::getCurrentCUDAStream()>>>(
        X.data_ptr<float>(), b.data_ptr<float>(), Y.data_ptr<float>(), c.data_ptr<float>(),
        rc.data_ptr<float>(), m, k, (float)reg, (float)lr);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

}  // namespace gb
