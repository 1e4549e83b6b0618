// Evaluation of a Linear-ReLU-Linear network on a (test) set, reduced ON THE DEVICE to a
// confusion matrix: only C*C integers ever reach the host (the reference ships all predictions to
// the CPU and calls scikit-learn, gossipy/model/handler.py:282-334).
//
// "simt" implementation: fp32 CUDA-core tile kernel (exact).  The tcgen05 implementation
// (bf16 operands, fp32 accumulation in TMEM, TMA-fed) is in mlp1_eval_tc.cu and is preferred
// whenever a bf16 copy of the test set is supplied.
#include "common.cuh"
#include "kernels.h"

namespace gb {

constexpr int EV_TS = 32;      // samples per CTA
constexpr int EV_KC = 32;      // K chunk
constexpr int EV_HMAX = 128;   // hidden units handled by one pass
constexpr int EV_THREADS = 256;

__global__ void __launch_bounds__(EV_THREADS)
mlp1_eval_simt_kernel(const float* __restrict__ row, const float* __restrict__ X,
                      const int64_t* __restrict__ y, int n, int IN, int H, int OUT, int n_classes,
                      int* __restrict__ cm, float* __restrict__ score1) {
    __shared__ float xs[EV_TS][EV_KC + 1];
    __shared__ float ws[EV_HMAX][EV_KC + 1];
    __shared__ float hs[EV_TS][EV_HMAX + 1];
    __shared__ float zs[EV_TS][17];
    const int tid = threadIdx.x;
    const int s0 = blockIdx.x * EV_TS;
    const float* W1 = row;
    const float* b1 = row + (size_t)H * IN;
    const float* W2 = b1 + H;
    const float* b2 = W2 + (size_t)OUT * H;
    const int j = tid % EV_HMAX;             // hidden unit of this thread
    const int sh = tid / EV_HMAX;            // which half of the sample tile
    constexpr int SPT = EV_TS / (EV_THREADS / EV_HMAX);   // samples per thread (16)
    float acc[SPT];
#pragma unroll
    for (int i = 0; i < SPT; ++i) acc[i] = 0.f;
    for (int k0 = 0; k0 < IN; k0 += EV_KC) {
        for (int e = tid; e < EV_TS * EV_KC; e += EV_THREADS) {
            const int r = e / EV_KC, c = e % EV_KC;
            xs[r][c] = (s0 + r < n && k0 + c < IN) ? X[(size_t)(s0 + r) * IN + k0 + c] : 0.f;
        }
        for (int e = tid; e < EV_HMAX * EV_KC; e += EV_THREADS) {
            const int r = e / EV_KC, c = e % EV_KC;
            ws[r][c] = (r < H && k0 + c < IN) ? W1[(size_t)r * IN + k0 + c] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < EV_KC; ++kk) {
            const float wv = ws[j][kk];
#pragma unroll
            for (int i = 0; i < SPT; ++i) acc[i] = fmaf(wv, xs[sh * SPT + i][kk], acc[i]);
        }
        __syncthreads();
    }
    const float bj = (j < H) ? b1[j] : 0.f;
#pragma unroll
    for (int i = 0; i < SPT; ++i) hs[sh * SPT + i][j] = (j < H) ? fmaxf(acc[i] + bj, 0.f) : 0.f;
    __syncthreads();
    for (int e = tid; e < EV_TS * OUT; e += EV_THREADS) {
        const int r = e / OUT, o = e % OUT;
        float z = b2[o];
        for (int jj = 0; jj < H; ++jj) z = fmaf(hs[r][jj], W2[(size_t)o * H + jj], z);
        zs[r][o] = z;
    }
    __syncthreads();
    if (tid < EV_TS && s0 + tid < n) {
        int best = 0; float bv = zs[tid][0];
        for (int o = 1; o < OUT; ++o) if (zs[tid][o] > bv) { bv = zs[tid][o]; best = o; }
        const int t = (int)y[s0 + tid];
        if (t >= 0 && t < n_classes && best < n_classes) atomicAdd(&cm[t * n_classes + best], 1);
        if (score1 != nullptr) score1[s0 + tid] = OUT > 1 ? zs[tid][1] : zs[tid][0];      // class-1 logit (AUC of 2-output nets)
    }
}

bool launch_mlp1_eval(const float* row, const float* X, const int64_t* y, int n, int IN, int H, int OUT,
                      int n_classes, int* cm, float* score1, cudaStream_t stream) {
    if (H > EV_HMAX || OUT > 16 || n <= 0) return false;
    mlp1_eval_simt_kernel<<<(n + EV_TS - 1) / EV_TS, EV_THREADS, 0, stream>>>(row, X, y, n, IN, H, OUT,
                                                                             n_classes, cm, score1);
    return true;
}

// force-load this file's kernels (CUDA loads functions lazily; loading one while another kernel spins
// on a cross-GPU flag could deadlock, so the extension loads everything up front)
void preload_eval() {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, mlp1_eval_simt_kernel);
}

}  // namespace gb
