// Flat fused optimizers: one launch over the whole parameter row (generic nn.Module path).
// Semantics: torch.optim.SGD (momentum / dampening / nesterov / weight decay) and Adam / AdamW.
// `scale` optionally multiplies the raw gradient per element (PartitionedTMH 1/age scaling,
// reference gossipy/model/handler.py:514-520).
#include "common.cuh"
#include "kernels.h"
#include <algorithm>
#include <cmath>

namespace gb {

__global__ void __launch_bounds__(256)
sgd_kernel(float* __restrict__ p, const float* __restrict__ g, int64_t n, float lr, float wd,
           float momentum, float* __restrict__ buf, float dampening, int nesterov, int first,
           const float* __restrict__ scale) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        float w = p[i];
        float d = g[i];
        if (scale) d *= scale[i];
        d = fmaf(wd, w, d);
        if (buf) {
            float b = first ? d : fmaf(momentum, buf[i], (1.f - dampening) * d);
            buf[i] = b;
            d = nesterov ? fmaf(momentum, b, d) : b;
        }
        p[i] = fmaf(-lr, d, w);
    }
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, int64_t n, float* __restrict__ m,
            float* __restrict__ v, float lr, float b1, float b2, float eps, float wd, int decoupled,
            float bc1, float bc2_sqrt) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        float w = p[i];
        float d = g[i];
        if (wd != 0.f) { if (decoupled) w *= (1.f - lr * wd); else d = fmaf(wd, w, d); }
        const float mi = fmaf(b1, m[i], (1.f - b1) * d);
        const float vi = fmaf(b2, v[i], (1.f - b2) * d * d);
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = w - (lr / bc1) * (mi / denom);
    }
}

static int blocks_for(int64_t n) {
    int64_t b = (n + 255) / 256;
    return (int)std::max<int64_t>(1, std::min<int64_t>(b, (int64_t)sm_count() * 8));
}

void launch_sgd(float* p, const float* g, int64_t n, float lr, float wd, float momentum, float* buf,
                float dampening, bool nesterov, bool first, const float* scale, cudaStream_t stream) {
    if (n <= 0) return;
    sgd_kernel<<<blocks_for(n), 256, 0, stream>>>(p, g, n, lr, wd, momentum, momentum != 0.f ? buf : nullptr,
                                                  dampening, nesterov ? 1 : 0, first ? 1 : 0, scale);
}

void launch_adam(float* p, const float* g, int64_t n, float* m, float* v, int64_t step, float lr,
                 float beta1, float beta2, float eps, float wd, bool decoupled, cudaStream_t stream) {
    if (n <= 0) return;
    const double bc1 = 1.0 - std::pow((double)beta1, (double)step);
    const double bc2 = 1.0 - std::pow((double)beta2, (double)step);
    adam_kernel<<<blocks_for(n), 256, 0, stream>>>(p, g, n, m, v, lr, beta1, beta2, eps, wd,
                                                   decoupled ? 1 : 0, (float)bc1, (float)std::sqrt(bc2));
}

// force-load this file's kernels (CUDA loads functions lazily; loading one while another kernel spins
// on a cross-GPU flag could deadlock, so the extension loads everything up front)
void preload_optim() {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, sgd_kernel);
    cudaFuncGetAttributes(&a, adam_kernel);
}

// the keyed sample order of one epoch as an index vector (generic autograd path: the mini-batches are gathered on the
// device with it; producing it here instead of copying a host-built vector keeps the host free to run ahead)
__global__ void keyed_perm_kernel(int64_t* __restrict__ out, int n, uint64_t key) {
    GbPerm perm; perm.init((uint32_t)n, key);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = (int64_t)perm((uint32_t)i);
}
// k keyed draws from range(n) with replacement (model/handler.py::SamplingTMH.draw_sample, ops/torch_ref.py::keyed_randint)
__global__ void keyed_randint_kernel(int64_t* __restrict__ out, int64_t k, uint64_t n, uint64_t key) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < k; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (int64_t)(gb_mix64(key ^ (uint64_t)i) % n);
}
void launch_keyed_randint(int64_t* out, int64_t k, int64_t n, uint64_t key, cudaStream_t stream) {
    if (k <= 0 || n <= 0) return;
    keyed_randint_kernel<<<(int)std::min<int64_t>(1024, (k + 255) / 256), 256, 0, stream>>>(out, k, (uint64_t)n, key);
}
void launch_keyed_perm(int64_t* out, int n, uint64_t key, cudaStream_t stream) {
    if (n <= 0) return;
    keyed_perm_kernel<<<std::min(1024, (n + 255) / 256), 256, 0, stream>>>(out, n, key);
}

}  // namespace gb
