// Flat fused optimizers: one launch over the whole parameter row (generic nn.Module path).
// Semantics: torch.optim.SGD (momentum / dampening / nesterov / weight decay) and Adam / AdamW.
// `scale` optionally multiplies the raw gradient per element (PartitionedTMH 1/age scaling,
// reference gossipy/model/handler.py:514-520).
#include "common.cuh"
#include "ops.h"
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

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
    const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
    int64_t b = (n + 255) / 256;
    return (int)std::max<int64_t>(1, std::min<int64_t>(b, (int64_t)sms * 8));
}

void sgd_step(at::Tensor p, at::Tensor g, int64_t n, double lr, double wd, double momentum,
              c10::optional<at::Tensor> buf, double dampening, bool nesterov, bool first,
              c10::optional<at::Tensor> scale) {
    TORCH_CHECK(p.is_cuda() && g.is_cuda() && p.scalar_type() == at::kFloat && g.scalar_type() == at::kFloat);
    TORCH_CHECK(n <= p.numel() && n <= g.numel());
    if (n == 0) return;
    c10::cuda::CUDAGuard guard(p.device());
    float* bp = (momentum != 0.0 && buf.has_value()) ? buf->data_ptr<float>() : nullptr;
    const float* sp = scale.has_value() ? scale->data_ptr<float>() : nullptr;
    sgd_kernel<<<blocks_for(n), 256, 0, at::cuda::getCurrentCUDAStream()>>>(
        p.data_ptr<float>(), g.data_ptr<float>(), n, (float)lr, (float)wd, (float)momentum, bp,
        (float)dampening, nesterov ? 1 : 0, first ? 1 : 0, sp);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void adam_step(at::Tensor p, at::Tensor g, int64_t n, at::Tensor m, at::Tensor v, int64_t step,
               double lr, double beta1, double beta2, double eps, double wd, bool decoupled) {
    TORCH_CHECK(p.is_cuda() && g.is_cuda() && m.is_cuda() && v.is_cuda());
    if (n == 0) return;
    c10::cuda::CUDAGuard guard(p.device());
    const double bc1 = 1.0 - std::pow(beta1, (double)step);
    const double bc2 = 1.0 - std::pow(beta2, (double)step);
    adam_kernel<<<blocks_for(n), 256, 0, at::cuda::getCurrentCUDAStream()>>>(
        p.data_ptr<float>(), g.data_ptr<float>(), n, m.data_ptr<float>(), v.data_ptr<float>(),
        (float)lr, (float)beta1, (float)beta2, (float)eps, (float)wd, decoupled ? 1 : 0, (float)bc1,
        (float)std::sqrt(bc2));
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

}  // namespace gb
