// Fused merge kernels over flat parameter rows.
//
// One launch reads the local row and the peer row -- which may be a mapped pointer into ANOTHER
// GPU's HBM (pull model over NVLink/NVSwitch: loads are pipelined, no remote atomics) -- and
// writes the weighted combination back in place.  No NCCL call and no separate elementwise
// launch sits on this path.  The generic (w_dst, w_src) pair covers: uniform average (.5,.5),
// age-weighted / limited merge (a/(a+b), b/(a+b)), adopt / pass-through / snapshot (0,1).
//
// Reference semantics: gossipy/model/handler.py:260-280, 666-688, 695-715; sampling.py:76-107,
// 201-234.
#include "common.cuh"
#include "ops.h"
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

namespace gb {

constexpr int kMergeThreads = 256;
constexpr int kUnroll = 4;  // 4 x 128-bit loads in flight per thread per operand (peer latency ~2 us)

__global__ void __launch_bounds__(kMergeThreads)
merge_pair_kernel(float* __restrict__ dst, const float* __restrict__ src, float wd, float ws,
                  int64_t lo, int64_t hi) {
    // scalar head until dst is 16-byte aligned, vector body, scalar tail
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    int64_t head = lo;
    while ((head < hi) && ((((uintptr_t)(dst + head)) & 15u) != 0)) ++head;
    const bool src_aligned = ((((uintptr_t)(src + head)) & 15u) == 0);
    for (int64_t i = lo + tid; i < head; i += nthreads)
        dst[i] = (wd == 0.f ? 0.f : wd * dst[i]) + ws * gb_ld_stream1(src + i);
    const int64_t nvec = (hi - head) / 4;
    float4* d4 = reinterpret_cast<float4*>(dst + head);
    if (src_aligned) {
        const float4* s4 = reinterpret_cast<const float4*>(src + head);
        int64_t i = tid;
        for (; i + (kUnroll - 1) * nthreads < nvec; i += kUnroll * nthreads) {
            float4 s[kUnroll], d[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) s[u] = gb_ld_stream(s4 + i + u * nthreads);
            if (wd != 0.f) {
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) d[u] = d4[i + u * nthreads];
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                float4 r;
                if (wd != 0.f) {
                    r.x = wd * d[u].x + ws * s[u].x; r.y = wd * d[u].y + ws * s[u].y;
                    r.z = wd * d[u].z + ws * s[u].z; r.w = wd * d[u].w + ws * s[u].w;
                } else {
                    r.x = ws * s[u].x; r.y = ws * s[u].y; r.z = ws * s[u].z; r.w = ws * s[u].w;
                }
                d4[i + u * nthreads] = r;
            }
        }
        for (; i < nvec; i += nthreads) {
            float4 s = gb_ld_stream(s4 + i), r;
            if (wd != 0.f) {
                float4 d = d4[i];
                r.x = wd * d.x + ws * s.x; r.y = wd * d.y + ws * s.y;
                r.z = wd * d.z + ws * s.z; r.w = wd * d.w + ws * s.w;
            } else { r.x = ws * s.x; r.y = ws * s.y; r.z = ws * s.z; r.w = ws * s.w; }
            d4[i] = r;
        }
    } else {
        for (int64_t i = tid; i < nvec; i += nthreads) {
            const float* s = src + head + 4 * i;
            float4 r, d = (wd != 0.f) ? d4[i] : make_float4(0, 0, 0, 0);
            r.x = wd * d.x + ws * gb_ld_stream1(s);     r.y = wd * d.y + ws * gb_ld_stream1(s + 1);
            r.z = wd * d.z + ws * gb_ld_stream1(s + 2); r.w = wd * d.w + ws * gb_ld_stream1(s + 3);
            d4[i] = r;
        }
    }
    for (int64_t i = head + 4 * nvec + tid; i < hi; i += nthreads)
        dst[i] = (wd == 0.f ? 0.f : wd * dst[i]) + ws * gb_ld_stream1(src + i);
}

// Strided blocks (partitioned-model merge): block s = (start, n_runs, run_len, stride); only the
// bytes of the partition are fetched from the peer.
__global__ void __launch_bounds__(kMergeThreads)
merge_segments_kernel(float* __restrict__ dst, const float* __restrict__ src,
                      const int64_t* __restrict__ seg, int n_seg, float wd, float ws) {
    for (int s = blockIdx.y; s < n_seg; s += gridDim.y) {
        const int64_t start = seg[4 * s], n_runs = seg[4 * s + 1], run_len = seg[4 * s + 2],
                      stride = seg[4 * s + 3];
        const int64_t total = n_runs * run_len;
        for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
             e += (int64_t)gridDim.x * blockDim.x) {
            const int64_t r = e / run_len, c = e - r * run_len;
            const int64_t pos = start + r * stride + c;
            dst[pos] = wd * dst[pos] + ws * gb_ld_stream1(src + pos);
        }
    }
}

// Sampled merge: gather of 4-byte words from the peer (documented lower NVLink efficiency).
// Duplicated indices are benign: every duplicate computes the same value from the same inputs only
// if reads precede writes -- so read both operands first, then write (two-phase within a thread);
// duplicates across threads race on identical values.
__global__ void __launch_bounds__(kMergeThreads)
merge_indexed_kernel(float* __restrict__ dst, const float* __restrict__ src,
                     const int64_t* __restrict__ idx, int64_t n, float wd, float ws,
                     float* __restrict__ scratch) {
    // phase 1: scratch[i] = merged value ; phase 2 (second launch with scratch==nullptr swap) writes
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = idx[i];
        scratch[i] = wd * dst[p] + ws * gb_ld_stream1(src + p);
    }
}
__global__ void __launch_bounds__(kMergeThreads)
scatter_kernel(float* __restrict__ dst, const int64_t* __restrict__ idx, int64_t n,
               const float* __restrict__ vals) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        dst[idx[i]] = vals[i];
}

// k-way weighted merge (All2All / PENS): pointer table passed by value in the launch parameters.
constexpr int kMaxWay = 32;
struct KwayArgs { const float* src[kMaxWay]; float w[kMaxWay]; float w0; int k; };

__global__ void __launch_bounds__(kMergeThreads)
merge_kway_kernel(float* __restrict__ dst, KwayArgs a, int64_t n) {
    const int64_t nvec = n / 4;
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
         i += (int64_t)gridDim.x * blockDim.x) {
        float4 d = d4[i];
        float4 acc = make_float4(a.w0 * d.x, a.w0 * d.y, a.w0 * d.z, a.w0 * d.w);
#pragma unroll 4
        for (int j = 0; j < a.k; ++j) {
            const float4 s = gb_ld_stream(reinterpret_cast<const float4*>(a.src[j]) + i);
            const float w = a.w[j];
            acc.x += w * s.x; acc.y += w * s.y; acc.z += w * s.z; acc.w += w * s.w;
        }
        d4[i] = acc;
    }
    for (int64_t i = 4 * nvec + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        float acc = a.w0 * dst[i];
        for (int j = 0; j < a.k; ++j) acc += a.w[j] * gb_ld_stream1(a.src[j] + i);
        dst[i] = acc;
    }
}

static int grid_for(int64_t work_items, int per_thread) {
    const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
    int64_t blocks = (work_items + (int64_t)kMergeThreads * per_thread - 1) / ((int64_t)kMergeThreads * per_thread);
    const int64_t cap = (int64_t)sms * 8;  // 8 resident CTAs of 256 threads per SM
    if (blocks < 1) blocks = 1;
    if (blocks > cap) blocks = cap;
    return (int)blocks;
}

void merge_pair(at::Tensor dst, at::Tensor src, double w_dst, double w_src, int64_t lo, int64_t hi) {
    TORCH_CHECK(dst.is_cuda() && dst.scalar_type() == at::kFloat && dst.is_contiguous());
    TORCH_CHECK(src.scalar_type() == at::kFloat && src.is_contiguous());
    TORCH_CHECK(0 <= lo && lo <= hi && hi <= dst.numel() && hi <= src.numel());
    if (hi == lo) return;
    c10::cuda::CUDAGuard guard(dst.device());
    auto stream = at::cuda::getCurrentCUDAStream();
    merge_pair_kernel<<<grid_for((hi - lo) / 4 + 1, kUnroll), kMergeThreads, 0, stream>>>(
        dst.data_ptr<float>(), src.data_ptr<float>(), (float)w_dst, (float)w_src, lo, hi);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void merge_segments(at::Tensor dst, at::Tensor src, at::Tensor seg, double w_dst, double w_src) {
    TORCH_CHECK(dst.is_cuda() && seg.is_cuda() && seg.scalar_type() == at::kLong && seg.is_contiguous());
    TORCH_CHECK(seg.dim() == 2 && seg.size(1) == 4);
    const int n_seg = (int)seg.size(0);
    if (n_seg == 0) return;
    c10::cuda::CUDAGuard guard(dst.device());
    auto stream = at::cuda::getCurrentCUDAStream();
    dim3 grid(8, n_seg < 1024 ? n_seg : 1024);
    merge_segments_kernel<<<grid, kMergeThreads, 0, stream>>>(
        dst.data_ptr<float>(), src.data_ptr<float>(), seg.data_ptr<int64_t>(), n_seg,
        (float)w_dst, (float)w_src);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void merge_indexed(at::Tensor dst, at::Tensor src, at::Tensor idx, double w_dst, double w_src) {
    TORCH_CHECK(dst.is_cuda() && idx.is_cuda() && idx.scalar_type() == at::kLong && idx.is_contiguous());
    const int64_t n = idx.numel();
    if (n == 0) return;
    c10::cuda::CUDAGuard guard(dst.device());
    auto stream = at::cuda::getCurrentCUDAStream();
    auto scratch = at::empty({n}, dst.options());
    merge_indexed_kernel<<<grid_for(n, 1), kMergeThreads, 0, stream>>>(
        dst.data_ptr<float>(), src.data_ptr<float>(), idx.data_ptr<int64_t>(), n, (float)w_dst,
        (float)w_src, scratch.data_ptr<float>());
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    scatter_kernel<<<grid_for(n, 1), kMergeThreads, 0, stream>>>(
        dst.data_ptr<float>(), idx.data_ptr<int64_t>(), n, scratch.data_ptr<float>());
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void merge_kway(at::Tensor dst, std::vector<at::Tensor> srcs, std::vector<double> weights) {
    TORCH_CHECK(dst.is_cuda() && dst.scalar_type() == at::kFloat && dst.is_contiguous());
    TORCH_CHECK(weights.size() == srcs.size() + 1, "need one weight per model incl. self");
    TORCH_CHECK((((uintptr_t)dst.data_ptr<float>()) & 15u) == 0, "row must be 16-byte aligned");
    c10::cuda::CUDAGuard guard(dst.device());
    auto stream = at::cuda::getCurrentCUDAStream();
    const int64_t n = dst.numel();
    double w0 = weights[0];
    size_t done = 0;
    if (srcs.empty()) {
        dst.mul_(w0);
        return;
    }
    while (done < srcs.size()) {
        KwayArgs a;
        a.k = (int)std::min<size_t>(kMaxWay, srcs.size() - done);
        a.w0 = (float)w0;
        for (int j = 0; j < a.k; ++j) {
            const at::Tensor& s = srcs[done + j];
            TORCH_CHECK(s.scalar_type() == at::kFloat && s.is_contiguous() && s.numel() >= n);
            TORCH_CHECK((((uintptr_t)s.data_ptr<float>()) & 15u) == 0);
            a.src[j] = s.data_ptr<float>();
            a.w[j] = (float)weights[1 + done + j];
        }
        merge_kway_kernel<<<grid_for(n / 4 + 1, 1), kMergeThreads, 0, stream>>>(dst.data_ptr<float>(), a, n);
        C10_CUDA_KERNEL_LAUNCH_CHECK();
        done += a.k;
        w0 = 1.0;  // later chunks accumulate onto the partial result
    }
}

}  // namespace gb
