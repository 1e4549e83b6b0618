// Fused merge kernels over flat parameter rows.
//
// One launch reads the local row and the peer row -- which may be a mapped pointer into ANOTHER
// GPU's HBM (pull model over NVLink/NVSwitch: loads are pipelined, no remote atomics) -- and
// writes the weighted combination back in place.  No NCCL call and no separate elementwise
// launch sits on this path.  The generic (w_dst, w_src) pair covers: uniform average (.5,.5),
// age-weighted / limited merge (a/(a+b), b/(a+b)), adopt / pass-through / snapshot (0,1).
//
// Cross-GPU handshake (PeerSync, kernels.h) is part of the same kernel: every CTA spins on the
// owner's `ready` generation flag (ld.acquire.sys) before its first peer load, and the last CTA to
// finish acknowledges the read with one red.release.sys.add on the owner's `done` counter.
//
// Reference semantics: gossipy/model/handler.py:260-280, 666-688, 695-715; sampling.py:76-107,
// 201-234.
#include "common.cuh"
#include "kernels.h"
#include <algorithm>

namespace gb {

constexpr int kMergeThreads = 256;
constexpr int kUnroll = 4;  // 4 x 128-bit loads in flight per thread per operand (peer latency ~2 us)

// ---- handshake helpers ---------------------------------------------------------------------------
// Tickets: zero-initialised CTA-arrival counters, one per handshaking launch, returned to zero by the last
// CTA.  The ring is far larger than the number of launches the host can run ahead (it waits for the metrics
// of round r while it enqueues round r+1); should two in-flight launches ever share a ticket the arrival count
// exceeds the grid size, which the kernels report as fault bit 4 instead of silently mis-acknowledging.
constexpr int kTickets = 1 << 16;
static uint32_t* g_tickets[16] = {nullptr};
static uint32_t g_ticket_next[16] = {0};
static uint32_t* g_fault[16] = {nullptr};

// per-device scratch of the handshake protocol; allocated by preload_merge() (extension start-up), never on
// the launch path: cudaMalloc is an implicit device-wide barrier and kernels of other GPUs may be spinning
static void ensure_device_scratch(int dev) {
    if (g_tickets[dev] == nullptr) {
        cudaMalloc(&g_tickets[dev], kTickets * sizeof(uint32_t));
        cudaMemset(g_tickets[dev], 0, kTickets * sizeof(uint32_t));
        cudaMalloc(&g_fault[dev], sizeof(uint32_t));
        cudaMemset(g_fault[dev], 0, sizeof(uint32_t));
    }
}
static uint32_t* next_ticket() {
    int dev = 0;
    cudaGetDevice(&dev);
    ensure_device_scratch(dev);
    return g_tickets[dev] + (g_ticket_next[dev]++ % kTickets);
}
uint32_t* device_fault_word() {
    int dev = 0;
    cudaGetDevice(&dev);
    ensure_device_scratch(dev);
    return g_fault[dev];
}

GB_DEVICE void peer_wait(const PeerSync& s) {
    if (s.ready != nullptr) {
        if (threadIdx.x == 0) gb_wait_flag(s.ready, s.gen, s.fault);
        __syncthreads();
    }
}
GB_DEVICE void peer_done(const PeerSync& s, uint32_t* ticket) {
    if (s.done != nullptr) {
        __syncthreads();                         // every thread of the CTA has consumed its peer loads
        if (threadIdx.x == 0) {
            __threadfence();
            const uint32_t total = gridDim.x * gridDim.y;
            const uint32_t seen = atomicAdd(ticket, 1u);
            if (seen == total - 1u) {
                *ticket = 0u;
                gb_red_release_sys_add(s.done, 1u);
            } else if (seen >= total && s.fault != nullptr) {
                atomicOr(s.fault, 4u);               // ticket shared by two in-flight launches
            }
        }
    }
}

__global__ void __launch_bounds__(kMergeThreads)
merge_pair_kernel(float* __restrict__ dst, const float* __restrict__ src, float wd, float ws,
                  int64_t lo, int64_t hi, PeerSync sync, uint32_t* ticket) {
    peer_wait(sync);
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
    peer_done(sync, ticket);
}

// Strided blocks (partitioned-model merge): block s = (start, n_runs, run_len, stride); only the
// bytes of the partition are fetched from the peer.
__global__ void __launch_bounds__(kMergeThreads)
merge_segments_kernel(float* __restrict__ dst, const float* __restrict__ src,
                      const int64_t* __restrict__ seg, int n_seg, float wd, float ws, PeerSync sync,
                      uint32_t* ticket) {
    peer_wait(sync);
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
    peer_done(sync, ticket);
}

// Sampled merge: gather of 4-byte words from the peer (documented lower NVLink efficiency).
// Indices may repeat: phase 1 computes every merged value from the UNMODIFIED rows into scratch,
// phase 2 scatters (duplicates then write identical values).
__global__ void __launch_bounds__(kMergeThreads)
merge_indexed_kernel(float* __restrict__ dst, const float* __restrict__ src,
                     const int64_t* __restrict__ idx, int64_t n, float wd, float ws,
                     float* __restrict__ scratch, PeerSync sync, uint32_t* ticket) {
    peer_wait(sync);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = idx[i];
        scratch[i] = wd * dst[p] + ws * gb_ld_stream1(src + p);
    }
    peer_done(sync, ticket);
}
__global__ void __launch_bounds__(kMergeThreads)
scatter_kernel(float* __restrict__ dst, const int64_t* __restrict__ idx, int64_t n,
               const float* __restrict__ vals) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        dst[idx[i]] = vals[i];
}

// k-way weighted merge (All2All / PENS): pointer table passed by value in the launch parameters.
struct KwayArgs {
    const float* src[kMaxWay]; float w[kMaxWay]; float w0; int k;
    const uint32_t* ready[kMaxWay]; uint32_t gen[kMaxWay]; uint32_t* done[kMaxWay];
    uint32_t* fault;
};

__global__ void __launch_bounds__(kMergeThreads)
merge_kway_kernel(float* __restrict__ dst, const KwayArgs a, int64_t n, uint32_t* ticket) {
    if (threadIdx.x < a.k && a.ready[threadIdx.x] != nullptr) {
        gb_wait_flag(a.ready[threadIdx.x], a.gen[threadIdx.x], a.fault);
    }
    __syncthreads();
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
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ticket, 1u) == gridDim.x - 1u) {
            *ticket = 0u;
            for (int j = 0; j < a.k; ++j)
                if (a.done[j] != nullptr) gb_red_release_sys_add(a.done[j], 1u);
        }
    }
}

// ---- stand-alone flag operations ---------------------------------------------------------------------
__global__ void flag_signal_kernel(uint32_t* flag, uint32_t value) {
    __threadfence_system();
    gb_st_release_sys(flag, value);
}
__global__ void flag_wait_kernel(const uint32_t* flag, uint32_t value, uint32_t* fault) {
    gb_wait_flag(flag, value, fault, 16u);                                     // flags only grow; 16 = a stream-ordered wait (acks of a row's readers)
}
__global__ void flag_add_kernel(uint32_t* flag, uint32_t value) {
    __threadfence_system();
    gb_red_release_sys_add(flag, value);
}

// ---- launchers -----------------------------------------------------------------------------------------
int sm_count() {
    static int cached[16] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (cached[dev] == 0) cudaDeviceGetAttribute(&cached[dev], cudaDevAttrMultiProcessorCount, dev);
    return cached[dev];
}

static int grid_for(int64_t work_items, int per_thread) {
    int64_t blocks = (work_items + (int64_t)kMergeThreads * per_thread - 1) / ((int64_t)kMergeThreads * per_thread);
    const int64_t cap = (int64_t)sm_count() * 8;  // 8 resident CTAs of 256 threads per SM
    return (int)std::max<int64_t>(1, std::min(blocks, cap));
}

void launch_merge_pair(float* dst, const float* src, float w_dst, float w_src, int64_t lo, int64_t hi,
                       PeerSync sync, cudaStream_t stream) {
    if (hi <= lo) return;
    uint32_t* ticket = sync.done ? next_ticket() : nullptr;
    merge_pair_kernel<<<grid_for((hi - lo) / 4 + 1, kUnroll), kMergeThreads, 0, stream>>>(
        dst, src, w_dst, w_src, lo, hi, sync, ticket);
}

void launch_merge_segments(float* dst, const float* src, const int64_t* seg, int n_seg, float w_dst,
                           float w_src, PeerSync sync, cudaStream_t stream) {
    if (n_seg <= 0) return;
    uint32_t* ticket = sync.done ? next_ticket() : nullptr;
    dim3 grid(8, n_seg < 1024 ? n_seg : 1024);
    merge_segments_kernel<<<grid, kMergeThreads, 0, stream>>>(dst, src, seg, n_seg, w_dst, w_src, sync, ticket);
}

void launch_merge_indexed(float* dst, const float* src, const int64_t* idx, int64_t n, float w_dst,
                          float w_src, float* scratch, PeerSync sync, cudaStream_t stream) {
    if (n <= 0) return;
    uint32_t* ticket = sync.done ? next_ticket() : nullptr;
    merge_indexed_kernel<<<grid_for(n, 1), kMergeThreads, 0, stream>>>(dst, src, idx, n, w_dst, w_src,
                                                                       scratch, sync, ticket);
    scatter_kernel<<<grid_for(n, 1), kMergeThreads, 0, stream>>>(dst, idx, n, scratch);
}

void launch_merge_kway(float* dst, const float* const* srcs, const float* weights, int k, int64_t n,
                       const PeerSync* syncs, cudaStream_t stream) {
    float w0 = weights[0];
    int done = 0;
    while (done < k) {
        KwayArgs a;
        a.k = std::min(kMaxWay, k - done);
        a.w0 = w0;
        a.fault = device_fault_word();
        for (int j = 0; j < kMaxWay; ++j) { a.ready[j] = nullptr; a.done[j] = nullptr; a.gen[j] = 0; a.src[j] = nullptr; a.w[j] = 0.f; }
        for (int j = 0; j < a.k; ++j) {
            a.src[j] = srcs[done + j];
            a.w[j] = weights[1 + done + j];
            if (syncs) { a.ready[j] = syncs[done + j].ready; a.gen[j] = syncs[done + j].gen; a.done[j] = syncs[done + j].done; }
        }
        merge_kway_kernel<<<grid_for(n / 4 + 1, 1), kMergeThreads, 0, stream>>>(dst, a, n, next_ticket());
        done += a.k;
        w0 = 1.f;  // later chunks accumulate onto the partial result
    }
}

void launch_flag_signal(uint32_t* flag, uint32_t value, cudaStream_t stream) {
    flag_signal_kernel<<<1, 1, 0, stream>>>(flag, value);
}
void launch_flag_wait(const uint32_t* flag, uint32_t value, cudaStream_t stream) {
    flag_wait_kernel<<<1, 1, 0, stream>>>(flag, value, device_fault_word());
}
void launch_flag_add(uint32_t* flag, uint32_t value, cudaStream_t stream) {
    flag_add_kernel<<<1, 1, 0, stream>>>(flag, value);
}

// force-load this file's kernels (CUDA loads functions lazily; loading one while another kernel spins
// on a cross-GPU flag could deadlock, so the extension loads everything up front)
void preload_merge() {
    device_fault_word();
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, merge_pair_kernel);
    cudaFuncGetAttributes(&a, merge_segments_kernel);
    cudaFuncGetAttributes(&a, merge_indexed_kernel);
    cudaFuncGetAttributes(&a, scatter_kernel);
    cudaFuncGetAttributes(&a, merge_kway_kernel);
    cudaFuncGetAttributes(&a, flag_signal_kernel);
    cudaFuncGetAttributes(&a, flag_wait_kernel);
    cudaFuncGetAttributes(&a, flag_add_kernel);
}

}  // namespace gb
