// All-reduce (weighted mean) of the gossip nodes' parameter rows -- one-shot for small rows, reduce-scatter +
// all-gather ("two-shot", further down) from 4 MiB.  One-shot: the all-to-all round of
// decentralised SGD on a clique (gossipy/simul.py:756-852, node.py:833-845) as ONE kernel per GPU.
//
// Every rank keeps its contribution in a SYMMETRIC buffer (same offset on every GPU).  With NVLink
// SHARP ("NVLS") the buffers are bound to one multicast object: a single
//     multimem.ld_reduce.relaxed.sys.global.add.v4.f32
// makes the NVSwitch fetch the 16 bytes at that offset from ALL GPUs, add them inside the switch and
// return the sum -- each GPU receives P*4 bytes instead of (W-1)*P*4 and issues no per-peer loads.
// Without multicast support the same kernel pulls the W-1 peer buffers with ordinary P2P loads.
//
// The kernel is self-synchronising (no NCCL, no host barrier): CTA 0 publishes "my input is complete"
// to every peer's flag block (st.release.sys), all CTAs wait until every peer has published this
// epoch, reduce, and the last CTA to finish tells every peer "I am done reading" and waits for the
// same message from all of them, so that on kernel exit the symmetric buffers may be overwritten.
#include "common.cuh"
#include "kernels.h"
#include <algorithm>

namespace gb {

constexpr int kNvlsThreads = 256;

GB_DEVICE float4 multimem_ld_reduce_add(const float* mc_addr) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc_addr) : "memory");
    return v;
}

template <bool MULTICAST>
__global__ void __launch_bounds__(kNvlsThreads)
allreduce_mean_kernel(float* __restrict__ out, const AllReduceArgs a, int64_t n, uint32_t* ticket, uint32_t* fault) {
    const int tid = threadIdx.x;
    // ---- start barrier ------------------------------------------------------------------------------
    if (blockIdx.x == 0 && tid < a.world) gb_st_release_sys(a.flags[tid] + a.rank, a.epoch);        // ready[rank] @ peer tid
    if (tid < a.world) {
        gb_wait_flag(a.flags[a.rank] + tid, a.epoch, fault, 2u);
    }
    __syncthreads();
    // ---- reduce ---------------------------------------------------------------------------------------
    const int64_t nvec = n / 4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    float4* o4 = reinterpret_cast<float4*>(out);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + tid; i < nvec; i += stride) {
        float4 acc;
        if (MULTICAST) {
            acc = multimem_ld_reduce_add(a.mc + 4 * i);
        } else {
            acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
            for (int r = 0; r < a.world; ++r) {
                const float4 v = gb_ld_stream(reinterpret_cast<const float4*>(a.bufs[r]) + i);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        o4[i] = make_float4(acc.x * a.scale, acc.y * a.scale, acc.z * a.scale, acc.w * a.scale);
    }
    // ---- end barrier: nobody leaves before every rank has finished reading ----------------------------------
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const bool last = atomicAdd(ticket, 1u) == gridDim.x - 1u;
        if (last) {
            *ticket = 0u;
            for (int r = 0; r < a.world; ++r) gb_st_release_sys(a.flags[r] + a.world + a.rank, a.epoch);  // done[rank] @ peer r
            for (int r = 0; r < a.world; ++r) {
                gb_wait_flag(a.flags[a.rank] + a.world + r, a.epoch, fault, 2u);
            }
        }
    }
}

// ---- two-shot form for large rows ---------------------------------------------------------------------------
// One-shot makes every GPU receive the whole vector (NVLS) or W-1 whole vectors (P2P).  For rows of several MiB
// the switch-side reduction is better used as reduce-scatter + all-gather: rank r reduces ONLY its 1/W slice
// (multimem.ld_reduce), scales it and writes the result back IN PLACE into every rank's symmetric buffer
// (multimem.st: one store, replicated by the switch) -- 2/W of the one-shot traffic per GPU.  Nobody else reads or
// writes slice r in this phase, so no extra buffer is needed.  A second kernel (same stream) waits until every rank
// has published "my slice is stored" and copies the now complete buffer to `out` (local HBM traffic only).
// Two kernels instead of a grid-wide spin: CTAs of one grid are not guaranteed to be co-resident next to other
// streams' kernels, and a CTA spinning on a flag that a not-yet-scheduled CTA must set would deadlock.
GB_DEVICE void multimem_st(float* mc_addr, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(mc_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <bool MULTICAST>
__global__ void __launch_bounds__(kNvlsThreads)
allreduce_rs_ag_kernel(const AllReduceArgs a, int64_t n, uint32_t* ticket, uint32_t* fault) {
    const int tid = threadIdx.x;
    if (blockIdx.x == 0 && tid < a.world) gb_st_release_sys(a.flags[tid] + a.rank, a.epoch);        // ready[rank] @ peer tid
    if (tid < a.world) gb_wait_flag(a.flags[a.rank] + tid, a.epoch, fault, 2u);
    __syncthreads();
    const int64_t nvec = n / 4;
    const int64_t per = (nvec + a.world - 1) / a.world;
    const int64_t lo = (int64_t)a.rank * per, hi = lo + per < nvec ? lo + per : nvec;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + tid; i < hi; i += stride) {
        float4 acc;
        if (MULTICAST) {
            acc = multimem_ld_reduce_add(a.mc + 4 * i);
        } else {
            acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
            for (int r = 0; r < a.world; ++r) {
                const float4 v = gb_ld_stream(reinterpret_cast<const float4*>(a.bufs[r]) + i);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        acc = make_float4(acc.x * a.scale, acc.y * a.scale, acc.z * a.scale, acc.w * a.scale);
        if (MULTICAST) {
            multimem_st(const_cast<float*>(a.mc) + 4 * i, acc);
        } else {
            for (int r = 0; r < a.world; ++r) reinterpret_cast<float4*>(const_cast<float*>(a.bufs[r]))[i] = acc;
        }
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence_system();
        if (atomicAdd(ticket, 1u) == gridDim.x - 1u) {
            *ticket = 0u;
            for (int r = 0; r < a.world; ++r) gb_st_release_sys(a.flags[r] + 2 * kMaxRanks + a.rank, a.epoch);   // stored[rank] @ peer r
        }
    }
}

__global__ void __launch_bounds__(kNvlsThreads)
allreduce_collect_kernel(float* __restrict__ out, const AllReduceArgs a, int64_t n, uint32_t* ticket, uint32_t* fault) {
    const int tid = threadIdx.x;
    if (tid < a.world) gb_wait_flag(a.flags[a.rank] + 2 * kMaxRanks + tid, a.epoch, fault, 2u);      // every slice is in my buffer
    __syncthreads();
    const int64_t nvec = n / 4;
    const float4* src = reinterpret_cast<const float4*>(a.bufs[a.rank]);
    float4* o4 = reinterpret_cast<float4*>(out);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + tid; i < nvec; i += (int64_t)gridDim.x * blockDim.x) o4[i] = src[i];
    // end barrier as in the one-shot kernel: the symmetric buffers may be overwritten once every rank has left
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(ticket, 1u) == gridDim.x - 1u) {
            *ticket = 0u;
            for (int r = 0; r < a.world; ++r) gb_st_release_sys(a.flags[r] + a.world + a.rank, a.epoch);
            for (int r = 0; r < a.world; ++r) gb_wait_flag(a.flags[a.rank] + a.world + r, a.epoch, fault, 2u);
        }
    }
}

static uint32_t* g_nvls_ticket[16] = {nullptr};
constexpr int64_t kTwoShotMinFloats = 1 << 20;        // 4 MiB rows and larger

bool launch_allreduce_mean(float* out, const AllReduceArgs& a, int64_t n, cudaStream_t stream) {
    if (a.world < 1 || a.world > kMaxRanks || n % 4 != 0) return false;
    int dev = 0;
    cudaGetDevice(&dev);
    if (g_nvls_ticket[dev] == nullptr) {
        cudaMalloc(&g_nvls_ticket[dev], 2 * sizeof(uint32_t));
        cudaMemset(g_nvls_ticket[dev], 0, 2 * sizeof(uint32_t));
    }
    const int64_t nvec = n / 4;
    if (n >= kTwoShotMinFloats && a.world > 1) {
        // NOTE: the stored[] flags live at words [2*kMaxRanks, 3*kMaxRanks) of every rank's flag block
        const int blocks = (int)std::min<int64_t>((nvec / a.world + kNvlsThreads * 4 - 1) / (kNvlsThreads * 4), (int64_t)sm_count());
        if (a.mc != nullptr)
            allreduce_rs_ag_kernel<true><<<std::max(1, blocks), kNvlsThreads, 0, stream>>>(a, n, g_nvls_ticket[dev], device_fault_word());
        else
            allreduce_rs_ag_kernel<false><<<std::max(1, blocks), kNvlsThreads, 0, stream>>>(a, n, g_nvls_ticket[dev], device_fault_word());
        const int blocks2 = (int)std::min<int64_t>((nvec + kNvlsThreads * 8 - 1) / (kNvlsThreads * 8), (int64_t)sm_count() * 2);
        allreduce_collect_kernel<<<std::max(1, blocks2), kNvlsThreads, 0, stream>>>(out, a, n, g_nvls_ticket[dev] + 1, device_fault_word());
        return true;
    }
    // few CTAs: the message is small (318 KB for the flagship MLP) and every CTA polls the flags
    int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((nvec + kNvlsThreads * 4 - 1) / (kNvlsThreads * 4),
                                                             (int64_t)sm_count()));
    if (a.mc != nullptr)
        allreduce_mean_kernel<true><<<blocks, kNvlsThreads, 0, stream>>>(out, a, n, g_nvls_ticket[dev], device_fault_word());
    else
        allreduce_mean_kernel<false><<<blocks, kNvlsThreads, 0, stream>>>(out, a, n, g_nvls_ticket[dev], device_fault_word());
    return true;
}

void preload_nvls() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (g_nvls_ticket[dev] == nullptr) {                 // not on the launch path (see merge.cu)
        cudaMalloc(&g_nvls_ticket[dev], 2 * sizeof(uint32_t));
        cudaMemset(g_nvls_ticket[dev], 0, 2 * sizeof(uint32_t));
    }
    cudaFuncAttributes at;
    cudaFuncGetAttributes(&at, allreduce_mean_kernel<true>);
    cudaFuncGetAttributes(&at, allreduce_mean_kernel<false>);
    cudaFuncGetAttributes(&at, allreduce_rs_ag_kernel<true>);
    cudaFuncGetAttributes(&at, allreduce_rs_ag_kernel<false>);
    cudaFuncGetAttributes(&at, allreduce_collect_kernel);
}

}  // namespace gb
