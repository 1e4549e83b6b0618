// Multi-process runtime: CUDA-IPC shared arenas, peer-mapped tensors and cross-GPU flags.
//
// One process per GPU (torch.distributed is used only to exchange the 64-byte IPC handles).  Each
// rank cudaMalloc's its arena, exports it, and maps every peer's arena; a peer row is then just a
// device pointer that the merge kernels dereference (loads travel over NVLink / NVSwitch).
// Cross-GPU ordering uses monotonically increasing 32-bit flags living in the arenas: the
// producer's stream runs `flag_signal` (st.release.sys after the data kernel), the consumer's
// stream runs `flag_wait` (ld.acquire.sys spin in a 1-thread kernel) before the merge kernel.
#include "common.cuh"
#include "ops.h"
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda.h>

namespace gb {

int64_t ipc_alloc(int64_t nbytes) {
    void* p = nullptr;
    C10_CUDA_CHECK(cudaMalloc(&p, (size_t)nbytes));
    C10_CUDA_CHECK(cudaMemset(p, 0, (size_t)nbytes));
    return (int64_t)(uintptr_t)p;
}

void ipc_free(int64_t ptr) { C10_CUDA_CHECK(cudaFree((void*)(uintptr_t)ptr)); }

pybind11::bytes ipc_get_handle(int64_t ptr) {
    cudaIpcMemHandle_t h;
    C10_CUDA_CHECK(cudaIpcGetMemHandle(&h, (void*)(uintptr_t)ptr));
    return pybind11::bytes(reinterpret_cast<const char*>(&h), sizeof(h));
}

int64_t ipc_open_handle(pybind11::bytes handle) {
    std::string s = handle;
    TORCH_CHECK(s.size() == sizeof(cudaIpcMemHandle_t), "bad IPC handle size");
    cudaIpcMemHandle_t h;
    memcpy(&h, s.data(), sizeof(h));
    void* p = nullptr;
    C10_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    return (int64_t)(uintptr_t)p;
}

void ipc_close_handle(int64_t ptr) { C10_CUDA_CHECK(cudaIpcCloseMemHandle((void*)(uintptr_t)ptr)); }

at::Tensor tensor_from_ptr(int64_t ptr, std::vector<int64_t> sizes, int64_t device, bool as_int32) {
    auto opts = at::TensorOptions().device(at::kCUDA, (c10::DeviceIndex)device)
                    .dtype(as_int32 ? at::kInt : at::kFloat);
    return at::from_blob((void*)(uintptr_t)ptr, sizes, [](void*) {}, opts);
}

__global__ void flag_signal_kernel(uint32_t* flag, uint32_t value) {
    __threadfence_system();
    gb_st_release_sys(flag, value);
}

__global__ void flag_wait_kernel(const uint32_t* flag, uint32_t value) {
    // flags only grow; wrap-around safe comparison
    while ((int32_t)(gb_ld_acquire_sys(flag) - value) < 0) __nanosleep(64);
}

void flag_signal(int64_t flag_ptr, int64_t value) {
    flag_signal_kernel<<<1, 1, 0, at::cuda::getCurrentCUDAStream()>>>((uint32_t*)(uintptr_t)flag_ptr, (uint32_t)value);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void flag_wait(int64_t flag_ptr, int64_t value) {
    flag_wait_kernel<<<1, 1, 0, at::cuda::getCurrentCUDAStream()>>>((const uint32_t*)(uintptr_t)flag_ptr, (uint32_t)value);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

int64_t device_sm_count() { return at::cuda::getCurrentDeviceProperties()->multiProcessorCount; }

}  // namespace gb
