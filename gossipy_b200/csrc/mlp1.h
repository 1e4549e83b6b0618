// Shared parameter block of the fused MLP training kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gb {

struct TrainParams {
    float* row; const float* X; const int64_t* y;
    int n, IN, H, OUT, B, epochs;
    float lr, wd; uint64_t key;
    const int64_t* part_id; const int64_t* ages; int n_parts;
    int Hs, C, nbuf;
    float* dbg;   // optional debug dump (tests): z1 of the first step, see mlp1_train_tc.cu
};

bool mlp1_train_tc(const TrainParams& p, cudaStream_t stream);

}  // namespace gb
