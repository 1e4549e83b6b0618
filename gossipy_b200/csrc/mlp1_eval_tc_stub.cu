#include <cuda_runtime.h>
#include <stdint.h>
namespace gb {
bool mlp1_eval_tc(const float*, const void*, const int64_t*, int, int, int, int, int, int*, cudaStream_t) { return false; }
}
