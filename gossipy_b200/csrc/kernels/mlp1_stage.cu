// Device-side data loader of the CTA-pair tcgen05 training kernel (mlp1_train_tc3.cu).
//
// One local update visits the node's shard in a keyed pseudo-random order (the engine's Feistel
// permutation, one key per epoch).  Instead of gathering 32 random rows per SGD step INSIDE the
// latency-critical training kernel, this kernel -- wide, one CTA per (step, feature half)
// -- writes the whole update's mini-batches ahead of time, already shuffled and already in the two
// shared-memory images the tensor core wants, so that the training kernel fetches each step's operands
// with two contiguous bulk copies (cp.async.bulk -> UBLKCP) and never touches an index:
//
//   XF[s][r]  forward operand  X_s[:, half r]   (N = batch rows, K = features)  K-major core matrices
//             [4 batch groups][FP/4 chunks][8 rows][16 B]
//   XT[s][r]  update operand   X_s[:, half r]^T (N = features,   K = batch)     K-major core matrices
//             [FP/8 feature groups][8 batch chunks][8 feature rows][16 B = 4 samples]
//   YS[s][32] labels (int32), -1 for the padding rows of a short last batch
//
// tcgen05 cannot transpose 32-bit (tf32) operands, hence the explicit second image.  Padding rows /
// columns are written as zeros.  r = CTA of the pair (feature half), FP = padded features per CTA.
#include "common.cuh"
#include "kernels.h"

namespace gb {

constexpr int ST_THREADS = 256;
constexpr int ST_B = 32;

__global__ void __launch_bounds__(ST_THREADS)
mlp1_stage_kernel(const float* __restrict__ X, const int64_t* __restrict__ y, int n, int IN, int B,
                  int epochs, uint64_t key, int FPC, int FP, float* __restrict__ xf,
                  float* __restrict__ xt, int* __restrict__ ys) {
    extern __shared__ __align__(16) float tile[];       // [32][FP + 1]
    __shared__ int ids[ST_B];
    const int s = blockIdx.x, r = blockIdx.y, tid = threadIdx.x;
    const int spe = (n + B - 1) / B;
    const int e = epochs > 0 ? s / spe : 0;
    const int pos = epochs > 0 ? (s % spe) * B : 0;
    const int bcur = min(B, n - pos);
    const int f0 = r * FPC;
    const int fcnt = max(0, min(FPC, IN - f0));
    const int ld = FP + 1;
    if (tid < ST_B) {
        int id = -1;
        if (tid < bcur) {
            GbPerm perm; perm.init((uint32_t)n, gb_mix64(key ^ (uint64_t)e));
            id = (int)perm((uint32_t)(pos + tid));
        }
        ids[tid] = id;
        if (r == 0) ys[(size_t)s * ST_B + tid] = id >= 0 ? (int)y[id] : -1;
    }
    __syncthreads();
    // gather the 32 rows of this mini-batch (my feature half) into shared memory.  The rows are random,
    // so the loop is a chain of DRAM round trips unless several loads are in flight per thread: 128-bit
    // loads, four issued back to back before the first one is consumed.
    if ((IN & 3) == 0) {                                 // f0, fcnt and every row start are multiples of 4 floats
        const int nch = FP >> 2, total = ST_B * nch;
        constexpr int U = 4;
        for (int i0 = tid; i0 < total; i0 += U * ST_THREADS) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * ST_THREADS;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < total) {
                    const int b = i / nch, c = (i - b * nch) << 2;
                    const int id = ids[b];
                    if (id >= 0 && c < fcnt) v[u] = __ldg(reinterpret_cast<const float4*>(X + (size_t)id * IN + f0 + c));
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * ST_THREADS;
                if (i < total) {
                    const int b = i / nch, c = (i - b * nch) << 2;
                    float* d = tile + b * ld + c;
                    d[0] = v[u].x; d[1] = v[u].y; d[2] = v[u].z; d[3] = v[u].w;
                }
            }
        }
    } else {
        for (int i = tid; i < ST_B * FP; i += ST_THREADS) {
            const int b = i / FP, c = i - b * FP;
            const int id = ids[b];
            tile[b * ld + c] = (id >= 0 && c < fcnt) ? X[(size_t)id * IN + f0 + c] : 0.f;
        }
    }
    __syncthreads();
    const size_t tile_floats = (size_t)ST_B * FP;
    const size_t base = ((size_t)s * 2 + r) * tile_floats;
    // XF: chunk q = ((g * nchunk + c) * 8 + r8) holds X[g*8 + r8][4c .. 4c+3]
    {
        const int nchunk = FP >> 2;
        float4* dst = reinterpret_cast<float4*>(xf + base);
        for (int q = tid; q < ST_B * nchunk; q += ST_THREADS) {
            const int r8 = q & 7, gc = q >> 3, c = gc % nchunk, g = gc / nchunk;
            const float* src = tile + (g * 8 + r8) * ld + 4 * c;
            dst[q] = make_float4(src[0], src[1], src[2], src[3]);
        }
    }
    // XT: chunk q = ((fg * 8 + bc) * 8 + fr) holds X[4bc .. 4bc+3][fg*8 + fr]
    {
        float4* dst = reinterpret_cast<float4*>(xt + base);
        const int nq = (FP >> 3) * 64;
        for (int q = tid; q < nq; q += ST_THREADS) {
            const int fr = q & 7, bc = (q >> 3) & 7, fg = q >> 6;
            const float* src = tile + (bc * 4) * ld + fg * 8 + fr;
            dst[q] = make_float4(src[0], src[ld], src[2 * ld], src[3 * ld]);
        }
    }
}

size_t mlp1_stage_bytes(int n, int IN, int B, int epochs, int* FPC_out, int* FP_out, int* steps_out) {
    const int FPC = ((IN / 2) + 3) & ~3;
    const int FP = (FPC + 15) & ~15;
    const int spe = (n + B - 1) / B;
    const int steps = epochs > 0 ? epochs * spe : 1;
    if (FPC_out) *FPC_out = FPC;
    if (FP_out) *FP_out = FP;
    if (steps_out) *steps_out = steps;
    const size_t tile = (size_t)ST_B * FP * 4;
    return (size_t)steps * (4 * tile + ST_B * 4);
}

// staging buffer layout: [XF: steps*2 tiles][XT: steps*2 tiles][YS: steps*32 int]
bool launch_mlp1_stage(const float* X, const int64_t* y, int n, int IN, int B, int epochs, uint64_t key,
                       void* staging, cudaStream_t stream) {
    int FPC, FP, steps;
    mlp1_stage_bytes(n, IN, B, epochs, &FPC, &FP, &steps);
    if (B > ST_B || IN - FPC > FPC || IN - FPC <= 0) return false;
    const size_t tile_floats = (size_t)ST_B * FP;
    float* xf = static_cast<float*>(staging);
    float* xt = xf + (size_t)steps * 2 * tile_floats;
    int* ys = reinterpret_cast<int*>(xt + (size_t)steps * 2 * tile_floats);
    const size_t smem = (size_t)ST_B * (FP + 1) * 4;
    static size_t configured = 0;
    if (smem > configured) {
        if (cudaFuncSetAttribute(mlp1_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return false;
        configured = smem;
    }
    mlp1_stage_kernel<<<dim3(steps, 2), ST_THREADS, smem, stream>>>(X, y, n, IN, B, epochs, key, FPC, FP, xf, xt, ys);
    return true;
}

// force-load this file's kernels (CUDA loads functions lazily; loading one while another kernel spins
// on a cross-GPU flag could deadlock, so the extension loads everything up front)
void preload_stage() {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, mlp1_stage_kernel);
}

}  // namespace gb
