// Shared device helpers for the gossipy_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define GB_DEVICE __device__ __forceinline__

// ---------------------------------------------------------------------------------------------
// counter-based randomness: must stay bit-identical to gossipy_b200/engine/rng.py
// ---------------------------------------------------------------------------------------------
GB_DEVICE uint64_t gb_mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

struct GbPerm {  // keyed permutation of range(n): 4-round Feistel + cycle walking
    uint64_t key; uint32_t n, half, hmask;
    GB_DEVICE void init(uint32_t n_, uint64_t key_) {
        n = n_; key = key_;
        uint32_t bits = (n_ <= 1) ? 2u : (32u - __clz(n_ - 1));
        if (bits < 2) bits = 2;
        bits += bits & 1u;
        half = bits >> 1; hmask = (1u << half) - 1u;
    }
    GB_DEVICE uint32_t operator()(uint32_t i) const {
        uint32_t x = i;
        while (true) {
            uint32_t l = x >> half, r = x & hmask;
#pragma unroll
            for (uint32_t rnd = 0; rnd < 4; ++rnd) {
                uint32_t f = (uint32_t)(gb_mix64(key ^ ((uint64_t)rnd << 56) ^ (uint64_t)r)) & hmask;
                uint32_t nl = r; r = l ^ f; l = nl;
            }
            x = (l << half) | r;
            if (x < n) return x;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// memory helpers
// ---------------------------------------------------------------------------------------------
// Streaming 128-bit load that does not allocate in L1: used for rows that may live in a PEER GPU's
// HBM (mapped through NVLink); such data is read exactly once per merge.  Deliberately NOT `.nc`:
// the row may have been written by another GPU while this kernel was already spinning on its
// `ready` flag, so the load must stay on the coherent path (ordered after the ld.acquire.sys).
GB_DEVICE float4 gb_ld_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
GB_DEVICE float gb_ld_stream1(const float* p) {
    float v;
    asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

GB_DEVICE void gb_cp_async16(void* smem_dst, const void* gmem_src) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gmem_src));
}
GB_DEVICE void gb_cp_async4(void* smem_dst, const void* gmem_src) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(s), "l"(gmem_src));
}
GB_DEVICE void gb_cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> GB_DEVICE void gb_cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N)); }

// cross-GPU flags (system scope) for the multi-process transport
GB_DEVICE void gb_st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
GB_DEVICE void gb_red_release_sys_add(uint32_t* p, uint32_t v) {
    asm volatile("red.release.sys.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
GB_DEVICE uint32_t gb_ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Bounded wait for a row's `ready` generation (cross-GPU handshake).  A lost publish must not hang the
// GPU forever: after ~30 s the waiter records fault bit 1 in `fault` (host-visible, checked with the
// metrics of every round) and carries on with whatever the row holds.
GB_DEVICE uint64_t gb_globaltimer() { uint64_t t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
constexpr uint64_t GB_WAIT_NS = 30ull * 1000ull * 1000ull * 1000ull;
GB_DEVICE void gb_wait_flag(const uint32_t* flag, uint32_t gen, uint32_t* fault, uint32_t code = 1u) {
    if ((int32_t)(gb_ld_acquire_sys(flag) - gen) >= 0) return;
    const uint64_t t0 = gb_globaltimer();
    unsigned it = 0;
    while ((int32_t)(gb_ld_acquire_sys(flag) - gen) < 0) {
        __nanosleep(40);
        if ((++it & 1023u) == 0u && gb_globaltimer() - t0 > GB_WAIT_NS) {
            if (fault != nullptr) atomicOr(fault, code);
            return;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// thread-block cluster helpers (raw PTX, no cooperative_groups dependency)
// ---------------------------------------------------------------------------------------------
GB_DEVICE uint32_t gb_cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
GB_DEVICE uint32_t gb_cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
GB_DEVICE void gb_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
GB_DEVICE void gb_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
GB_DEVICE void gb_cluster_sync() { gb_cluster_arrive(); gb_cluster_wait(); }
// address of `smem_ptr` (a pointer into MY shared memory) inside CTA `rank` of the cluster
GB_DEVICE uint32_t gb_map_shared(const void* smem_ptr, uint32_t rank) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_ptr), r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
GB_DEVICE void gb_st_cluster(uint32_t addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" :: "r"(addr), "f"(v) : "memory");
}
GB_DEVICE void gb_st_cluster4(uint32_t addr, float4 v) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

GB_DEVICE float gb_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
