// murmura_b200 — shared device helpers (sm_100a only).
//
// Every kernel in this extension works on the *flat parameter arena*: one fp32 row ("slot") per
// virtual federated node, laid out as  [ float state (Pf, padded) | int buffers as float (Pi, padded) ].
// Rows of other GPUs are reached through peer-mapped base pointers (cudaIpc / NVLink 5).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mb {

constexpr int kMaxRanks = 16;

// ---- memory access -------------------------------------------------------------------------
// Streaming 128-bit load that does not pollute L1 (neighbour tiles are read exactly once).  Coherent on purpose (no `.nc`): peer
// rows are written while a consumer may already be resident and spinning on the epoch flag, and the acquire on that flag only
// orders the coherent path.
__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// System-scope acquire/release used for the per-rank "published epoch" flags that live in the
// peer-mapped control page (SURVEY §5.8 control plane).
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// ReLU that propagates NaN like torch.relu (fmaxf would launder a NaN — e.g. from an attacked, negative running variance — into 0,
// and a poisoned candidate would then score a finite loss where the reference scores NaN and rejects it).
__device__ __forceinline__ float relu_nan(float x) { return x < 0.f ? 0.f : x; }

// ---- reductions ----------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// Block-wide sum; `scratch` must hold >= 32 floats. Result valid in every thread.
__device__ __forceinline__ float block_sum(float v, float* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    const int nwarps = (blockDim.x + 31) >> 5;
    float r = (threadIdx.x < nwarps) ? scratch[threadIdx.x] : 0.f;
    if (warp == 0) r = warp_sum(r);
    if (threadIdx.x == 0) scratch[0] = r;
    __syncthreads();
    return scratch[0];
}

// ---- Philox4x32-10 counter RNG (in-register noise for the Gaussian attack injector) ---------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0; k.y += W1;
    }
    return c;
}
__device__ __forceinline__ float u32_to_unit(uint32_t x) {       // (0, 1]
    return (float)(x >> 8) * (1.0f / 16777216.0f) + (0.5f / 16777216.0f);
}
// Four standard normals from one Philox block (two Box–Muller pairs).
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint64_t stream, uint64_t idx) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const float r0 = sqrtf(-2.f * __logf(u32_to_unit(r.x))), r1 = sqrtf(-2.f * __logf(u32_to_unit(r.z)));
    float s0, c0, s1, c1;
    __sincosf(6.283185307179586f * u32_to_unit(r.y), &s0, &c0);
    __sincosf(6.283185307179586f * u32_to_unit(r.w), &s1, &c1);
    return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

// ---- edge table ----------------------------------------------------------------------------
// CSR over the destination slots hosted on this GPU. The first entry of every row is the
// destination itself (read from the *live* arena); the others are neighbours, read from the
// published buffer of (rank, slot) — local or peer-mapped.
struct EdgeTable {
    const int* row_ptr;    // [V+1]
    const int* src_rank;   // [E]
    const int* src_slot;   // [E]
    const float* mask;     // [E] 1 = alive, 0 = dropped (fault injection / timed-out peer)
};

struct PeerView {
    const float* const* pub;   // [G] peer-mapped base pointers of the published buffers
    size_t parity_off;         // element offset of the current round's parity plane
    size_t stride;             // elements per slot (P_all)
};

__device__ __forceinline__ const float* edge_src(const PeerView& pv, const EdgeTable& et, int e) {
    return pv.pub[et.src_rank[e]] + pv.parity_off + (size_t)et.src_slot[e] * pv.stride;
}

// Wait until every rank has published `epoch`. One thread per rank spins on the local control
// page; ranks that stay silent past `timeout_cycles` are recorded in *timed_out (bit per rank)
// and treated as missing neighbours (the reference's deadline-driven partial aggregation).
__device__ __forceinline__ void wait_published(const uint32_t* flags, int G, uint32_t epoch,
                                               long long timeout_cycles, uint32_t* timed_out) {
    if (flags != nullptr && G > 1) {
        if (threadIdx.x < G) {
            const long long t0 = clock64();
            while ((int32_t)(ld_acquire_sys(flags + threadIdx.x) - epoch) < 0) {
                if (clock64() - t0 > timeout_cycles) { atomicOr(timed_out, 1u << threadIdx.x); break; }
                __nanosleep(100);
            }
        }
        __syncthreads();
    }
}

}  // namespace mb
