// murmura_b200 — fused neighbour-exchange + aggregation kernels (sm_100a).
//
// Reference call sites replaced: K1 exchange (murmura/core/network.py:105-139,
// murmura/distributed/node_process.py:221-276), K2 FedAvg (aggregation/base.py:76-115), K3 Krum
// (aggregation/krum.py:55-75), K4 BALANCE (aggregation/balance.py:82-175), K5 Sketchguard
// (aggregation/sketchguard.py:71-261), K6 UBAR (aggregation/ubar.py:101-249), K7 EvidentialTrust
// (aggregation/evidential_trust.py:177-212,289-381), K9/K10 attack injectors (attacks/gaussian.py:79-90,
// attacks/directed.py:79-89).  There is no send kernel and no NCCL call: a round is
//   publish  (copy live → published[parity], attack fused, release the epoch flag to every peer)
//   distance / sketch / filter kernels (tiny; wait on the flags, read peers' tiles over NVLink)
//   weighted_gather (streams the accepted neighbours' tiles from peer memory, writes live in place)
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda_fp8.h>
#include "common.cuh"

namespace mb {

constexpr int kThreads = 256;
constexpr int kMaxRow = 128;          // max (degree + 1) per destination row kept in shared memory

// =============================================================================================
// publish: live → published[parity] with the attack fused in, int buffers mirrored as floats,
// then (last block) release-store the epoch into every peer's control page.
// =============================================================================================
struct PublishArgs {
    float* live;                 // [V][stride]
    float* pub;                  // local published plane for this parity: [S][stride]
    size_t stride;
    int Pf;                      // number of real float elements (<= Pf_pad)
    int Pf_pad;                  // float region length (multiple of 4), tail starts here
    const long long* ints;       // [V][n_int] live int buffers (may be null)
    int n_int;
    const float* scale;          // [V] multiplicative attack (1 for honest, lambda for directed deviation)
    const float* noise_std;      // [V] additive Gaussian std (0 for honest)
    const int* node_gid;         // [V] global node ids (Philox stream)
    unsigned long long seed;
    unsigned long long round;
    uint32_t* const* peer_flags; // [G] control-page flag arrays of every rank (may be null when G == 1)
    int G;
    int my_rank;
    uint32_t epoch;
    unsigned int* ticket;        // device counter, zero on entry, reset by the last block
};

__global__ void __launch_bounds__(kThreads) publish_kernel(PublishArgs a) {
    const int v = blockIdx.y;
    const float4* src = reinterpret_cast<const float4*>(a.live + (size_t)v * a.stride);
    float4* dst = reinterpret_cast<float4*>(a.pub + (size_t)v * a.stride);
    const int n4 = a.Pf_pad >> 2;
    const float sc = a.scale[v], sd = a.noise_std[v];
    const uint64_t stream = a.round * 1000003ull + (uint64_t)a.node_gid[v];
    auto transform = [&](float4 x, int i) {
        if (sc != 1.f) { x.x *= sc; x.y *= sc; x.z *= sc; x.w *= sc; }
        if (sd != 0.f) {
            const float4 z = philox_normal4(a.seed, stream, (uint64_t)i);
            const int base = i << 2;                       // keep the zero padding exactly zero
            if (base + 0 < a.Pf) x.x = fmaf(sd, z.x, x.x);
            if (base + 1 < a.Pf) x.y = fmaf(sd, z.y, x.y);
            if (base + 2 < a.Pf) x.z = fmaf(sd, z.z, x.z);
            if (base + 3 < a.Pf) x.w = fmaf(sd, z.w, x.w);
        }
        return x;
    };
    constexpr int U = 4;                                   // independent 128-bit loads in flight per thread
    const int step = gridDim.x * blockDim.x;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * step < n4; i += U * step) {
        float4 x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) x[u] = ld_stream(src + i + u * step);
#pragma unroll
        for (int u = 0; u < U; ++u) st_stream(dst + i + u * step, transform(x[u], i + u * step));
    }
    for (; i < n4; i += step) st_stream(dst + i, transform(src[i], i));
    if (blockIdx.x == 0 && a.n_int > 0) {                  // int buffers ride along as floats (never attacked)
        float* lt = a.live + (size_t)v * a.stride + a.Pf_pad;
        float* pt = a.pub + (size_t)v * a.stride + a.Pf_pad;
        for (int k = threadIdx.x; k < a.n_int; k += blockDim.x) {
            const float f = (float)a.ints[(size_t)v * a.n_int + k];
            lt[k] = f; pt[k] = f;
        }
    }
    // ---- epoch flag: the last block to finish publishes to every rank --------------------------
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int total = gridDim.x * gridDim.y;
        if (atomicAdd(a.ticket, 1u) == total - 1) {
            __threadfence_system();
            *a.ticket = 0;
            if (a.peer_flags != nullptr)
                for (int g = 0; g < a.G; ++g) st_release_sys(a.peer_flags[g] + a.my_rank, a.epoch);
        }
    }
}

// =============================================================================================
// publish + per-rank column sum (full-mesh FedAvg): the same pass that writes the published rows also writes
// rsum = Σ_v published_v, so the aggregate needs ONE row per rank from the fabric instead of every node's row.
// One thread owns a float4 column of all V local rows (V is small); bytes = publish bytes + one extra row written.
// =============================================================================================
__global__ void __launch_bounds__(kThreads) publish_sum_kernel(PublishArgs a, float* __restrict__ rsum, int V) {
    const int n4 = a.Pf_pad >> 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
        int v = 0;
        auto one = [&](float4 x, int vv) {
            const float sc = a.scale[vv], sd = a.noise_std[vv];
            if (sc != 1.f) { x.x *= sc; x.y *= sc; x.z *= sc; x.w *= sc; }
            if (sd != 0.f) {
                const float4 z = philox_normal4(a.seed, a.round * 1000003ull + (uint64_t)a.node_gid[vv], (uint64_t)i);
                const int base = i << 2;
                if (base + 0 < a.Pf) x.x = fmaf(sd, z.x, x.x);
                if (base + 1 < a.Pf) x.y = fmaf(sd, z.y, x.y);
                if (base + 2 < a.Pf) x.z = fmaf(sd, z.z, x.z);
                if (base + 3 < a.Pf) x.w = fmaf(sd, z.w, x.w);
            }
            st_stream(reinterpret_cast<float4*>(a.pub + (size_t)vv * a.stride) + i, x);
            sum.x += x.x; sum.y += x.y; sum.z += x.z; sum.w += x.w;
        };
        for (; v + 4 <= V; v += 4) {                       // 4 independent 128-bit loads in flight
            float4 x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) x[u] = ld_stream(reinterpret_cast<const float4*>(a.live + (size_t)(v + u) * a.stride) + i);
#pragma unroll
            for (int u = 0; u < 4; ++u) one(x[u], v + u);
        }
        for (; v < V; ++v) one(ld_stream(reinterpret_cast<const float4*>(a.live + (size_t)v * a.stride) + i), v);
        st_stream(reinterpret_cast<float4*>(rsum) + i, sum);
    }
    if (blockIdx.x == 0 && a.n_int > 0) {
        for (int v = 0; v < V; ++v) {
            float* lt = a.live + (size_t)v * a.stride + a.Pf_pad;
            float* pt = a.pub + (size_t)v * a.stride + a.Pf_pad;
            for (int k = threadIdx.x; k < a.n_int; k += blockDim.x) {
                const float f = (float)a.ints[(size_t)v * a.n_int + k];
                lt[k] = f; pt[k] = f;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        if (atomicAdd(a.ticket, 1u) == gridDim.x - 1) {
            __threadfence_system();
            *a.ticket = 0;
            if (a.peer_flags != nullptr)
                for (int g = 0; g < a.G; ++g) st_release_sys(a.peer_flags[g] + a.my_rank, a.epoch);
        }
    }
}

// =============================================================================================
// weighted_gather: out_v = Σ_e w_e · θ_src(e)   (self edge reads the live row; in-place safe)
// =============================================================================================
struct GatherArgs {
    float* live;
    PeerView pv;
    EdgeTable et;
    const float* w;              // [E]
    int len4;                    // float4s per row to process (float region)
    int renorm;                  // divide by Σ of surviving weights (FedAvg mean over present neighbours)
    const uint32_t* flags; int G; uint32_t epoch; long long timeout; uint32_t* timed_out;
};

__global__ void __launch_bounds__(kThreads) weighted_gather_kernel(GatherArgs a) {
    __shared__ const float4* s_src[kMaxRow];
    __shared__ float s_w[kMaxRow];
    __shared__ int s_n;
    __shared__ float s_wself;
    wait_published(a.flags, a.G, a.epoch, a.timeout, a.timed_out);
    const int v = blockIdx.y;
    const int e0 = a.et.row_ptr[v], e1 = a.et.row_ptr[v + 1];
    float4* out = reinterpret_cast<float4*>(a.live + (size_t)v * a.pv.stride);
    if (threadIdx.x == 0) {
        const uint32_t dead = a.timed_out ? *a.timed_out : 0u;
        int n = 0;
        float wself = a.w[e0] * a.et.mask[e0], tot = wself;
        for (int e = e0 + 1; e < e1; ++e) {
            float w = a.w[e] * a.et.mask[e];
            if ((dead >> a.et.src_rank[e]) & 1u) w = 0.f;
            if (w == 0.f) continue;
            s_src[n] = reinterpret_cast<const float4*>(edge_src(a.pv, a.et, e));
            s_w[n] = w; tot += w; ++n;
        }
        if (a.renorm && tot != 0.f) { wself /= tot; for (int k = 0; k < n; ++k) s_w[k] /= tot; }
        s_n = n; s_wself = wself;
    }
    __syncthreads();
    const int n = s_n;
    const float wself = s_wself;
    if (n == 0 && wself == 1.f) return;                      // keep own state untouched
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.len4; i += gridDim.x * blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (wself != 0.f) {                                   // own row: plain (coherent) load, written back in place below
            const float4 o = out[i];
            acc = make_float4(wself * o.x, wself * o.y, wself * o.z, wself * o.w);
        }
        int k = 0;
        for (; k + 4 <= n; k += 4) {            // 4 independent 128-bit (possibly NVLink) loads in flight per thread
            const float4 x0 = ld_stream(s_src[k] + i), x1 = ld_stream(s_src[k + 1] + i);
            const float4 x2 = ld_stream(s_src[k + 2] + i), x3 = ld_stream(s_src[k + 3] + i);
            const float w0 = s_w[k], w1 = s_w[k + 1], w2 = s_w[k + 2], w3 = s_w[k + 3];
            acc.x = fmaf(w0, x0.x, acc.x); acc.y = fmaf(w0, x0.y, acc.y); acc.z = fmaf(w0, x0.z, acc.z); acc.w = fmaf(w0, x0.w, acc.w);
            acc.x = fmaf(w1, x1.x, acc.x); acc.y = fmaf(w1, x1.y, acc.y); acc.z = fmaf(w1, x1.z, acc.z); acc.w = fmaf(w1, x1.w, acc.w);
            acc.x = fmaf(w2, x2.x, acc.x); acc.y = fmaf(w2, x2.y, acc.y); acc.z = fmaf(w2, x2.z, acc.z); acc.w = fmaf(w2, x2.w, acc.w);
            acc.x = fmaf(w3, x3.x, acc.x); acc.y = fmaf(w3, x3.y, acc.y); acc.z = fmaf(w3, x3.z, acc.z); acc.w = fmaf(w3, x3.w, acc.w);
        }
        for (; k < n; ++k) {
            const float4 x = ld_stream(s_src[k] + i);
            const float w = s_w[k];
            acc.x = fmaf(w, x.x, acc.x); acc.y = fmaf(w, x.y, acc.y); acc.z = fmaf(w, x.z, acc.z); acc.w = fmaf(w, x.w, acc.w);
        }
        out[i] = acc;
    }
}

// =============================================================================================
// NVLS full-mesh FedAvg: out = (1/N)·Σ_{all nodes} θ   with the cross-GPU sum done INSIDE the NVSwitch.
// The published planes of all ranks are bound to one multicast object; `multimem.ld_reduce.add.v4.f32` on the multicast
// address returns Σ_ranks of the word at that offset, so each GPU ingests S·P floats instead of (N - V)·P.
// A Byzantine destination's own term is its live row, not its (attacked) published row: corrected locally.
// =============================================================================================
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc_addr) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc_addr) : "memory");
    return v;
}

struct NvlsArgs {
    float* live;                 // [V][stride] local live rows (written)
    const float* pub_local;      // local published plane for this parity [S][stride]
    const float* mc_pub;         // multicast address of the same plane
    size_t stride;
    int V, S, len4;
    float inv_n;
    const uint8_t* byz;          // [V] 1 = destination's own term must be its live row
    const uint32_t* flags; int G; uint32_t epoch; long long timeout; uint32_t* timed_out;
};

__global__ void __launch_bounds__(kThreads) nvls_fedavg_kernel(NvlsArgs a) {
    wait_published(a.flags, a.G, a.epoch, a.timeout, a.timed_out);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.len4; i += gridDim.x * blockDim.x) {
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < a.S; ++s) {
            const float4 r = multimem_ld_reduce_add(a.mc_pub + (size_t)s * a.stride + ((size_t)i << 2));
            sum.x += r.x; sum.y += r.y; sum.z += r.z; sum.w += r.w;
        }
        for (int v = 0; v < a.V; ++v) {
            float4 o = sum;
            float4* dst = reinterpret_cast<float4*>(a.live + (size_t)v * a.stride) + i;
            if (a.byz[v]) {
                const float4 p = reinterpret_cast<const float4*>(a.pub_local + (size_t)v * a.stride)[i];
                const float4 l = *dst;
                o.x += l.x - p.x; o.y += l.y - p.y; o.z += l.z - p.z; o.w += l.w - p.w;
            }
            o.x *= a.inv_n; o.y *= a.inv_n; o.z *= a.inv_n; o.w *= a.inv_n;
            *dst = o;
        }
    }
}

// Two-shot all-reduce of the per-rank sum rows, written as ONE kernel over peer memory: this GPU reduces slice
// [len·r/G, len·(r+1)/G) of the rsum rows of all ranks (one `multimem.ld_reduce` per 16 bytes when the arena is bound to an NVLS
// multicast object — the switch adds; G peer loads otherwise) and immediately scatters the reduced values into the `total` row of
// EVERY rank (`multimem.st`, or G peer stores), then raises its epoch flag (last block).  Per GPU and direction the fabric carries
// one row instead of G−1 (peer loads) or G (every GPU asking the switch for the whole row).  Liveness is the frozen mask; with a
// missing rank the kernel does nothing and `fedavg_fullmesh` falls back to the one-shot reduction over the ranks that arrived.
__device__ __forceinline__ void multimem_st_v4(float* mc_addr, const float4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(mc_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

struct ReduceScatterArgs {
    const float* const* peer_rsum; const float* mc_rsum; float* const* peer_tot; float* mc_tot;
    int len4, G, my_rank; const uint32_t* timed_out;
    uint32_t* const* peer_flags; uint32_t epoch; unsigned int* ticket;
};

__global__ void __launch_bounds__(kThreads) fullmesh_reduce_scatter_kernel(ReduceScatterArgs a) {
    const uint32_t dead = a.timed_out ? *a.timed_out : 0u;
    if (dead == 0u) {
        const int lo = (int)((long long)a.len4 * a.my_rank / a.G), hi = (int)((long long)a.len4 * (a.my_rank + 1) / a.G);
        for (int i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x) {
            float4 sum;
            if (a.mc_rsum) sum = multimem_ld_reduce_add(a.mc_rsum + ((size_t)i << 2));
            else {
                sum = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int r = 0; r < a.G; ++r) {
                    const float4 x = ld_stream(reinterpret_cast<const float4*>(a.peer_rsum[r]) + i);
                    sum.x += x.x; sum.y += x.y; sum.z += x.z; sum.w += x.w;
                }
            }
            if (a.mc_tot) multimem_st_v4(a.mc_tot + ((size_t)i << 2), sum);
            else for (int r = 0; r < a.G; ++r) st_stream(reinterpret_cast<float4*>(a.peer_tot[r]) + i, sum);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        if (atomicAdd(a.ticket, 1u) == gridDim.x - 1) {
            __threadfence_system();
            *a.ticket = 0;
            for (int g = 0; g < a.G; ++g) st_release_sys(a.peer_flags[g] + a.my_rank, a.epoch);
        }
    }
}

// Full-mesh FedAvg on the per-rank sums: T = Σ_ranks rsum_r (one `multimem.ld_reduce` when the arena is bound to an NVLS
// multicast object and every rank arrived, otherwise G−1 peer loads), then every local node gets T/N — a Byzantine
// destination swaps its own (attacked) published term for its live row.  Bytes per GPU: one row in over the fabric,
// V rows written: the minimal-byte schedule ("all destinations share the source set").  Liveness is the frozen mask.
struct FullMeshArgs {
    float* live; const float* pub_local; const float* const* peer_rsum; const float* mc_rsum; const float* tot_local;
    size_t stride; int V, len4, G, N; const uint8_t* byz; const int* rank_nodes; const uint32_t* timed_out;
};

__global__ void __launch_bounds__(kThreads) fedavg_fullmesh_kernel(FullMeshArgs a) {
    const uint32_t dead = a.timed_out ? *a.timed_out : 0u;
    int n_alive = a.N;
    if (dead) for (int r = 0; r < a.G; ++r) if ((dead >> r) & 1u) n_alive -= a.rank_nodes[r];
    const float inv = 1.f / (float)max(n_alive, 1);
    const bool use_mc = a.mc_rsum != nullptr && dead == 0u && a.G > 1;
    const bool use_tot = a.tot_local != nullptr && dead == 0u;           // two-shot: the reduced row is already in local HBM
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.len4; i += gridDim.x * blockDim.x) {
        float4 sum;
        if (use_tot) sum = ld_stream(reinterpret_cast<const float4*>(a.tot_local) + i);
        else if (use_mc) sum = multimem_ld_reduce_add(a.mc_rsum + ((size_t)i << 2));
        else {
            sum = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = 0; r < a.G; ++r) {
                if ((dead >> r) & 1u) continue;
                const float4 x = ld_stream(reinterpret_cast<const float4*>(a.peer_rsum[r]) + i);
                sum.x += x.x; sum.y += x.y; sum.z += x.z; sum.w += x.w;
            }
        }
        for (int v = 0; v < a.V; ++v) {
            float4 o = sum;
            float4* dst = reinterpret_cast<float4*>(a.live + (size_t)v * a.stride) + i;
            if (a.byz[v]) {
                const float4 p = reinterpret_cast<const float4*>(a.pub_local + (size_t)v * a.stride)[i];
                const float4 l = *dst;
                o.x += l.x - p.x; o.y += l.y - p.y; o.z += l.z - p.z; o.w += l.w - p.w;
            }
            o.x *= inv; o.y *= inv; o.z *= inv; o.w *= inv;
            st_stream(dst, o);
        }
    }
}

// Stream-ordered wait for all ranks' epoch flags (used before library / TMA kernels that cannot spin themselves).
__global__ void wait_epoch_kernel(const uint32_t* flags, int G, uint32_t epoch, long long timeout, uint32_t* timed_out) {
    wait_published(flags, G, epoch, timeout, timed_out);
}

// =============================================================================================
// weighted_gather, TMA-bulk variant: neighbour tiles are pulled by the copy engine (`cp.async.bulk` global → shared,
// completion on an mbarrier; SASS: UBLKCP) through an 8-stage ring, so the (NVLink) loads of the next units are in flight
// while the CTA folds the current one — no registers are tied up by outstanding loads.  Unit = (tile of 2048 floats, source).
// =============================================================================================
constexpr int kBulkTile = 2048;                 // floats per unit (8 KiB)
constexpr int kBulkStages = 8;
constexpr int kBulkThreads = 128;               // 16 floats per thread per unit

__device__ __forceinline__ uint32_t bulk_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bulk_smem_u32(bar)), "r"(parity) : "memory");
        if (done) break;
        if (++spins > (1u << 28)) __trap();
    }
}

__global__ void __launch_bounds__(kBulkThreads) weighted_gather_bulk_kernel(GatherArgs a) {
    extern __shared__ __align__(128) uint8_t bulk_smem[];
    __shared__ const float* s_src[kMaxRow];
    __shared__ float s_w[kMaxRow];
    __shared__ int s_n;
    __shared__ __align__(8) uint64_t full_bar[kBulkStages];
    __shared__ __align__(8) uint64_t empty_bar[kBulkStages];
    wait_published(a.flags, a.G, a.epoch, a.timeout, a.timed_out);
    const int v = blockIdx.y;
    const int e0 = a.et.row_ptr[v], e1 = a.et.row_ptr[v + 1];
    float* out = a.live + (size_t)v * a.pv.stride;
    if (threadIdx.x == 0) {
        const uint32_t dead = a.timed_out ? *a.timed_out : 0u;
        int n = 0; float tot = 0.f;
        for (int e = e0; e < e1; ++e) {
            float w = a.w[e] * a.et.mask[e];
            if (e != e0 && ((dead >> a.et.src_rank[e]) & 1u)) w = 0.f;
            if (w == 0.f) continue;
            s_src[n] = (e == e0) ? out : edge_src(a.pv, a.et, e);
            s_w[n] = w; tot += w; ++n;
        }
        if (a.renorm && tot != 0.f) for (int k = 0; k < n; ++k) s_w[k] /= tot;
        s_n = n;
        for (int s = 0; s < kBulkStages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bulk_smem_u32(&full_bar[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bulk_smem_u32(&empty_bar[s])), "r"(kBulkThreads));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");               // rows were produced / acquired through the generic proxy
    }
    __syncthreads();
    const int n = s_n;
    if (n == 0) return;
    if (n == 1 && s_src[0] == out && s_w[0] == 1.f) return;           // keep own state untouched
    const int len = a.len4 << 2;
    const int ntiles = (len + kBulkTile - 1) / kBulkTile;
    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;      // tiles blockIdx.x, +gridDim.x, …
    const long long units = (long long)my_tiles * n;
    auto issue = [&](long long u) {                                     // thread 0 only
        const int stage = (int)(u % kBulkStages);
        const uint32_t use = (uint32_t)(u / kBulkStages);
        if (use > 0) bulk_wait(&empty_bar[stage], (use - 1) & 1u);      // all 128 readers released the slot
        const int tile = (int)blockIdx.x + (int)(u / n) * (int)gridDim.x;
        const int k = (int)(u % n);
        const int off = tile * kBulkTile;
        const uint32_t bytes = (uint32_t)(min(kBulkTile, len - off) * 4);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bulk_smem_u32(&full_bar[stage])), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(bulk_smem_u32(bulk_smem + stage * kBulkTile * 4)), "l"(s_src[k] + off), "r"(bytes),
                        "r"(bulk_smem_u32(&full_bar[stage])) : "memory");
    };
    if (threadIdx.x == 0)
        for (long long u = 0; u < units && u < kBulkStages - 1; ++u) issue(u);
    float4 acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long u = 0; u < units; ++u) {
        if (threadIdx.x == 0 && u + kBulkStages - 1 < units) issue(u + kBulkStages - 1);
        const int stage = (int)(u % kBulkStages);
        bulk_wait(&full_bar[stage], (uint32_t)((u / kBulkStages) & 1));
        const int k = (int)(u % n);
        const float w = s_w[k];
        const float4* t = reinterpret_cast<const float4*>(bulk_smem + stage * kBulkTile * 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) {                                   // thread owns float4 #(j*128 + tid) of the tile: conflict-free
            const float4 x = t[j * kBulkThreads + threadIdx.x];
            acc[j].x = fmaf(w, x.x, acc[j].x); acc[j].y = fmaf(w, x.y, acc[j].y); acc[j].z = fmaf(w, x.z, acc[j].z); acc[j].w = fmaf(w, x.w, acc[j].w);
        }
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bulk_smem_u32(&empty_bar[stage])) : "memory");
        if (k == n - 1) {                                               // tile complete → write back in place
            const int tile = (int)blockIdx.x + (int)(u / n) * (int)gridDim.x;
            const int off4 = tile * (kBulkTile / 4);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i4 = off4 + j * kBulkThreads + threadIdx.x;
                if (i4 < a.len4) reinterpret_cast<float4*>(out)[i4] = acc[j];
                acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
}

// int-buffer tail: ints_v = trunc(Σ_e w_tail_e · tail_src(e))  (per-aggregator rules)
__global__ void tail_blend_kernel(float* live, PeerView pv, EdgeTable et, const float* w_tail, int Pf_pad,
                                  long long* ints, int n_int, const uint32_t* timed_out) {
    const int v = blockIdx.x;
    const int e0 = et.row_ptr[v], e1 = et.row_ptr[v + 1];
    const uint32_t dead = timed_out ? *timed_out : 0u;
    for (int k = threadIdx.x; k < n_int; k += blockDim.x) {
        float acc = 0.f;
        for (int e = e0; e < e1; ++e) {
            float w = w_tail[e] * et.mask[e];
            if (e != e0 && ((dead >> et.src_rank[e]) & 1u)) w = 0.f;
            if (w == 0.f) continue;
            const float* src = (e == e0) ? live + (size_t)v * pv.stride : edge_src(pv, et, e);
            acc = fmaf(w, src[Pf_pad + k], acc);
        }
        ints[(size_t)v * n_int + k] = (long long)truncf(acc);
    }
}

// =============================================================================================
// edge_distances: d2[e] = Σ (own - θ_src(e))², n2[v] = Σ own²  over the first len4 float4s
// =============================================================================================
struct DistArgs {
    const float* live;
    PeerView pv;
    EdgeTable et;
    int len4;
    float* d2;                   // [E] (zeroed by caller)
    float* n2;                   // [V] (zeroed by caller)
    const uint32_t* flags; int G; uint32_t epoch; long long timeout; uint32_t* timed_out;
};

__global__ void __launch_bounds__(kThreads) edge_distances_kernel(DistArgs a) {
    __shared__ const float4* s_src[kMaxRow];
    __shared__ float s_acc[kMaxRow];
    __shared__ int s_eid[kMaxRow];
    __shared__ int s_n;
    __shared__ float s_scratch[32];
    wait_published(a.flags, a.G, a.epoch, a.timeout, a.timed_out);
    const int v = blockIdx.y;
    const int e0 = a.et.row_ptr[v], e1 = a.et.row_ptr[v + 1];
    const float4* own = reinterpret_cast<const float4*>(a.live + (size_t)v * a.pv.stride);
    if (threadIdx.x == 0) {
        const uint32_t dead = a.timed_out ? *a.timed_out : 0u;
        int n = 0;
        for (int e = e0 + 1; e < e1; ++e) {
            if (a.et.mask[e] == 0.f || ((dead >> a.et.src_rank[e]) & 1u)) continue;
            s_src[n] = reinterpret_cast<const float4*>(edge_src(a.pv, a.et, e));
            s_eid[n] = e; s_acc[n] = 0.f; ++n;
        }
        s_n = n;
    }
    __syncthreads();
    const int n = s_n;
    const int lane = threadIdx.x & 31;
    float own_sq = 0.f;
    constexpr int U = 4;                          // float4s per thread per tile
    const int tile = blockDim.x * U;
    for (int base = blockIdx.x * tile; base < a.len4; base += gridDim.x * tile) {
        float4 o[U]; bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = base + u * blockDim.x + threadIdx.x;
            ok[u] = i < a.len4;
            o[u] = ok[u] ? own[i] : make_float4(0.f, 0.f, 0.f, 0.f);
            own_sq += o[u].x * o[u].x + o[u].y * o[u].y + o[u].z * o[u].z + o[u].w * o[u].w;
        }
        for (int k = 0; k < n; ++k) {
            float4 x[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                x[u] = ok[u] ? ld_stream(s_src[k] + base + u * blockDim.x + threadIdx.x) : o[u];
            float p = 0.f;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float dx = o[u].x - x[u].x, dy = o[u].y - x[u].y, dz = o[u].z - x[u].z, dw = o[u].w - x[u].w;
                p += dx * dx + dy * dy + dz * dz + dw * dw;
            }
            p = warp_sum(p);
            if (lane == 0) atomicAdd(&s_acc[k], p);
        }
    }
    own_sq = block_sum(own_sq, s_scratch);
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(&a.n2[v], own_sq);
    for (int k = threadIdx.x; k < n; k += blockDim.x) atomicAdd(&a.d2[s_eid[k]], s_acc[k]);
}

// =============================================================================================
// pairwise (exact fp32, shared-memory staged): D[v][i][j] = Σ (c_i - c_j)² over the float region
// for the m = deg+1 candidates of destination v.  Small-m / fallback path of Krum; the large-m
// path is the tcgen05 Gram kernel (gram_tcgen05.cu).
// =============================================================================================
constexpr int kPairM = 32;
constexpr int kPairTile = 256;       // floats per candidate per tile

struct PairArgs {
    const float* live; PeerView pv; EdgeTable et; int len;   // len = floats (multiple of 4)
    float* D;                                                // [V][kPairM][kPairM], zeroed
    const uint32_t* flags; int G; uint32_t epoch; long long timeout; uint32_t* timed_out;
};

__global__ void __launch_bounds__(kThreads) pairwise_kernel(PairArgs a) {
    extern __shared__ float s_tile[];                        // [m][kPairTile]
    __shared__ const float* s_src[kPairM];
    __shared__ float s_acc[kPairM * kPairM];
    wait_published(a.flags, a.G, a.epoch, a.timeout, a.timed_out);
    const int v = blockIdx.y;
    const int e0 = a.et.row_ptr[v];
    const int m = min(a.et.row_ptr[v + 1] - e0, kPairM);
    for (int k = threadIdx.x; k < m; k += blockDim.x)
        s_src[k] = (k == 0) ? a.live + (size_t)v * a.pv.stride : edge_src(a.pv, a.et, e0 + k);
    for (int k = threadIdx.x; k < kPairM * kPairM; k += blockDim.x) s_acc[k] = 0.f;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int npairs = m * (m - 1) / 2;
    for (int base = blockIdx.x * kPairTile; base < a.len; base += gridDim.x * kPairTile) {
        const int cnt4 = min(kPairTile, a.len - base) >> 2;
        for (int idx = threadIdx.x; idx < m * (kPairTile / 4); idx += blockDim.x) {
            const int c = idx / (kPairTile / 4), q = idx % (kPairTile / 4);
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q < cnt4) x = ld_stream(reinterpret_cast<const float4*>(s_src[c] + base) + q);
            reinterpret_cast<float4*>(s_tile + c * kPairTile)[q] = x;
        }
        __syncthreads();
        for (int p = warp; p < npairs; p += nwarps) {
            int i = 0, rem = p;                              // unrank p → (i < j)
            while (rem >= m - 1 - i) { rem -= m - 1 - i; ++i; }
            const int j = i + 1 + rem;
            float s = 0.f;
#pragma unroll
            for (int t = 0; t < kPairTile / 32; ++t) {
                const float d = s_tile[i * kPairTile + t * 32 + lane] - s_tile[j * kPairTile + t * 32 + lane];
                s = fmaf(d, d, s);
            }
            s = warp_sum(s);
            if (lane == 0) s_acc[i * kPairM + j] += s;       // pair p is owned by exactly one warp
        }
        __syncthreads();
    }
    float* Dv = a.D + (size_t)v * kPairM * kPairM;
    for (int p = threadIdx.x; p < kPairM * kPairM; p += blockDim.x) {
        const int i = p / kPairM, j = p % kPairM;
        if (i < j && j < m) { atomicAdd(&Dv[i * kPairM + j], s_acc[p]); atomicAdd(&Dv[j * kPairM + i], s_acc[p]); }
    }
}

// ---- exact refinement of Gram-derived distances ---------------------------------------------------------------------------
// D = G_aa + G_bb − 2·G_ab from a TF32 Gram carries an absolute error ∝ (‖a‖² + ‖b‖²); once honest models have converged the
// true ‖a − b‖² is far below it (cancellation).  Every pair with D < τ·(‖a‖² + ‖b‖²) is therefore recomputed exactly
// (Σ (a_k − b_k)² in fp32, the reference's definition: aggregation/base.py:118-135) — the fp32 fallback of SURVEY §7.3-2.
// Pass 1: (pair, chunk) CTAs of flagged pairs stream both rows and add their partial sums into `scratch`; pass 2 substitutes.
__device__ __forceinline__ void unrank_pair(int p, int& i, int& j) {
    i = 0; int rem = p;
    while (rem >= kPairM - 1 - i) { rem -= kPairM - 1 - i; ++i; }
    j = i + 1 + rem;
}

__global__ void __launch_bounds__(kThreads) krum_refine_kernel(PairArgs a, const float* __restrict__ norms, float tau,
                                                               float* __restrict__ scratch, int chunks) {
    __shared__ float red[32];
    const int v = blockIdx.y;
    const int p = blockIdx.x / chunks, chunk = blockIdx.x - p * chunks;
    int i, j; unrank_pair(p, i, j);
    const int e0 = a.et.row_ptr[v];
    const int m = min(a.et.row_ptr[v + 1] - e0, kPairM);
    if (j >= m) return;
    const float d = a.D[((size_t)v * kPairM + i) * kPairM + j];
    if (!(d < tau * (norms[v * kPairM + i] + norms[v * kPairM + j]))) return;
    const float* xi = (i == 0) ? a.live + (size_t)v * a.pv.stride : edge_src(a.pv, a.et, e0 + i);
    const float* xj = edge_src(a.pv, a.et, e0 + j);
    const int n4 = a.len >> 2;
    const int per = (n4 + chunks - 1) / chunks;
    const int lo = chunk * per, hi = min(n4, lo + per);
    float s = 0.f;
    for (int q = lo + threadIdx.x; q < hi; q += blockDim.x) {
        const float4 x = ld_stream(reinterpret_cast<const float4*>(xi) + q), y = ld_stream(reinterpret_cast<const float4*>(xj) + q);
        const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
        s = fmaf(d0, d0, s); s = fmaf(d1, d1, s); s = fmaf(d2, d2, s); s = fmaf(d3, d3, s);
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) atomicAdd(&scratch[((size_t)v * kPairM + i) * kPairM + j], s);
}

__global__ void krum_refine_apply_kernel(EdgeTable et, float* __restrict__ D, const float* __restrict__ norms, const float* __restrict__ scratch,
                                         float tau, float* __restrict__ nref) {
    const int v = blockIdx.x;
    const int m = min(et.row_ptr[v + 1] - et.row_ptr[v], kPairM);
    int cnt = 0;
    for (int p = threadIdx.x; p < kPairM * (kPairM - 1) / 2; p += blockDim.x) {
        int i, j; unrank_pair(p, i, j);
        if (j >= m) continue;
        float* dij = D + ((size_t)v * kPairM + i) * kPairM + j;
        if (*dij < tau * (norms[v * kPairM + i] + norms[v * kPairM + j])) {
            const float e = scratch[((size_t)v * kPairM + i) * kPairM + j];
            *dij = e; D[((size_t)v * kPairM + j) * kPairM + i] = e;
            ++cnt;
        }
    }
    if (nref && cnt) atomicAdd(nref + v, (float)cnt);
}

// =============================================================================================
// Count-Sketch: s[h(k)] += σ(k)·θ[k]; table packed as uint16 = bucket | (sign<0)<<15
// =============================================================================================
__global__ void __launch_bounds__(kThreads) count_sketch_kernel(const float* base, size_t stride, const int* slots,
                                                                 const uint16_t* table, int Pf, int K, float* out) {
    extern __shared__ float s_hist[];
    const int row = blockIdx.y;
    const float* src = base + (size_t)slots[row] * stride;
    for (int k = threadIdx.x; k < K; k += blockDim.x) s_hist[k] = 0.f;
    __syncthreads();
    const int n4 = Pf >> 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        const float4 x = ld_stream(reinterpret_cast<const float4*>(src) + i);
        const uint2 t = reinterpret_cast<const uint2*>(table)[i];
        const uint32_t t0 = t.x & 0xffffu, t1 = t.x >> 16, t2 = t.y & 0xffffu, t3 = t.y >> 16;
        atomicAdd(&s_hist[t0 & 0x7fffu], (t0 & 0x8000u) ? -x.x : x.x);
        atomicAdd(&s_hist[t1 & 0x7fffu], (t1 & 0x8000u) ? -x.y : x.y);
        atomicAdd(&s_hist[t2 & 0x7fffu], (t2 & 0x8000u) ? -x.z : x.z);
        atomicAdd(&s_hist[t3 & 0x7fffu], (t3 & 0x8000u) ? -x.w : x.w);
    }
    if (blockIdx.x == 0) {                                   // ragged end (Pf not a multiple of 4)
        for (int i = (n4 << 2) + threadIdx.x; i < Pf; i += blockDim.x) {
            const uint32_t t = table[i];
            atomicAdd(&s_hist[t & 0x7fffu], (t & 0x8000u) ? -src[i] : src[i]);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x)
        if (s_hist[k] != 0.f) atomicAdd(&out[(size_t)row * K + k], s_hist[k]);
}

// Block-scaled fp8 (OCP MX): e4m3 payload + one ue8m0 scale per 32 elements — the tcgen05
// `mxf8f6f4` operand format — for the published sketches.
__global__ void sketch_quant_mxfp8_kernel(const float* sk, int K, int Kpad, uint8_t* q, uint8_t* scales) {
    const int row = blockIdx.y;
    const int grp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (grp * 32 >= Kpad) return;
    const int k = grp * 32 + lane;
    const float x = (k < K) ? sk[(size_t)row * K + k] : 0.f;
    const float amax = warp_max(fabsf(x));
    int e = 0;                                               // scale = 2^e, amax/scale <= 448
    if (amax > 0.f) { frexpf(amax / 448.f, &e); }           // amax/448 = f·2^e with f in [0.5,1) → 2^e >= amax/448
    e = max(-127, min(127, e));
    const float inv = exp2f((float)-e);
    const __nv_fp8_e4m3 h(x * inv);
    q[(size_t)row * Kpad + k] = *reinterpret_cast<const uint8_t*>(&h);
    if (lane == 0) scales[(size_t)row * (Kpad / 32) + grp] = (uint8_t)(e + 127);
}

__device__ __forceinline__ float mxfp8_load(const uint8_t* q, const uint8_t* sc, int k) {
    __nv_fp8_e4m3 h; *reinterpret_cast<uint8_t*>(&h) = q[k];
    return (float)h * exp2f((float)((int)sc[k >> 5] - 127));
}

// =============================================================================================
// Filters — one small block per destination; produce per-edge weights for weighted_gather
// =============================================================================================
struct FilterCommon {
    EdgeTable et;
    float* w;            // [E] float-region weights
    float* w_tail;       // [E] int-tail weights
    const uint32_t* timed_out;
    float* stats;        // [V][4]: accepted, evaluated, threshold, aux
};

__device__ __forceinline__ bool edge_alive(const FilterCommon& c, int e) {
    const uint32_t dead = c.timed_out ? *c.timed_out : 0u;
    return c.et.mask[e] != 0.f && !((dead >> c.et.src_rank[e]) & 1u);
}

// FedAvg: unit weights (renormalised inside weighted_gather); ints keep own.
__global__ void fedavg_weights_kernel(FilterCommon c, int V) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const int e0 = c.et.row_ptr[v], e1 = c.et.row_ptr[v + 1];
    for (int e = e0; e < e1; ++e) { c.w[e] = 1.f; c.w_tail[e] = (e == e0) ? 1.f : 0.f; }
}

// BALANCE / Sketchguard decision given distances: accept d <= thr, closest fallback, blend weights.
//   tail_first = 0: tail averaged like floats (BALANCE); 1: alpha*own + (1-alpha)*first accepted (Sketchguard)
__device__ void threshold_accept(const FilterCommon& c, int v, const float* dist /*per edge, sqrt'ed*/, float thr,
                                 float alpha, int min_neighbors, int tail_first, float* rate_out) {
    const int e0 = c.et.row_ptr[v], e1 = c.et.row_ptr[v + 1];
    int alive = 0, accepted = 0, closest = -1; float best = INFINITY;
    for (int e = e0 + 1; e < e1; ++e) {
        c.w[e] = 0.f; c.w_tail[e] = 0.f;
        if (!edge_alive(c, e)) continue;
        ++alive;
        if (dist[e] < best) { best = dist[e]; closest = e; }
        if (dist[e] <= thr) { c.w[e] = 1.f; ++accepted; }
    }
    if (rate_out) *rate_out = (float)accepted / (float)max(1, alive);
    c.stats[v * 4 + 0] = (float)accepted; c.stats[v * 4 + 1] = (float)alive; c.stats[v * 4 + 2] = thr;
    int fallback = -1;
    if (accepted < min_neighbors && alive > 0 && c.w[closest] == 0.f) { c.w[closest] = 1.f; ++accepted; fallback = closest; }
    if (accepted == 0) { c.w[e0] = 1.f; c.w_tail[e0] = 1.f; return; }
    const float wn = (1.f - alpha) / (float)accepted;
    int first = -1;
    for (int e = e0 + 1; e < e1; ++e) if (c.w[e] != 0.f) { if (e != fallback && first < 0) first = e; c.w[e] = wn; }
    if (first < 0) first = fallback;                 // accepted list order: thresholded ones, then the fallback
    c.w[e0] = alpha; c.w_tail[e0] = alpha;
    if (tail_first) c.w_tail[first] = 1.f - alpha;
    else for (int e = e0 + 1; e < e1; ++e) c.w_tail[e] = c.w[e];
}

__global__ void balance_filter_kernel(FilterCommon c, int V, const float* d2, const float* n2, float* dist_out,
                                      float factor /* γ·exp(-κ t/T) */, float alpha, int min_neighbors) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    for (int e = c.et.row_ptr[v]; e < c.et.row_ptr[v + 1]; ++e) dist_out[e] = sqrtf(d2[e]);
    threshold_accept(c, v, dist_out, factor * sqrtf(n2[v]), alpha, min_neighbors, 0, nullptr);
}

// Sketchguard: sketch distances (fp32 or mxfp8 published sketches, local or peer), adaptive threshold
// with the attack factor from the node's last three acceptance rates, blend weights.
struct SketchFilterArgs {
    FilterCommon c;
    const float* own_sketch;            // [V][K] sketches of the live states (fp32, local)
    const float* const* peer_sketch;    // [G] fp32 published sketches  [plane][S][K]   (fp32 mode)
    const uint8_t* const* peer_q;       // [G] e4m3 payloads            [plane][S][Kpad] (fp8 mode)
    const uint8_t* const* peer_sc;      // [G] ue8m0 scales             [plane][S][Kpad/32]
    size_t plane_slots;                 // parity * S
    int K, Kpad, fp8;
    float factor, alpha; int min_neighbors;
    float* hist;                        // [V][4]: ring of the last 3 acceptance rates + count
    float* dist_out;                    // [E]
    const uint32_t* flags; int G; uint32_t epoch; long long timeout; uint32_t* timed_out;
};

__global__ void __launch_bounds__(128) sketchguard_filter_kernel(SketchFilterArgs a) {
    __shared__ float s_scratch[32];
    wait_published(a.flags, a.G, a.epoch, a.timeout, a.timed_out);
    const int v = blockIdx.x;
    const int e0 = a.c.et.row_ptr[v], e1 = a.c.et.row_ptr[v + 1];
    const float* own = a.own_sketch + (size_t)v * a.K;
    float nrm = 0.f;
    for (int k = threadIdx.x; k < a.K; k += blockDim.x) nrm += own[k] * own[k];
    nrm = sqrtf(block_sum(nrm, s_scratch));
    for (int e = e0 + 1; e < e1; ++e) {
        float s = 0.f;
        if (edge_alive(a.c, e)) {
            const size_t row = a.plane_slots + (size_t)a.c.et.src_slot[e];
            const int r = a.c.et.src_rank[e];
            if (a.fp8) {
                const uint8_t* q = a.peer_q[r] + row * a.Kpad;
                const uint8_t* sc = a.peer_sc[r] + row * (a.Kpad / 32);
                for (int k = threadIdx.x; k < a.K; k += blockDim.x) { const float d = own[k] - mxfp8_load(q, sc, k); s = fmaf(d, d, s); }
            } else {
                const float* o = a.peer_sketch[r] + row * a.K;
                for (int k = threadIdx.x; k < a.K; k += blockDim.x) { const float d = own[k] - o[k]; s = fmaf(d, d, s); }
            }
        }
        s = block_sum(s, s_scratch);
        if (threadIdx.x == 0) a.dist_out[e] = sqrtf(s);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        float* h = a.hist + v * 4;
        const int cnt = (int)h[3];
        float af = 1.f;
        if (cnt >= 3 && (h[0] + h[1] + h[2]) / 3.f < 0.3f) af = 1.5f;
        float rate = 0.f;
        threshold_accept(a.c, v, a.dist_out, a.factor * af * nrm, a.alpha, a.min_neighbors, 1, &rate);
        h[cnt % 3] = rate; h[3] = (float)(cnt + 1);      // ring holds the last three rates
        a.c.stats[v * 4 + 3] = af;
    }
}

// UBAR stage 1: shortlist the max(min_nb, int(rho·d)) closest alive neighbours (stable by edge order).
__global__ void ubar_stage1_kernel(FilterCommon c, int V, const float* d2, float rho, int min_neighbors,
                                   float* cand /*[E] 0/1*/, float* rank_out /*[E]*/) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const int e0 = c.et.row_ptr[v], e1 = c.et.row_ptr[v + 1];
    int alive = 0;
    for (int e = e0 + 1; e < e1; ++e) alive += edge_alive(c, e);
    const int pick = max(min_neighbors, (int)(rho * (float)alive));
    int chosen = 0;
    cand[e0] = 0.f; rank_out[e0] = -1.f;
    for (int e = e0 + 1; e < e1; ++e) {
        cand[e] = 0.f; rank_out[e] = 1e9f;
        if (!edge_alive(c, e)) continue;
        int rank = 0;
        for (int f = e0 + 1; f < e1; ++f)
            if (f != e && edge_alive(c, f) && (d2[f] < d2[e] || (d2[f] == d2[e] && f < e))) ++rank;
        rank_out[e] = (float)rank;
        if (rank < pick) { cand[e] = 1.f; ++chosen; }
    }
    c.stats[v * 4 + 0] = (float)chosen; c.stats[v * 4 + 1] = (float)alive;
}

// UBAR stage 2: keep candidates with loss <= own loss (best-loss fallback); blend weights.
__global__ void ubar_stage2_kernel(FilterCommon c, int V, const float* cand, const float* rank, const float* loss,
                                   const float* own_loss, float alpha, int use_loss) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const int e0 = c.et.row_ptr[v], e1 = c.et.row_ptr[v + 1];
    int kept = 0, ncand = 0, best = -1; float best_loss = INFINITY, best_rank = 1e9f;
    for (int e = e0 + 1; e < e1; ++e) {
        c.w[e] = 0.f; c.w_tail[e] = 0.f;
        if (cand[e] == 0.f) continue;
        ++ncand;
        if (use_loss) {
            if (loss[e] < best_loss || (loss[e] == best_loss && rank[e] < best_rank)) { best_loss = loss[e]; best_rank = rank[e]; best = e; }
            if (loss[e] <= own_loss[v]) { c.w[e] = 1.f; ++kept; }
        } else { c.w[e] = 1.f; ++kept; }
    }
    if (kept == 0 && best >= 0) { c.w[best] = 1.f; kept = 1; }
    c.stats[v * 4 + 2] = (float)kept; c.stats[v * 4 + 3] = (float)ncand;
    if (kept == 0) { c.w[e0] = 1.f; c.w_tail[e0] = 1.f; return; }
    int first = -1; float first_rank = 1e9f;                 // dict order of the reference = stage-1 distance order
    for (int e = e0 + 1; e < e1; ++e) if (c.w[e] != 0.f) { c.w[e] = (1.f - alpha) / (float)kept; if (rank[e] < first_rank) { first_rank = rank[e]; first = e; } }
    c.w[e0] = alpha; c.w_tail[e0] = alpha; c.w_tail[first] = 1.f - alpha;
}

// Krum: scores from the dense per-destination distance table (squared), single winner.
__global__ void krum_select_kernel(FilterCommon c, int V, const float* D /*[V][32][32] squared*/, int num_compromised,
                                   int* winner_out) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const int e0 = c.et.row_ptr[v], e1 = c.et.row_ptr[v + 1];
    const int m = e1 - e0;
    for (int e = e0; e < e1; ++e) { c.w[e] = 0.f; c.w_tail[e] = 0.f; }
    int win = 0;
    // Krum over the states that actually arrived (own + alive neighbours), like the reference's deadline-driven partial
    // aggregation (distributed/node_process.py:241-244): a dropped / timed-out edge shrinks m instead of disabling the rule
    int alive[kPairM]; int ma = 0;
    if (m <= kPairM) {
        alive[ma++] = 0;
        for (int e = e0 + 1; e < e1; ++e) if (edge_alive(c, e)) alive[ma++] = e - e0;
    }
    if (ma >= 1 && (float)num_compromised < (float)(ma - 2) * 0.5f) {
        const float* Dv = D + (size_t)v * kPairM * kPairM;
        const int keep = max(1, ma - num_compromised - 2);
        float best = INFINITY;
        for (int a = 0; a < ma; ++a) {
            const int i = alive[a];
            float row[kPairM]; int n = 0;
            for (int b = 0; b < ma; ++b) if (b != a) {         // insertion sort of the distances to the other arrived states
                float d = sqrtf(fmaxf(Dv[i * kPairM + alive[b]], 0.f));
                int p = n++;
                while (p > 0 && row[p - 1] > d) { row[p] = row[p - 1]; --p; }
                row[p] = d;
            }
            float s = 0.f;
            for (int k = 0; k < keep && k < n; ++k) s += row[k];
            if (s < best) { best = s; win = i; }
        }
    }
    c.w[e0 + win] = 1.f; c.w_tail[e0 + win] = 1.f;
    if (winner_out) winner_out[v] = win;
    c.stats[v * 4 + 0] = (float)win; c.stats[v * 4 + 1] = (float)m;
}

// EvidentialTrust: per-edge (vacuity, accuracy) → trust → EMA → threshold → normalised weights.
__global__ void trust_filter_kernel(FilterCommon c, int V, const float* vac, const float* acc, const int* src_gid, int N,
                                    float* ema /*[V][N]*/, float* ema_valid /*[V][N]*/, float accuracy_weight,
                                    float vacuity_threshold, float momentum, int adaptive, float threshold,
                                    float self_weight, float* trust_out /*[E]*/) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const int e0 = c.et.row_ptr[v], e1 = c.et.row_ptr[v + 1];
    float total = 0.f; int accepted = 0, evaluated = 0;
    for (int e = e0 + 1; e < e1; ++e) {
        c.w[e] = 0.f; c.w_tail[e] = 0.f; trust_out[e] = 0.f;
        if (!edge_alive(c, e)) continue;
        ++evaluated;
        float t = (1.f - vac[e]) * (accuracy_weight * acc[e] + (1.f - accuracy_weight));
        if (vac[e] > vacuity_threshold) t *= expf(-(vac[e] - vacuity_threshold));
        t = fminf(1.f, fmaxf(0.f, t));
        if (adaptive) {
            const int g = src_gid[e];
            float* slot = ema + (size_t)v * N + g; float* ok = ema_valid + (size_t)v * N + g;
            if (*ok != 0.f) t = momentum * t + (1.f - momentum) * (*slot);
            *slot = t; *ok = 1.f;
        }
        trust_out[e] = t;
        if (t >= threshold) { c.w[e] = t; total += t; ++accepted; }
    }
    c.stats[v * 4 + 0] = (float)accepted; c.stats[v * 4 + 1] = (float)evaluated; c.stats[v * 4 + 2] = threshold;
    c.w_tail[e0] = 1.f;                                        // ints keep own (average_states copies states[0])
    if (accepted == 0) { c.w[e0] = 1.f; return; }
    for (int e = e0 + 1; e < e1; ++e) if (c.w[e] != 0.f) c.w[e] = (1.f - self_weight) * c.w[e] / total;
    c.w[e0] = self_weight;
}

}  // namespace mb

// =============================================================================================
// Host wrappers
// =============================================================================================
using torch::Tensor;

namespace {

inline cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream(); }

inline int grid_x_for(int work_items, int per_block, int rows) {
    // ~4 waves of 148 SMs across all rows, never more blocks than work
    int want = std::max(1, (148 * 8) / std::max(1, rows));
    int have = std::max(1, (work_items + per_block - 1) / per_block);
    return std::min(want, have);
}

mb::EdgeTable make_et(const Tensor& row_ptr, const Tensor& src_rank, const Tensor& src_slot, const Tensor& mask) {
    TORCH_CHECK(row_ptr.dtype() == torch::kInt32 && src_rank.dtype() == torch::kInt32 && src_slot.dtype() == torch::kInt32);
    TORCH_CHECK(mask.dtype() == torch::kFloat32);
    return mb::EdgeTable{row_ptr.data_ptr<int>(), src_rank.data_ptr<int>(), src_slot.data_ptr<int>(), mask.data_ptr<float>()};
}

struct Sync {   // optional cross-GPU flag wait parameters
    const uint32_t* flags = nullptr; int G = 1; uint32_t epoch = 0; long long timeout = 0; uint32_t* timed_out = nullptr;
};
Sync make_sync(int64_t flags_ptr, int64_t G, int64_t epoch, double timeout_ms, int64_t timed_out_ptr) {
    Sync s;
    s.flags = reinterpret_cast<const uint32_t*>(flags_ptr);
    s.G = (int)G; s.epoch = (uint32_t)epoch;
    static int khz = 0;                                  // clock64() ticks at the SM clock: cycles per millisecond = kHz
    if (khz == 0) {
        int dev = 0; cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev) != cudaSuccess || khz <= 0) khz = 1900000;
    }
    s.timeout = (long long)(timeout_ms * (double)khz);
    s.timed_out = reinterpret_cast<uint32_t*>(timed_out_ptr);
    return s;
}

}  // namespace

void publish(Tensor live, int64_t pub_ptr, int64_t stride, int64_t V, int64_t Pf, int64_t Pf_pad, c10::optional<Tensor> ints,
             Tensor scale, Tensor noise_std, Tensor node_gid, int64_t seed, int64_t round, int64_t peer_flags_ptr,
             int64_t G, int64_t my_rank, int64_t epoch, Tensor ticket) {
    if (V == 0) return;
    c10::cuda::CUDAGuard guard(live.device());
    mb::PublishArgs a;
    a.live = live.data_ptr<float>(); a.pub = reinterpret_cast<float*>(pub_ptr); a.stride = (size_t)stride;
    a.Pf = (int)Pf; a.Pf_pad = (int)Pf_pad;
    a.ints = ints.has_value() && ints->numel() > 0 ? reinterpret_cast<const long long*>(ints->data_ptr<int64_t>()) : nullptr;
    a.n_int = a.ints ? (int)ints->size(1) : 0;
    a.scale = scale.data_ptr<float>(); a.noise_std = noise_std.data_ptr<float>(); a.node_gid = node_gid.data_ptr<int>();
    a.seed = (unsigned long long)seed; a.round = (unsigned long long)round;
    a.peer_flags = reinterpret_cast<uint32_t* const*>(peer_flags_ptr); a.G = (int)G; a.my_rank = (int)my_rank; a.epoch = (uint32_t)epoch;
    a.ticket = reinterpret_cast<unsigned int*>(ticket.data_ptr<int>());
    dim3 grid(grid_x_for((int)Pf_pad / 4, mb::kThreads, (int)V), (unsigned)V);
    mb::publish_kernel<<<grid, mb::kThreads, 0, cur_stream()>>>(a);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void publish_sum(Tensor live, int64_t pub_ptr, int64_t rsum_ptr, int64_t stride, int64_t V, int64_t Pf, int64_t Pf_pad,
                 c10::optional<Tensor> ints, Tensor scale, Tensor noise_std, Tensor node_gid, int64_t seed, int64_t round,
                 int64_t peer_flags_ptr, int64_t G, int64_t my_rank, int64_t epoch, Tensor ticket) {
    if (V == 0) return;
    c10::cuda::CUDAGuard guard(live.device());
    mb::PublishArgs a;
    a.live = live.data_ptr<float>(); a.pub = reinterpret_cast<float*>(pub_ptr); a.stride = (size_t)stride;
    a.Pf = (int)Pf; a.Pf_pad = (int)Pf_pad;
    a.ints = ints.has_value() && ints->numel() > 0 ? reinterpret_cast<const long long*>(ints->data_ptr<int64_t>()) : nullptr;
    a.n_int = a.ints ? (int)ints->size(1) : 0;
    a.scale = scale.data_ptr<float>(); a.noise_std = noise_std.data_ptr<float>(); a.node_gid = node_gid.data_ptr<int>();
    a.seed = (unsigned long long)seed; a.round = (unsigned long long)round;
    a.peer_flags = reinterpret_cast<uint32_t* const*>(peer_flags_ptr); a.G = (int)G; a.my_rank = (int)my_rank; a.epoch = (uint32_t)epoch;
    a.ticket = reinterpret_cast<unsigned int*>(ticket.data_ptr<int>());
    const int grid = grid_x_for((int)Pf_pad / 4, mb::kThreads, 1);
    mb::publish_sum_kernel<<<grid, mb::kThreads, 0, cur_stream()>>>(a, reinterpret_cast<float*>(rsum_ptr), (int)V);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void fullmesh_reduce_scatter(Tensor anchor, int64_t peer_rsum_tbl, int64_t mc_rsum_ptr, int64_t peer_tot_tbl, int64_t mc_tot_ptr, int64_t len,
                             int64_t G, int64_t my_rank, int64_t timed_out_ptr, int64_t peer_flags_ptr, int64_t epoch, Tensor ticket) {
    c10::cuda::CUDAGuard guard(anchor.device());
    TORCH_CHECK(G >= 2 && peer_flags_ptr != 0, "fullmesh_reduce_scatter needs >= 2 ranks");
    mb::ReduceScatterArgs a;
    a.peer_rsum = reinterpret_cast<const float* const*>(peer_rsum_tbl); a.mc_rsum = reinterpret_cast<const float*>(mc_rsum_ptr);
    a.peer_tot = reinterpret_cast<float* const*>(peer_tot_tbl); a.mc_tot = reinterpret_cast<float*>(mc_tot_ptr);
    a.len4 = (int)(len / 4); a.G = (int)G; a.my_rank = (int)my_rank; a.timed_out = reinterpret_cast<const uint32_t*>(timed_out_ptr);
    a.peer_flags = reinterpret_cast<uint32_t* const*>(peer_flags_ptr); a.epoch = (uint32_t)epoch;
    a.ticket = reinterpret_cast<unsigned int*>(ticket.data_ptr<int>());
    const int slice4 = (a.len4 + a.G - 1) / a.G;
    const int grid = std::max(1, std::min(148 * 4, (slice4 + mb::kThreads - 1) / mb::kThreads));
    mb::fullmesh_reduce_scatter_kernel<<<grid, mb::kThreads, 0, cur_stream()>>>(a);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void fedavg_fullmesh(Tensor live, int64_t pub_local_ptr, int64_t peer_rsum_tbl, int64_t mc_rsum_ptr, int64_t stride, int64_t V,
                     int64_t len, int64_t N, int64_t G, Tensor byz, Tensor rank_nodes, int64_t timed_out_ptr, int64_t tot_local_ptr) {
    if (V == 0) return;
    c10::cuda::CUDAGuard guard(live.device());
    TORCH_CHECK(rank_nodes.dtype() == torch::kInt32 && rank_nodes.numel() >= G && byz.numel() >= V);
    mb::FullMeshArgs a;
    a.live = live.data_ptr<float>(); a.pub_local = reinterpret_cast<const float*>(pub_local_ptr);
    a.peer_rsum = reinterpret_cast<const float* const*>(peer_rsum_tbl); a.mc_rsum = reinterpret_cast<const float*>(mc_rsum_ptr);
    a.tot_local = reinterpret_cast<const float*>(tot_local_ptr);
    a.stride = (size_t)stride; a.V = (int)V; a.len4 = (int)(len / 4); a.G = (int)G; a.N = (int)N;
    a.byz = byz.data_ptr<uint8_t>(); a.rank_nodes = rank_nodes.data_ptr<int>();
    a.timed_out = reinterpret_cast<const uint32_t*>(timed_out_ptr);
    const int grid = std::max(1, std::min(148 * 8, (a.len4 + mb::kThreads - 1) / mb::kThreads));
    mb::fedavg_fullmesh_kernel<<<grid, mb::kThreads, 0, cur_stream()>>>(a);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void weighted_gather(Tensor live, int64_t peer_pub_ptr, int64_t parity_off, int64_t stride, int64_t V, Tensor row_ptr,
                     Tensor src_rank, Tensor src_slot, Tensor mask, Tensor w, int64_t len, bool renorm,
                     int64_t flags_ptr, int64_t G, int64_t epoch, double timeout_ms, int64_t timed_out_ptr, bool use_tma) {
    if (V == 0) return;
    c10::cuda::CUDAGuard guard(live.device());
    mb::GatherArgs a;
    a.live = live.data_ptr<float>();
    a.pv = mb::PeerView{reinterpret_cast<const float* const*>(peer_pub_ptr), (size_t)parity_off, (size_t)stride};
    a.et = make_et(row_ptr, src_rank, src_slot, mask);
    a.w = w.data_ptr<float>(); a.len4 = (int)(len / 4); a.renorm = renorm ? 1 : 0;
    Sync s = make_sync(flags_ptr, G, epoch, timeout_ms, timed_out_ptr);
    a.flags = s.flags; a.G = s.G; a.epoch = s.epoch; a.timeout = s.timeout; a.timed_out = s.timed_out;
    if (use_tma) {
        const int smem = mb::kBulkStages * mb::kBulkTile * 4;
        static bool attr = false;
        if (!attr) { cudaFuncSetAttribute(mb::weighted_gather_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = true; }
        const int ntiles = (int)((len + mb::kBulkTile - 1) / mb::kBulkTile);
        dim3 grid(std::max(1, std::min(ntiles, (148 * 3) / std::max<int>(1, (int)V))), (unsigned)V);
        mb::weighted_gather_bulk_kernel<<<grid, mb::kBulkThreads, smem, cur_stream()>>>(a);
    } else {
        dim3 grid(grid_x_for(a.len4, mb::kThreads, (int)V), (unsigned)V);
        mb::weighted_gather_kernel<<<grid, mb::kThreads, 0, cur_stream()>>>(a);
    }
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void nvls_fedavg(Tensor live, int64_t pub_local_ptr, int64_t mc_pub_ptr, int64_t stride, int64_t V, int64_t S, int64_t len,
                 int64_t N, Tensor byz, int64_t flags_ptr, int64_t G, int64_t epoch, double timeout_ms, int64_t timed_out_ptr) {
    if (V == 0) return;
    c10::cuda::CUDAGuard guard(live.device());
    TORCH_CHECK(mc_pub_ptr != 0, "NVLS multicast pointer is null (multicast not supported on this system)");
    mb::NvlsArgs a;
    a.live = live.data_ptr<float>(); a.pub_local = reinterpret_cast<const float*>(pub_local_ptr);
    a.mc_pub = reinterpret_cast<const float*>(mc_pub_ptr); a.stride = (size_t)stride;
    a.V = (int)V; a.S = (int)S; a.len4 = (int)(len / 4); a.inv_n = 1.f / (float)N; a.byz = byz.data_ptr<uint8_t>();
    Sync s = make_sync(flags_ptr, G, epoch, timeout_ms, timed_out_ptr);
    a.flags = s.flags; a.G = s.G; a.epoch = s.epoch; a.timeout = s.timeout; a.timed_out = s.timed_out;
    const int grid = std::max(1, std::min(148 * 4, (a.len4 + mb::kThreads - 1) / mb::kThreads));
    mb::nvls_fedavg_kernel<<<grid, mb::kThreads, 0, cur_stream()>>>(a);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void wait_epoch(Tensor anchor, int64_t flags_ptr, int64_t G, int64_t epoch, double timeout_ms, int64_t timed_out_ptr) {
    if (G <= 1 || flags_ptr == 0) return;
    c10::cuda::CUDAGuard guard(anchor.device());
    Sync s = make_sync(flags_ptr, G, epoch, timeout_ms, timed_out_ptr);
    mb::wait_epoch_kernel<<<1, 32, 0, cur_stream()>>>(s.flags, s.G, s.epoch, s.timeout, s.timed_out);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void tail_blend(Tensor live, int64_t peer_pub_ptr, int64_t parity_off, int64_t stride, int64_t V, Tensor row_ptr, Tensor src_rank,
                Tensor src_slot, Tensor mask, Tensor w_tail, int64_t Pf_pad, Tensor ints, int64_t timed_out_ptr) {
    if (V == 0 || ints.numel() == 0) return;
    c10::cuda::CUDAGuard guard(live.device());
    mb::PeerView pv{reinterpret_cast<const float* const*>(peer_pub_ptr), (size_t)parity_off, (size_t)stride};
    mb::tail_blend_kernel<<<(unsigned)V, 64, 0, cur_stream()>>>(live.data_ptr<float>(), pv, make_et(row_ptr, src_rank, src_slot, mask),
                                                                w_tail.data_ptr<float>(), (int)Pf_pad, reinterpret_cast<long long*>(ints.data_ptr<int64_t>()),
                                                                (int)ints.size(1), reinterpret_cast<const uint32_t*>(timed_out_ptr));
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void edge_distances(Tensor live, int64_t peer_pub_ptr, int64_t parity_off, int64_t stride, int64_t V, Tensor row_ptr, Tensor src_rank,
                    Tensor src_slot, Tensor mask, int64_t len, Tensor d2, Tensor n2, int64_t flags_ptr, int64_t G, int64_t epoch,
                    double timeout_ms, int64_t timed_out_ptr) {
    if (V == 0) return;
    c10::cuda::CUDAGuard guard(live.device());
    mb::DistArgs a;
    a.live = live.data_ptr<float>();
    a.pv = mb::PeerView{reinterpret_cast<const float* const*>(peer_pub_ptr), (size_t)parity_off, (size_t)stride};
    a.et = make_et(row_ptr, src_rank, src_slot, mask);
    a.len4 = (int)(len / 4); a.d2 = d2.data_ptr<float>(); a.n2 = n2.data_ptr<float>();
    Sync s = make_sync(flags_ptr, G, epoch, timeout_ms, timed_out_ptr);
    a.flags = s.flags; a.G = s.G; a.epoch = s.epoch; a.timeout = s.timeout; a.timed_out = s.timed_out;
    cudaMemsetAsync(a.d2, 0, d2.numel() * sizeof(float), cur_stream());
    cudaMemsetAsync(a.n2, 0, n2.numel() * sizeof(float), cur_stream());
    dim3 grid(grid_x_for(a.len4, mb::kThreads * 4, (int)V), (unsigned)V);
    mb::edge_distances_kernel<<<grid, mb::kThreads, 0, cur_stream()>>>(a);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void pairwise_distances(Tensor live, int64_t peer_pub_ptr, int64_t parity_off, int64_t stride, int64_t V, Tensor row_ptr, Tensor src_rank,
                        Tensor src_slot, Tensor mask, int64_t len, Tensor D, int64_t max_m, int64_t flags_ptr, int64_t G, int64_t epoch,
                        double timeout_ms, int64_t timed_out_ptr) {
    if (V == 0) return;
    c10::cuda::CUDAGuard guard(live.device());
    TORCH_CHECK(max_m <= mb::kPairM, "pairwise_distances supports at most ", mb::kPairM, " candidates per node");
    mb::PairArgs a;
    a.live = live.data_ptr<float>();
    a.pv = mb::PeerView{reinterpret_cast<const float* const*>(peer_pub_ptr), (size_t)parity_off, (size_t)stride};
    a.et = make_et(row_ptr, src_rank, src_slot, mask);
    a.len = (int)len; a.D = D.data_ptr<float>();
    Sync s = make_sync(flags_ptr, G, epoch, timeout_ms, timed_out_ptr);
    a.flags = s.flags; a.G = s.G; a.epoch = s.epoch; a.timeout = s.timeout; a.timed_out = s.timed_out;
    cudaMemsetAsync(a.D, 0, D.numel() * sizeof(float), cur_stream());
    const size_t smem = (size_t)max_m * mb::kPairTile * sizeof(float);
    dim3 grid(grid_x_for((int)len, mb::kPairTile, (int)V), (unsigned)V);
    mb::pairwise_kernel<<<grid, mb::kThreads, smem, cur_stream()>>>(a);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// D [V][32][32] (Gram-derived squared distances) → pairs below τ·(‖a‖²+‖b‖²) replaced by their exact fp32 value; `scratch` like D.
void krum_refine(Tensor live, int64_t peer_pub_ptr, int64_t parity_off, int64_t stride, int64_t V, Tensor row_ptr, Tensor src_rank,
                 Tensor src_slot, Tensor mask, int64_t len, Tensor D, Tensor norms, double tau, Tensor scratch, c10::optional<Tensor> nref) {
    if (V == 0) return;
    c10::cuda::CUDAGuard guard(live.device());
    TORCH_CHECK(D.numel() == V * mb::kPairM * mb::kPairM && scratch.numel() == D.numel() && norms.numel() == V * mb::kPairM);
    mb::PairArgs a;
    a.live = live.data_ptr<float>();
    a.pv = mb::PeerView{reinterpret_cast<const float* const*>(peer_pub_ptr), (size_t)parity_off, (size_t)stride};
    a.et = make_et(row_ptr, src_rank, src_slot, mask);
    a.len = (int)len; a.D = D.data_ptr<float>();
    a.flags = nullptr; a.G = 1; a.epoch = 0; a.timeout = 0; a.timed_out = nullptr;
    cudaMemsetAsync(scratch.data_ptr<float>(), 0, scratch.numel() * sizeof(float), cur_stream());
    const int chunks = (int)std::max<int64_t>(1, std::min<int64_t>(16, len / (64 * 1024)));
    dim3 grid((unsigned)(mb::kPairM * (mb::kPairM - 1) / 2 * chunks), (unsigned)V);
    mb::krum_refine_kernel<<<grid, mb::kThreads, 0, cur_stream()>>>(a, norms.data_ptr<float>(), (float)tau, scratch.data_ptr<float>(), chunks);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    mb::krum_refine_apply_kernel<<<(unsigned)V, 128, 0, cur_stream()>>>(a.et, D.data_ptr<float>(), norms.data_ptr<float>(), scratch.data_ptr<float>(),
                                                                       (float)tau, nref.has_value() ? nref->data_ptr<float>() : nullptr);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void count_sketch(int64_t base_ptr, int64_t stride, Tensor slots, Tensor table, int64_t Pf, int64_t K, Tensor out) {
    const int rows = (int)slots.numel();
    if (rows == 0) return;
    c10::cuda::CUDAGuard guard(out.device());
    TORCH_CHECK(K <= 8192, "sketch_size up to 8192 supported");
    cudaMemsetAsync(out.data_ptr<float>(), 0, (size_t)rows * K * sizeof(float), cur_stream());
    dim3 grid(grid_x_for((int)Pf / 4, mb::kThreads * 4, rows), (unsigned)rows);
    mb::count_sketch_kernel<<<grid, mb::kThreads, (size_t)K * sizeof(float), cur_stream()>>>(
        reinterpret_cast<const float*>(base_ptr), (size_t)stride, slots.data_ptr<int>(),
        reinterpret_cast<const uint16_t*>(table.data_ptr()), (int)Pf, (int)K, out.data_ptr<float>());
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void sketch_quant_mxfp8(Tensor sk, int64_t q_ptr, int64_t sc_ptr, int64_t Kpad) {
    const int rows = (int)sk.size(0), K = (int)sk.size(1);
    if (rows == 0) return;
    c10::cuda::CUDAGuard guard(sk.device());
    dim3 grid(((int)Kpad / 32 + 3) / 4, rows);
    mb::sketch_quant_mxfp8_kernel<<<grid, 128, 0, cur_stream()>>>(sk.data_ptr<float>(), K, (int)Kpad,
                                                                  reinterpret_cast<uint8_t*>(q_ptr), reinterpret_cast<uint8_t*>(sc_ptr));
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

namespace {
mb::FilterCommon make_fc(Tensor row_ptr, Tensor src_rank, Tensor src_slot, Tensor mask, Tensor w, Tensor w_tail, int64_t timed_out_ptr, Tensor stats) {
    return mb::FilterCommon{make_et(row_ptr, src_rank, src_slot, mask), w.data_ptr<float>(), w_tail.data_ptr<float>(),
                            reinterpret_cast<const uint32_t*>(timed_out_ptr), stats.data_ptr<float>()};
}
}  // namespace

void fedavg_weights(int64_t V, Tensor row_ptr, Tensor src_rank, Tensor src_slot, Tensor mask, Tensor w, Tensor w_tail, Tensor stats) {
    if (V == 0) return;
    c10::cuda::CUDAGuard guard(w.device());
    mb::fedavg_weights_kernel<<<((int)V + 63) / 64, 64, 0, cur_stream()>>>(make_fc(row_ptr, src_rank, src_slot, mask, w, w_tail, 0, stats), (int)V);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void balance_filter(int64_t V, Tensor row_ptr, Tensor src_rank, Tensor src_slot, Tensor mask, Tensor w, Tensor w_tail, Tensor stats,
                    Tensor d2, Tensor n2, Tensor dist_out, double factor, double alpha, int64_t min_neighbors, int64_t timed_out_ptr) {
    if (V == 0) return;
    c10::cuda::CUDAGuard guard(w.device());
    mb::balance_filter_kernel<<<((int)V + 63) / 64, 64, 0, cur_stream()>>>(
        make_fc(row_ptr, src_rank, src_slot, mask, w, w_tail, timed_out_ptr, stats), (int)V, d2.data_ptr<float>(), n2.data_ptr<float>(),
        dist_out.data_ptr<float>(), (float)factor, (float)alpha, (int)min_neighbors);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void sketchguard_filter(int64_t V, Tensor row_ptr, Tensor src_rank, Tensor src_slot, Tensor mask, Tensor w, Tensor w_tail, Tensor stats,
                        Tensor own_sketch, int64_t peer_sketch_ptr, int64_t peer_q_ptr, int64_t peer_sc_ptr, int64_t plane_slots,
                        int64_t K, int64_t Kpad, bool fp8, double factor, double alpha, int64_t min_neighbors, Tensor hist,
                        Tensor dist_out, int64_t flags_ptr, int64_t G, int64_t epoch, double timeout_ms, int64_t timed_out_ptr) {
    if (V == 0) return;
    c10::cuda::CUDAGuard guard(w.device());
    mb::SketchFilterArgs a;
    a.c = make_fc(row_ptr, src_rank, src_slot, mask, w, w_tail, timed_out_ptr, stats);
    a.own_sketch = own_sketch.data_ptr<float>();
    a.peer_sketch = reinterpret_cast<const float* const*>(peer_sketch_ptr);
    a.peer_q = reinterpret_cast<const uint8_t* const*>(peer_q_ptr);
    a.peer_sc = reinterpret_cast<const uint8_t* const*>(peer_sc_ptr);
    a.plane_slots = (size_t)plane_slots; a.K = (int)K; a.Kpad = (int)Kpad; a.fp8 = fp8 ? 1 : 0;
    a.factor = (float)factor; a.alpha = (float)alpha; a.min_neighbors = (int)min_neighbors;
    a.hist = hist.data_ptr<float>(); a.dist_out = dist_out.data_ptr<float>();
    Sync s = make_sync(flags_ptr, G, epoch, timeout_ms, timed_out_ptr);
    a.flags = s.flags; a.G = s.G; a.epoch = s.epoch; a.timeout = s.timeout; a.timed_out = s.timed_out;
    mb::sketchguard_filter_kernel<<<(unsigned)V, 128, 0, cur_stream()>>>(a);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void ubar_stage1(int64_t V, Tensor row_ptr, Tensor src_rank, Tensor src_slot, Tensor mask, Tensor w, Tensor w_tail, Tensor stats,
                 Tensor d2, double rho, int64_t min_neighbors, Tensor cand, Tensor rank, int64_t timed_out_ptr) {
    if (V == 0) return;
    c10::cuda::CUDAGuard guard(w.device());
    mb::ubar_stage1_kernel<<<((int)V + 63) / 64, 64, 0, cur_stream()>>>(
        make_fc(row_ptr, src_rank, src_slot, mask, w, w_tail, timed_out_ptr, stats), (int)V, d2.data_ptr<float>(), (float)rho,
        (int)min_neighbors, cand.data_ptr<float>(), rank.data_ptr<float>());
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void ubar_stage2(int64_t V, Tensor row_ptr, Tensor src_rank, Tensor src_slot, Tensor mask, Tensor w, Tensor w_tail, Tensor stats,
                 Tensor cand, Tensor rank, Tensor loss, Tensor own_loss, double alpha, bool use_loss) {
    if (V == 0) return;
    c10::cuda::CUDAGuard guard(w.device());
    mb::ubar_stage2_kernel<<<((int)V + 63) / 64, 64, 0, cur_stream()>>>(
        make_fc(row_ptr, src_rank, src_slot, mask, w, w_tail, 0, stats), (int)V, cand.data_ptr<float>(), rank.data_ptr<float>(),
        loss.data_ptr<float>(), own_loss.data_ptr<float>(), (float)alpha, use_loss ? 1 : 0);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void krum_select(int64_t V, Tensor row_ptr, Tensor src_rank, Tensor src_slot, Tensor mask, Tensor w, Tensor w_tail, Tensor stats,
                 Tensor D, int64_t num_compromised, Tensor winner, int64_t timed_out_ptr) {
    if (V == 0) return;
    c10::cuda::CUDAGuard guard(w.device());
    mb::krum_select_kernel<<<((int)V + 31) / 32, 32, 0, cur_stream()>>>(
        make_fc(row_ptr, src_rank, src_slot, mask, w, w_tail, timed_out_ptr, stats), (int)V, D.data_ptr<float>(), (int)num_compromised,
        winner.data_ptr<int>());
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void trust_filter(int64_t V, Tensor row_ptr, Tensor src_rank, Tensor src_slot, Tensor mask, Tensor w, Tensor w_tail, Tensor stats,
                  Tensor vac, Tensor acc, Tensor src_gid, int64_t N, Tensor ema, Tensor ema_valid, double accuracy_weight,
                  double vacuity_threshold, double momentum, bool adaptive, double threshold, double self_weight, Tensor trust_out,
                  int64_t timed_out_ptr) {
    if (V == 0) return;
    c10::cuda::CUDAGuard guard(w.device());
    mb::trust_filter_kernel<<<((int)V + 63) / 64, 64, 0, cur_stream()>>>(
        make_fc(row_ptr, src_rank, src_slot, mask, w, w_tail, timed_out_ptr, stats), (int)V, vac.data_ptr<float>(), acc.data_ptr<float>(),
        src_gid.data_ptr<int>(), (int)N, ema.data_ptr<float>(), ema_valid.data_ptr<float>(), (float)accuracy_weight,
        (float)vacuity_threshold, (float)momentum, adaptive ? 1 : 0, (float)threshold, (float)self_weight, trust_out.data_ptr<float>());
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}
