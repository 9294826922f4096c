// murmura_b200 — grouped MLP forward on 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
// Replaces the per-neighbour model evaluation of UBAR stage 2 (reference murmura/aggregation/ubar.py:152-222:
// load_state_dict + forward per candidate), EvidentialTrust (aggregation/evidential_trust.py:214-316: deepcopy(model) +
// forward per neighbour) and DMTT model scoring (dmtt/node_process.py:309-363: fresh model per received state) for the
// MLP families (Linear [+BatchNorm1d] +ReLU … + Linear/EvidentialHead).
//
// ONE launch per layer evaluates EVERY (destination node, candidate weights) pair hosted on the GPU:
//   group g:  Y_g[M_g, N] = act( BN_g( X_g[M_g, K] · W_gᵀ + b_g ) )
// where W_g / b_g / BN statistics are read IN PLACE from the candidate's published arena row — local HBM or a peer
// GPU's memory over NVLink (plain coalesced loads; rows are not 16-byte aligned for K = 561, so no TMA here) — staged
// into 128-byte-swizzled shared memory and multiplied by `tcgen05.mma.kind::tf32` (fp32 operands consumed directly,
// fp32 accumulation in TMEM).  Eval-mode BatchNorm, bias, ReLU and the Dirichlet head (softplus + 1) are fused into the
// TMEM → register epilogue, so a 3-layer HAR classifier is 3 launches for all edges instead of ~12 kernels per edge.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "common.cuh"

namespace mb {

constexpr int kLinBM = 128;          // rows of X per CTA  (UMMA M)
constexpr int kLinBN = 64;           // output features per CTA (UMMA N)
constexpr int kLinBK = 32;           // fp32 per k-block = one 128-byte swizzle row
constexpr int kLinStages = 3;
constexpr int kLinThreads = 128;
constexpr int kLinStageBytes = (kLinBM + kLinBN) * kLinBK * 4;      // 24 KiB

struct LinearGroup {                 // one (destination, candidate) pair — lives in a device array
    const float* X;                  // [M][ldx]   activations of the destination's samples
    const float* W;                  // [N][K]     candidate weights (row-major, K contiguous) — possibly peer memory
    const float* bias;               // [N] or null
    const float* bn_mean;            // [N] or null → eval-mode BatchNorm folded into the epilogue
    const float* bn_var;
    const float* bn_gamma;
    const float* bn_beta;
    float* Y;                        // [M][ldy]
    long long M;
};

__device__ __forceinline__ uint32_t lin_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void lin_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(lin_smem_u32(bar)), "r"(parity) : "memory");
        if (done) break;
        if (++spins > (1u << 22)) __trap();
    }
}

// K-major, 128-byte swizzle (8-row groups 1024 B apart), sm_100 descriptor version 1.
__device__ __forceinline__ uint64_t lin_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// Stage a [rows × 32] fp32 tile (row-major source with `ld` floats between rows) into swizzled smem; zero-fill OOB.
template <int ROWS>
__device__ __forceinline__ void stage_tile(uint8_t* tile, const float* __restrict__ src, long long ld, long long row0, long long nrows,
                                           int k0, int K, bool vec_ok) {
#pragma unroll
    for (int j = 0; j < ROWS * 8 / kLinThreads; ++j) {
        const int q = threadIdx.x + j * kLinThreads;
        const int r = q >> 3, c = q & 7;
        const long long grow = row0 + r;
        const int k = k0 + (c << 2);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (grow < nrows && k < K) {
            const float* p = src + grow * ld + k;
            if (vec_ok && k + 3 < K) v = ld_stream(reinterpret_cast<const float4*>(p));
            else {
                v.x = p[0];
                if (k + 1 < K) v.y = p[1];
                if (k + 2 < K) v.z = p[2];
                if (k + 3 < K) v.w = p[3];
            }
        }
        *reinterpret_cast<float4*>(tile + r * 128 + ((c ^ (r & 7)) << 4)) = v;      // 128B swizzle: chunk ^= row % 8
    }
}

__global__ void __launch_bounds__(kLinThreads) grouped_linear_tf32_kernel(const LinearGroup* __restrict__ groups, int K, int N, int ldx,
                                                                          int ldy, int act, float eps) {
    extern __shared__ uint8_t lin_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(lin_smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t mma_done[kLinStages];
    __shared__ uint32_t tmem_base_smem;

    const LinearGroup g = groups[blockIdx.z];
    const long long m0 = (long long)blockIdx.y * kLinBM;
    const int n0 = blockIdx.x * kLinBN;
    if (m0 >= g.M) return;                                          // uniform per CTA
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0) {
        if (lane == 0) {
            for (int s = 0; s < kLinStages; ++s)
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(lin_smem_u32(&mma_done[s])));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(lin_smem_u32(&tmem_base_smem)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    const bool vec_x = ((ldx & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.X) & 15) == 0);
    const bool vec_w = ((K & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.W) & 15) == 0);
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kLinBN >> 3) << 17) | ((uint32_t)(kLinBM >> 4) << 24);
    const int nk = (K + kLinBK - 1) / kLinBK;

    for (int it = 0; it < nk; ++it) {
        const int stage = it % kLinStages;
        if (it >= kLinStages) lin_mbar_wait(&mma_done[stage], (uint32_t)((it / kLinStages - 1) & 1));   // slot free again?
        uint8_t* tileA = smem + stage * kLinStageBytes;
        uint8_t* tileB = tileA + kLinBM * kLinBK * 4;
        stage_tile<kLinBM>(tileA, g.X, ldx, m0, g.M, it * kLinBK, K, vec_x);
        stage_tile<kLinBN>(tileB, g.W, K, n0, N, it * kLinBK, K, vec_w);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy smem writes → visible to the tensor core
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a0 = lin_smem_u32(tileA), b0 = lin_smem_u32(tileB);
#pragma unroll
            for (int k = 0; k < kLinBK / 8; ++k) {
                const uint32_t acc = (it > 0 || k > 0) ? 1u : 0u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                             :: "r"(tmem_base), "l"(lin_desc(a0 + k * 32)), "l"(lin_desc(b0 + k * 32)), "r"(idesc), "r"(acc) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                         :: "r"(lin_smem_u32(&mma_done[stage])) : "memory");
        }
    }
    const int last = nk - 1;
    lin_mbar_wait(&mma_done[last % kLinStages], (uint32_t)((last / kLinStages) & 1));                  // all MMAs retired
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: TMEM → registers → bias / BatchNorm(eval) / activation → global ------------------------------
    const long long row = m0 + warp * 32 + lane;
#pragma unroll
    for (int c0 = 0; c0 < kLinBN; c0 += 16) {
        uint32_t r[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row < g.M) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int n = n0 + c0 + j;
                if (n >= N) break;
                float v = __uint_as_float(r[j]);
                if (g.bias) v += g.bias[n];
                if (g.bn_mean) v = (v - g.bn_mean[n]) * rsqrtf(g.bn_var[n] + eps) * (g.bn_gamma ? g.bn_gamma[n] : 1.f) + (g.bn_beta ? g.bn_beta[n] : 0.f);
                if (act == 1) v = v < 0.f ? 0.f : v;        // NaN-propagating ReLU (torch.relu semantics)
                else if (act == 2) v = (v > 20.f ? v : log1pf(expf(v))) + 1.f;          // Dirichlet head: softplus + 1
                g.Y[row * ldy + n] = v;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(64));
}

// ---- grouped metrics: softmax-CE or Dirichlet statistics of each group's output rows --------------------------------
struct EvalGroup { const float* out; const long long* targets; long long M; };

__global__ void grouped_eval_kernel(const EvalGroup* __restrict__ groups, int C, int ld, int dirichlet, float* __restrict__ stats /*[G][8]*/) {
    const EvalGroup g = groups[blockIdx.y];
    const long long rowi = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (rowi >= g.M) return;
    const float* z = g.out + rowi * ld;
    float* st = stats + (size_t)blockIdx.y * 8;
    const int t = (int)g.targets[rowi];
    float mx = -INFINITY, S = 0.f; int arg = 0;
    for (int c = lane; c < C; c += 32) { const float x = z[c]; S += x; if (x > mx) { mx = x; arg = c; } }
    S = warp_sum(S);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, mx, o); const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
    }
    if (dirichlet) {
        float ent = 0.f, sq = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float p = z[c] / S;
            ent -= p * logf(p + 1e-10f);
            const float d = ((c == t) ? 1.f : 0.f) - p;
            sq = fmaf(d, d, sq);
        }
        ent = warp_sum(ent); sq = warp_sum(sq);
        if (lane == 0) {
            atomicAdd(st + 0, sq); atomicAdd(st + 1, arg == t ? 1.f : 0.f); atomicAdd(st + 2, 1.f);
            atomicAdd(st + 3, (float)C / S); atomicAdd(st + 4, ent); atomicAdd(st + 5, S);
        }
    } else {
        float se = 0.f;
        for (int c = lane; c < C; c += 32) se += __expf(z[c] - mx);
        se = warp_sum(se);
        if (lane == 0) { atomicAdd(st + 0, __logf(se) + mx - z[t]); atomicAdd(st + 1, arg == t ? 1.f : 0.f); atomicAdd(st + 2, 1.f); }
    }
}

}  // namespace mb

using torch::Tensor;

// groups: int64 tensor [G][9] on the device = {X, W, bias, bn_mean, bn_var, bn_gamma, bn_beta, Y, M} (pointers as integers)
void grouped_linear_tf32(Tensor groups, int64_t G, int64_t max_m, int64_t K, int64_t N, int64_t ldx, int64_t ldy, int64_t act, double eps) {
    if (G == 0 || max_m == 0) return;
    c10::cuda::CUDAGuard guard(groups.device());
    TORCH_CHECK(groups.dtype() == torch::kInt64 && groups.is_contiguous() && groups.size(1) == 9, "groups must be int64 [G][9]");
    static_assert(sizeof(mb::LinearGroup) == 9 * 8, "LinearGroup layout");
    const int smem = mb::kLinStages * mb::kLinStageBytes + 1024;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(mb::grouped_linear_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr = true; }
    dim3 grid((unsigned)((N + mb::kLinBN - 1) / mb::kLinBN), (unsigned)((max_m + mb::kLinBM - 1) / mb::kLinBM), (unsigned)G);
    mb::grouped_linear_tf32_kernel<<<grid, mb::kLinThreads, smem, at::cuda::getCurrentCUDAStream()>>>(
        reinterpret_cast<const mb::LinearGroup*>(groups.data_ptr<int64_t>()), (int)K, (int)N, (int)ldx, (int)ldy, (int)act, (float)eps);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// groups: int64 [G][3] = {out, targets, M};  stats [G][8] is accumulated into (zero it first)
void grouped_eval(Tensor groups, int64_t G, int64_t max_m, int64_t C, int64_t ld, bool dirichlet, Tensor stats) {
    if (G == 0 || max_m == 0) return;
    c10::cuda::CUDAGuard guard(groups.device());
    TORCH_CHECK(groups.dtype() == torch::kInt64 && groups.is_contiguous() && groups.size(1) == 3);
    static_assert(sizeof(mb::EvalGroup) == 3 * 8, "EvalGroup layout");
    dim3 grid((unsigned)((max_m * 32 + 255) / 256), (unsigned)G);
    mb::grouped_eval_kernel<<<grid, 256, 0, at::cuda::getCurrentCUDAStream()>>>(
        reinterpret_cast<const mb::EvalGroup*>(groups.data_ptr<int64_t>()), (int)C, (int)ld, dirichlet ? 1 : 0, stats.data_ptr<float>());
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}
