// murmura_b200 — grouped implicit-GEMM convolution / linear layers on 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
// This is the training hot loop of a federated node (reference murmura/core/node.py:59-109: forward → loss → backward →
// SGD through stock autograd, i.e. cuDNN / cuBLAS library kernels) and the foreign-weight forward of UBAR stage 2 /
// EvidentialTrust / DMTT scoring (murmura/aggregation/ubar.py:152-202, murmura/dmtt/node_process.py:309-363) as ONE
// kernel family.  A launch processes the same layer of EVERY virtual node hosted on the GPU (blockIdx.z = node × split-K
// slice); weights, biases and BatchNorm statistics are read (and, in W mode, updated) IN PLACE in the nodes' arena rows.
//
//   mode F (fprop)  Y[p][co]  = Σ_{tap,ci} X[pix(p,tap)][ci] · W[co][tap][ci]      (+bias, eval-BN, residual, ReLU / softplus+1)
//   mode D (dgrad)  dX[q][ci] (+)= Σ_{tap,co} dY[pix'(q,tap)][co] · W[co][tap][ci]
//   mode W (wgrad)  W[co][tap][ci] += α · Σ_p X[pix(p,tap)][ci] · dY[p][co]          (α = −lr: the SGD step is the epilogue)
//
// Activations are NHWC fp32, weights (Cout, KH, KW, Cin) fp32 — the physical layout of the arena.  Operand tiles are
// gathered by cp.async (16 B, zero-fill for padding / out-of-range taps) straight into 128-byte-swizzled shared memory:
// K-major tiles (SWIZZLE_128B) for im2col rows and fprop weights, MN-major tiles (SWIZZLE_128B_BASE32B, the tf32 transposing
// layout) for dgrad weights and both wgrad operands — their reduction index is the slow axis in memory, so no
// transposition pass is ever needed.  One elected thread issues
// `tcgen05.mma.kind::tf32` (fp32 operands consumed in place, fp32 accumulation in TMEM); a 4-stage mbarrier ring
// decouples the 4 loader warps from the MMA warp; the loader warps then run the epilogue out of TMEM (`tcgen05.ld`).
// Only taps that touch at least one real pixel are visited ("live taps": a 3×3 conv on a 1×1 map is a 1×1 conv).
// Split-K slices reduce with `red.global.add` (the SGD update of W mode is a sum anyway).
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAException.h>
#include <pybind11/pybind11.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "conv_common.cuh"
#include "pdl.cuh"

namespace py = pybind11;

namespace mb {

// Source pixel of GEMM row (b, y, x) for tap (kh, kw).  F / W: the input pixel under the tap; D: the output pixel whose
// tap (kh, kw) lands on input pixel (y, x) (exists only when the offset is a multiple of the stride).
template <int MODE>
__device__ __forceinline__ bool gather_pixel(const ConvGemmParams& p, int pk, int kh, int kw, int& pix) {
    const int b = pk >> 16, y = (pk >> 8) & 255, x = pk & 255;
    int sy, sx;
    if (MODE == kModeD) {
        const int ty = y + p.pad - kh, tx = x + p.pad - kw;
        if (ty < 0 || tx < 0) return false;
        if (p.stride == 1) { sy = ty; sx = tx; }
        else {
            sy = ty / p.stride; sx = tx / p.stride;
            if (sy * p.stride != ty || sx * p.stride != tx) return false;
        }
    } else {
        sy = y * p.stride + kh - p.pad; sx = x * p.stride + kw - p.pad;
        if (sy < 0 || sx < 0) return false;
    }
    if (sy >= p.SH || sx >= p.SW) return false;
    pix = (b * p.SH + sy) * p.SW + sx;
    return true;
}

// ---- operand A, K-major (F: im2col rows of X, D: gathered rows of dY); tile = 128 rows × 32 k, 128B swizzle ---------------
template <int MODE>
__device__ __forceinline__ void load_a_rows(const ConvGemmParams& p, const float* __restrict__ Xg, uint32_t sa, int kb,
                                            const int (&pk)[8], int tid) {
    const int c = tid & 7, r0 = tid >> 3;
    const int k = kb * kCgBK + c * 4;
    const bool kval = k < p.K;
    int cc = 0, kh = 0, kw = 0;
    if (kval) {
        const int lt = k / p.C; cc = k - lt * p.C;
        const int tap = p.taps[lt]; kh = tap / p.KW; kw = tap - kh * p.KW;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int r = r0 + 16 * j;
        const uint32_t dst = sa + r * 128 + ((c ^ (r & 7)) << 4);
        int pix = 0;
        const bool ok = kval && pk[j] >= 0 && gather_pixel<MODE>(p, pk[j], kh, kw, pix);
        cp_async16(dst, ok ? Xg + (long long)pix * p.lds + cc : Xg, ok ? 16 : 0);
    }
}

// ---- operand B of mode F, K-major: weight rows W[n][tap][ci]; tile = BN rows × 32 k ------------------------------------------
template <int BN>
__device__ __forceinline__ void load_b_weights_k(const ConvGemmParams& p, const float* __restrict__ Wg, uint32_t sb, int kb, int n0, int tid) {
    if (p.vecB == 4) {
        const int c = tid & 7, r0 = tid >> 3;
        const int k = kb * kCgBK + c * 4;
        const bool kval = k < p.K;
        int koff = 0;
        if (kval) { const int lt = k / p.C; koff = (int)p.taps[lt] * p.Cw_real + (k - lt * p.C); }
#pragma unroll
        for (int j = 0; j < BN / 16; ++j) {
            const int r = r0 + 16 * j, n = n0 + r;
            const uint32_t dst = sb + r * 128 + ((c ^ (r & 7)) << 4);
            const bool ok = kval && n < p.N;
            cp_async16(dst, ok ? Wg + (long long)n * p.wrow + koff : Wg, ok ? 16 : 0);
        }
    } else {                                     // rows not 16-byte aligned (first layers: K = 561, 7·7·3, 5·5·1 …): 4-byte copies
#pragma unroll 4
        for (int i = 0; i < BN * kCgBK / kCgLoaders; ++i) {
            const int e = tid + kCgLoaders * i;
            const int r = e >> 5, kk = e & 31, n = n0 + r;
            const int k = kb * kCgBK + kk;
            const uint32_t dst = sb + r * 128 + ((((kk >> 2) ^ (r & 7)) << 4) | ((kk & 3) << 2));
            bool ok = k < p.K && n < p.N;
            int koff = 0;
            if (ok) { const int lt = k / p.C, cc = k - lt * p.C; ok = cc < p.Cw_real; koff = (int)p.taps[lt] * p.Cw_real + cc; }
            cp_async4(dst, ok ? Wg + (long long)n * p.wrow + koff : Wg, ok ? 4 : 0);
        }
    }
}

// MN-major tile [8 k-groups of 4][ROWS/32 atoms][4 k-rows][128 B] (SWIZZLE_128B_BASE32B): address of the 16-byte chunk `c`
// (4 consecutive M/N rows) of k-row `j` — the 32-byte chunk index is XOR-ed with (j % 4).
template <int ROWS>
__device__ __forceinline__ uint32_t mn_chunk_addr(uint32_t base, int j, int c) {
    const int r = j & 3, c16 = c & 7;
    return base + (((j >> 2) * (ROWS / 32) + (c >> 3)) << 9) + (r << 7) + ((((c16 >> 1) ^ r)) << 5) + ((c16 & 1) << 4);
}

// ---- operand B of mode D, MN-major: W[co][tap][ci] with k = (tap, co), rows n = ci ------------------------------------------
template <int BN>
__device__ __forceinline__ void load_b_weights_mn(const ConvGemmParams& p, const float* __restrict__ Wg, uint32_t sb, int kb, int n0, int tid) {
    constexpr int CPR = BN / 4, KSTEP = kCgLoaders / CPR, ITER = kCgBK / KSTEP;
    const int c = tid % CPR, j0 = tid / CPR;
    const int n = n0 + c * 4;
    const bool nval = n < p.N;
#pragma unroll
    for (int i = 0; i < ITER; ++i) {
        const int j = j0 + KSTEP * i;
        const int k = kb * kCgBK + j;
        const bool ok = nval && k < p.K;
        long long off = 0;
        bool ok2 = ok;
        if (ok) { const int lt = k / p.C, co = k - lt * p.C; ok2 = co < p.Ck_real; if (ok2) off = (long long)co * p.wrow + (int)p.taps[lt] * p.Cw_real + n; }
        cp_async16(mn_chunk_addr<BN>(sb, j, c), Wg + off, ok2 ? 16 : 0);
    }
}

// ---- operand A of mode W, MN-major: im2col of X with k = output pixel, rows m = (tap, ci) (+ the all-ones bias row) -------------
__device__ __forceinline__ void load_a_cols(const ConvGemmParams& p, const float* __restrict__ Xg, uint32_t sa, int kb, int m0, int tid) {
    const int c = tid & 31, j0 = tid >> 5;
    const int m = m0 + c * 4;
    const int Mreal = p.ntaps * p.C;
    int kind = 0, cc = 0, kh = 0, kw = 0;                    // 0 = zero rows, 1 = im2col rows, 2 = ones (bias gradient)
    if (m < Mreal) { const int lt = m / p.C; cc = m - lt * p.C; const int tap = p.taps[lt]; kh = tap / p.KW; kw = tap - kh * p.KW; kind = 1; }
    else if (m == Mreal && p.ones_row) kind = 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int j = j0 + 4 * i;
        const int pidx = kb * kCgBK + j;
        bool ok = kind != 0 && pidx < p.K;
        const float* src = Xg;
        if (ok) {
            if (kind == 2) src = p.ones;
            else {
                int pix = 0;
                ok = gather_pixel<kModeW>(p, __ldg(p.ptab + pidx), kh, kw, pix);
                if (ok) src = Xg + (long long)pix * p.lds + cc;
            }
        }
        cp_async16(mn_chunk_addr<kCgBM>(sa, j, c), src, ok ? 16 : 0);
    }
}

// ---- operand B of mode W, MN-major: dY[p][co] with k = output pixel, rows n = co ----------------------------------------------
template <int BN>
__device__ __forceinline__ void load_b_grad(const ConvGemmParams& p, const float* __restrict__ Yg, uint32_t sb, int kb, int n0, int tid) {
    constexpr int CPR = BN / 4, KSTEP = kCgLoaders / CPR, ITER = kCgBK / KSTEP;
    const int c = tid % CPR, j0 = tid / CPR;
    const int n = n0 + c * 4;
    const bool nval = n < p.N;                                 // the padding columns [N, ldy) of dY are zero
#pragma unroll
    for (int i = 0; i < ITER; ++i) {
        const int j = j0 + KSTEP * i;
        const int pidx = kb * kCgBK + j;
        const bool ok = nval && pidx < p.K;
        cp_async16(mn_chunk_addr<BN>(sb, j, c), ok ? Yg + (long long)pidx * p.ldy + n : Yg, ok ? 16 : 0);
    }
}

__device__ __forceinline__ long long gtimer() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define CG_STAMP(slot) do { if (p.dbg) p.dbg[(((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 + (slot)] = gtimer(); } while (0)

template <int MODE, int BN>
__global__ void __launch_bounds__(kCgThreads) conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
    constexpr int NS = kCgStages, LOOK = kCgStages - 1;
    constexpr int STAGE = kCgABytes + BN * 128;
    constexpr bool A_MN = MODE == kModeW, B_MN = MODE != kModeF;
    extern __shared__ uint8_t cg_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(cg_smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[NS];
    __shared__ __align__(8) uint64_t empty_bar[NS];
    __shared__ __align__(8) uint64_t accum_bar;
    __shared__ uint32_t tmem_base_smem;
    __shared__ float bn_ss[2 * BN];

    pdl_launch_dependents();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = (int)blockIdx.z / p.splitk, split = (int)blockIdx.z - g * p.splitk;
    const int m0 = (int)blockIdx.x * kCgBM, n0 = (int)blockIdx.y * BN;
    const int kb_begin = split * p.kb_per_split;
    const int nkb = min(p.kb_total, kb_begin + p.kb_per_split) - kb_begin;       // ≥ 1 by construction of the split

    if (warp == 4) {
        if (lane == 0) {
            for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], kCgLoaders); mbar_init(&empty_bar[s], 1); }
            mbar_init(&accum_bar, 1);
            mbar_init_fence();
        }
        __syncwarp();
        tmem_alloc<BN>(&tmem_base_smem);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    pdl_wait();                                 // predecessor grid complete + flushed: global memory may be touched from here on
    if (tid == 0) { CG_STAMP(0); CG_STAMP(1); }

    float* row = p.row_tab ? reinterpret_cast<float*>(p.row_tab[g]) : p.arena + (long long)(p.gmap ? p.gmap[g] : g) * p.arena_gs;
    float* Wg = row + p.w_off;
    const float* Xg = p.X + (long long)g * p.x_gs;
    float* Yg = p.Y + (long long)g * p.y_gs;
    const uint32_t smem0 = smem_u32(smem);

    if (warp < 4) {
        // ================= loaders: cp.async ring, LOOK stages in flight =================
        int pk[8];
        if (MODE != kModeW) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { const int r = m0 + (tid >> 3) + 16 * j; pk[j] = r < p.M ? __ldg(p.ptab + r) : -1; }
        }
        for (int it = 0; it < nkb + LOOK; ++it) {
            if (it < nkb) {
                const int s = it % NS;
                if (it >= NS) mbar_wait(&empty_bar[s], (uint32_t)((it / NS - 1) & 1));      // MMAs that read this slot have retired
                const uint32_t sa = smem0 + s * STAGE, sb = sa + kCgABytes;
                const int kb = kb_begin + it;
                if (MODE == kModeW) { load_a_cols(p, Xg, sa, kb, m0, tid); load_b_grad<BN>(p, Yg, sb, kb, n0, tid); }
                else {
                    load_a_rows<MODE>(p, Xg, sa, kb, pk, tid);
                    if (MODE == kModeF) load_b_weights_k<BN>(p, Wg, sb, kb, n0, tid);
                    else load_b_weights_mn<BN>(p, Wg, sb, kb, n0, tid);
                }
            }
            cp_async_commit();
            if (tid == 0 && it == 0) CG_STAMP(2);
            if (it >= LOOK) {
                cp_async_wait<LOOK>();                          // this thread's copies of k-block it−LOOK have landed
                fence_proxy_async();
                mbar_arrive(&full_bar[(it - LOOK) % NS]);
            }
        }
        // ================= epilogue: TMEM → registers → global =================
        if (tid == 0) CG_STAMP(3);
        if (MODE != kModeW) epilogue_prepare_bn<BN>(p, row, n0, tid, bn_ss);
        mbar_wait_backoff(&accum_bar, 0);
        tc_fence_after();
        if (tid == 0) CG_STAMP(5);
        const int r = m0 + warp * 32 + lane;
        if (MODE != kModeW) epilogue_rows<BN>(p, tmem_base, warp, r < p.M ? (long long)r : -1ll, n0, g, split, row, Yg, bn_ss);
        else epilogue_wgrad<BN>(p, tmem_base, warp, r, n0, row);
    } else if (lane == 0) {
        // ================= MMA issuer: 4 × (128 × BN × 8) tf32 MMAs per k-block =================
        constexpr uint32_t idesc = umma_idesc_tf32(kCgBM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
        for (int it = 0; it < nkb; ++it) {
            const int s = it % NS;
            mbar_wait(&full_bar[s], (uint32_t)((it / NS) & 1));
            tc_fence_after();
            if (it == 0) CG_STAMP(4);
            const uint32_t sa = smem0 + s * STAGE, sb = sa + kCgABytes;
#pragma unroll
            for (int k = 0; k < kCgBK / 8; ++k) {
                // MN-major (BASE32B): atoms along M/N are 512 B apart (LBO), the next 4 k's are (rows/32)·512 B apart (SBO);
                // one UMMA_K = 8 step = two k-groups
                const uint32_t a_lbo = p.mn_swap ? (kCgBM / 32) * 512 : 512, a_sbo = p.mn_swap ? 512 : (kCgBM / 32) * 512;
                const uint32_t b_lbo = p.mn_swap ? (BN / 32) * 512 : 512, b_sbo = p.mn_swap ? 512 : (BN / 32) * 512;
                const uint64_t ad = A_MN ? umma_desc(sa + k * (kCgBM / 32) * 1024, a_lbo, a_sbo, kLayoutSw128Base32)
                                         : umma_desc_sw128(sa + k * 32, 0, 1024);
                const uint64_t bd = B_MN ? umma_desc(sb + k * (BN / 32) * 1024, b_lbo, b_sbo, kLayoutSw128Base32)
                                         : umma_desc_sw128(sb + k * 32, 0, 1024);
                umma_tf32(tmem_base, ad, bd, idesc, (it > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&empty_bar[s]);
        }
        umma_commit(&accum_bar);
    }
    if (tid == 0) CG_STAMP(6);
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<BN>(tmem_base);
    if (tid == 0) CG_STAMP(7);
}

}  // namespace mb

namespace {

template <int MODE, int BN>
void launch_conv(const mb::ConvGemmParams& p, dim3 grid, cudaStream_t stream) {
    constexpr int smem = mb::kCgStages * (mb::kCgABytes + BN * 128) + 1024;
    static bool attr = false;
    if (!attr) {
        C10_CUDA_CHECK(cudaFuncSetAttribute(mb::conv_gemm_kernel<MODE, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    C10_CUDA_CHECK(mbhost::launch(mb::conv_gemm_kernel<MODE, BN>, grid, dim3(mb::kCgThreads), smem, stream, dim3(1, 1, 1), p));
}

}  // namespace

// One grouped layer launch; `d` is the plan built by murmura_b200/ops/conv_plan.py (field names = ConvGemmParams).
// Pointers travel as integers (sub-buffers of the trainer's workspace / the arena); returns the number of CTAs.
int64_t conv_gemm(py::dict d) {
    mb::ConvGemmParams p;
    int mode, G, bn;
    mbhost::fill_conv_params(d, p, mode, G, bn);
    TORCH_CHECK(p.stride >= 1 && p.SH <= 256 && p.SW <= 256 && p.ptab != nullptr);
    TORCH_CHECK(p.vecB == 4 || (mode == mb::kModeF), "conv_gemm: 4-byte weight copies exist for fprop only");
    if (mode == mb::kModeW) TORCH_CHECK(!p.ones_row || p.ones != nullptr, "conv_gemm: ones buffer missing");
    const int m_ext = mode == mb::kModeW ? p.M + (p.ones_row ? 1 : 0) : p.M;
    dim3 grid((unsigned)((m_ext + mb::kCgBM - 1) / mb::kCgBM), (unsigned)((p.N + bn - 1) / bn), (unsigned)(G * p.splitk));
    TORCH_CHECK(grid.z <= 65535 && grid.y <= 65535, "conv_gemm: grid too large");
    auto stream = at::cuda::getCurrentCUDAStream().stream();
    if (mode == mb::kModeF) { if (bn == 64) launch_conv<mb::kModeF, 64>(p, grid, stream); else launch_conv<mb::kModeF, 128>(p, grid, stream); }
    else if (mode == mb::kModeD) { if (bn == 64) launch_conv<mb::kModeD, 64>(p, grid, stream); else launch_conv<mb::kModeD, 128>(p, grid, stream); }
    else { if (bn == 64) launch_conv<mb::kModeW, 64>(p, grid, stream); else launch_conv<mb::kModeW, 128>(p, grid, stream); }
    return (int64_t)grid.x * grid.y * grid.z;
}
