// murmura_b200 — TMA-fed variant of the grouped implicit-GEMM conv / linear kernels (TMA → swizzled smem → tcgen05 → TMEM), sm_100a.
//
// Same math and epilogues as conv_tcgen05.cu (modes F / D / W, see there), but the operand tiles are produced by ONE thread
// issuing `cp.async.bulk.tensor` instructions instead of 128 threads computing im2col addresses:
//
//   * activations are 5-D tensors {C, W, H, B, node}; the im2col tile of a tap is the box {32 channels, row width, rows, images}
//     at the tap's offset — borders come from TMA's out-of-bounds zero fill, strided convolutions from `elementStrides`;
//     GEMM M tiles are whole images (small maps) or strips of image rows, so a tile is exactly one box;
//   * weights are {KH·KW·Cin, Cout, node} (fprop, K-major) or {Cin, tap, Cout, node} (dgrad, MN-major) tensors over the arena;
//   * wgrad reduces over boxes of ≤ 32 output pixels: X-window boxes (A) and dY boxes (B) land MN-major
//     (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B = the UMMA SWIZZLE_128B_BASE32B layout tf32 needs for transposed operands).
//
// Warp roles (192 threads): warps 0-3 epilogue (TMEM → registers → global / red.add), warp 4 TMEM allocator + MMA issuer,
// warp 5 TMA producer.  4-stage mbarrier ring (expect_tx / tcgen05.commit).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAException.h>
#include <cuda.h>
#include "conv_common.cuh"
#include "pdl.cuh"

namespace py = pybind11;

namespace mb {

constexpr int kCtThreads = 192;

// One tap set of modes F / D.  A stride-s dgrad is s² independent stride-1 problems, one per parity class (py, px) of the input
// pixel: only the taps with kh ≡ py + pad (mod s) reach that class, and they read dY at a constant offset (dx, dy).
struct TapClass { int ntaps, py, px; unsigned char tap[32]; signed char dx[32], dy[32]; };

struct alignas(64) ConvTmaParams {
    CUtensorMap mapA;               // F: X {C,W,H,B,G};  D: dY {C,W,H,B,G};  W: X {C,W,H,B,G} (pixel-block box)
    CUtensorMap mapB;               // F: weights {K,Cout,S};  D: weights {Cin,T,Cout,S};  W: dY {C,W,H,B,G}
    ConvGemmParams g;
    int RH, RW;                     // F / D: plane of the GEMM rows of one class (F: output plane, D: input plane / stride)
    int Bt, TH, tpi, RT;            // tile = Bt whole images (tpi == 1) or a strip of TH rows (tpi strips per image); RT rows
    int ncls, mt_per_cls;           // classes share the grid: blockIdx.x = class · mt_per_cls + tile
    int osc, OPH, OPW;              // output pixel of class row (y, x) = (y·osc + py, x·osc + px) in the [OPH, OPW] plane
    int ystep;                      // source row of tile row y0 = y0·ystep + dy (F: conv stride, D: 1)
    int nB;                         // images per node
    int PK, bh, bb, bpi;            // W: pixels per k-block = RW·bh·bb; k-blocks per image when bb == 1
    int prefill;                    // W: zero the stages before the first load (PK < 32) / write the all-ones bias atom
    int g_off;                      // first group of this launch (scoring launches one group range per source GPU)
    TapClass cls[4];
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const void* map, int c0, int c1, int c2, int c3, int c4, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d_u(uint32_t dst, const void* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

template <int MODE, int BN>
__global__ void __launch_bounds__(kCtThreads) conv_tma_kernel(const __grid_constant__ ConvTmaParams P) {
    constexpr int NS = kCgStages;
    constexpr int STAGE = kCgABytes + BN * 128;
    constexpr bool A_MN = MODE == kModeW, B_MN = MODE != kModeF;
    const ConvGemmParams& p = P.g;
    extern __shared__ uint8_t ct_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ct_smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[NS];
    __shared__ __align__(8) uint64_t empty_bar[NS];
    __shared__ __align__(8) uint64_t accum_bar;
    __shared__ uint32_t tmem_base_smem;
    __shared__ float bn_ss[2 * BN];

    pdl_launch_dependents();                    // the next kernel's prologue may overlap this one (it blocks in pdl_wait)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gl = (int)blockIdx.z / p.splitk, split = (int)blockIdx.z - gl * p.splitk;
    const int g = gl + P.g_off;
    const int n0 = (int)blockIdx.y * BN;
    const int kb_begin = split * p.kb_per_split;
    const int ci = MODE == kModeW ? 0 : (int)blockIdx.x / P.mt_per_cls;             // tap class of this CTA (F / D)
    const TapClass& cls = P.cls[ci];
    const int kb_total = MODE == kModeW ? p.kb_total : cls.ntaps * (p.C / kCgBK);
    const int nkb = max(0, min(kb_total, kb_begin + p.kb_per_split) - kb_begin);    // 0: nothing to add for this slice
    const uint32_t smem0 = smem_u32(smem);

    if (warp == 4) {
        if (lane == 0) {
            for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
            mbar_init(&accum_bar, 1);
            mbar_init_fence();
        }
        __syncwarp();
        tmem_alloc<BN>(&tmem_base_smem);
    } else if (warp == 5 && lane == 0) {
        tma_prefetch_desc(&P.mapA);
        tma_prefetch_desc(&P.mapB);
    }
    if (MODE == kModeW && P.prefill) {
        // rows PK..31 of every MN-major tile are never written by TMA: they must read as zero; the bias-gradient atom is all ones
        float4* s4 = reinterpret_cast<float4*>(smem);
        for (int i = tid; i < NS * STAGE / 16; i += kCtThreads) s4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        const int Mreal = p.ntaps * p.C, m0 = (int)blockIdx.x * kCgBM;
        if (p.ones_row && Mreal >= m0 && Mreal < m0 + kCgBM) {
            const int a = (Mreal - m0) >> 5;
            for (int s = 0; s < NS; ++s) {
                float* atom = reinterpret_cast<float*>(smem + s * STAGE + a * 4096);
                for (int i = tid; i < P.PK * 32; i += kCtThreads) atom[i] = 1.f;
            }
        }
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    pdl_wait();                                 // predecessor grid complete + flushed: global memory may be touched from here on
    const int slot = p.gmap ? p.gmap[g] : g;

    // tile origin in the row space of modes F / D
    int b0 = 0, y0 = 0;
    const int mt = MODE == kModeW ? 0 : (int)blockIdx.x - ci * P.mt_per_cls;
    if (MODE != kModeW) {
        if (P.tpi == 1) b0 = mt * P.Bt;
        else { b0 = mt / P.tpi; y0 = (mt - b0 * P.tpi) * P.TH; }
    }

    if (warp == 5) {
        if (lane == 0) {
            // ================= TMA producer =================
            const int Mreal = p.ntaps * p.C, m0 = (int)blockIdx.x * kCgBM;
            for (int it = 0; it < nkb; ++it) {
                const int s = it % NS;
                if (it >= NS) mbar_wait(&empty_bar[s], (uint32_t)((it / NS - 1) & 1));
                const uint32_t sa = smem0 + s * STAGE, sb = sa + kCgABytes;
                const int kb = kb_begin + it;
                if (MODE != kModeW) {
                    const int k = kb * kCgBK;
                    const int lt = k / p.C, c0 = k - lt * p.C;
                    const int tap = cls.tap[lt];
                    const int xs = cls.dx[lt], ys = y0 * P.ystep + cls.dy[lt];
                    mbar_expect_tx(&full_bar[s], (uint32_t)(P.RT * 128 + BN * 128));
                    tma_load_5d(sa, &P.mapA, c0, xs, ys, b0, g, &full_bar[s]);
                    if (MODE == kModeF) tma_load_3d_u(sb, &P.mapB, tap * p.Cw_real + c0, n0, slot, &full_bar[s]);
                    else {
#pragma unroll
                        for (int a = 0; a < BN / 32; ++a) tma_load_4d(sb + a * 4096, &P.mapB, n0 + 32 * a, tap, c0, slot, &full_bar[s]);
                    }
                } else {
                    int pb0, py0;
                    if (P.bpi == 1) { pb0 = kb * P.bb; py0 = 0; }
                    else { pb0 = kb / P.bpi; py0 = (kb - pb0 * P.bpi) * P.bh; }
                    int na = 0;
#pragma unroll
                    for (int a = 0; a < 4; ++a) na += (m0 + 32 * a < Mreal) ? 1 : 0;
                    mbar_expect_tx(&full_bar[s], (uint32_t)((na + BN / 32) * P.PK * 128));
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const int m = m0 + 32 * a;
                        if (m < Mreal) {
                            const int lt = m / p.C, ci0 = m - lt * p.C;
                            const int tap = p.taps[lt], kh = tap / p.KW, kw = tap - kh * p.KW;
                            tma_load_5d(sa + a * 4096, &P.mapA, ci0, kw - p.pad, py0 * p.stride + kh - p.pad, pb0, g, &full_bar[s]);
                        }
                    }
#pragma unroll
                    for (int a = 0; a < BN / 32; ++a) tma_load_5d(sb + a * 4096, &P.mapB, n0 + 32 * a, 0, py0, pb0, g, &full_bar[s]);
                }
            }
        }
    } else if (warp == 4) {
        if (lane == 0) {
            // ================= MMA issuer =================
            constexpr uint32_t idesc = umma_idesc_tf32(kCgBM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
            const uint32_t mn_lbo = p.mn_swap ? 512 : 4096, mn_sbo = p.mn_swap ? 4096 : 512;   // MN tiles: [atom][32 k-rows][128 B]
            for (int it = 0; it < nkb; ++it) {
                const int s = it % NS;
                mbar_wait(&full_bar[s], (uint32_t)((it / NS) & 1));
                tc_fence_after();
                const uint32_t sa = smem0 + s * STAGE, sb = sa + kCgABytes;
#pragma unroll
                for (int k = 0; k < kCgBK / 8; ++k) {
                    const uint64_t ad = A_MN ? umma_desc(sa + k * 1024, mn_lbo, mn_sbo, kLayoutSw128Base32) : umma_desc_sw128(sa + k * 32, 0, 1024);
                    const uint64_t bd = B_MN ? umma_desc(sb + k * 1024, mn_lbo, mn_sbo, kLayoutSw128Base32) : umma_desc_sw128(sb + k * 32, 0, 1024);
                    umma_tf32(tmem_base, ad, bd, idesc, (it > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);
            }
            if (nkb > 0) umma_commit(&accum_bar);
        }
    } else {
        // ================= epilogue =================
        float* row = p.arena + (long long)slot * p.arena_gs;
        if (MODE != kModeW) epilogue_prepare_bn<BN>(p, row, n0, tid, bn_ss);
        if (nkb > 0) {
            mbar_wait_backoff(&accum_bar, 0);
            tc_fence_after();
            const int r = warp * 32 + lane;
            if (MODE != kModeW) {
                long long grow = -1;
                if (r < P.RT) {
                    const int hw = P.RH * P.RW;
                    const int bi = r / hw, rem = r - bi * hw;                 // whole-image tiles: bi-th image of the tile; strips: bi = 0
                    const int y = y0 + rem / P.RW, x = rem - (rem / P.RW) * P.RW;
                    const int b = b0 + bi;
                    if (y < P.RH && b < P.nB)
                        grow = ((long long)b * P.OPH + (y * P.osc + cls.py)) * P.OPW + (x * P.osc + cls.px);
                }
                epilogue_rows<BN>(p, tmem_base, warp, grow, n0, g, split, row, p.Y + (long long)g * p.y_gs, bn_ss);
            } else {
                epilogue_wgrad<BN>(p, tmem_base, warp, (int)blockIdx.x * kCgBM + r, n0, row);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<BN>(tmem_base);
}

}  // namespace mb

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
        TORCH_CHECK(e == cudaSuccess && q == cudaDriverEntryPointSuccess && ptr, "cuTensorMapEncodeTiled unavailable");
        fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

template <int MODE, int BN>
void launch_tma(const mb::ConvTmaParams& P, dim3 grid, cudaStream_t stream) {
    constexpr int smem = mb::kCgStages * (mb::kCgABytes + BN * 128) + 1024;
    static bool attr = false;
    if (!attr) {
        C10_CUDA_CHECK(cudaFuncSetAttribute(mb::conv_tma_kernel<MODE, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    C10_CUDA_CHECK(mbhost::launch(mb::conv_tma_kernel<MODE, BN>, grid, dim3(mb::kCtThreads), smem, stream, dim3(1, 1, 1), P));
}
}  // namespace

// fp32 tiled tensor map (rank ≤ 5): dims / box in elements (innermost first), strides in bytes for dims 1..rank-1.
// swizzle: 0 none, 3 = 128B (K-major UMMA tiles), 4 = 128B_ATOM_32B (MN-major tf32 tiles).  Returns the 128 descriptor bytes.
py::bytes tma_encode(int64_t ptr, std::vector<int64_t> dims, std::vector<int64_t> strides_bytes, std::vector<int64_t> box,
                     std::vector<int64_t> elem_strides, int64_t swizzle) {
    const int rank = (int)dims.size();
    TORCH_CHECK(rank >= 1 && rank <= 5 && (int)strides_bytes.size() == rank - 1 && (int)box.size() == rank && (int)elem_strides.size() == rank);
    cuuint64_t gdim[5], gstr[4]; cuuint32_t b[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = (cuuint64_t)dims[i]; b[i] = (cuuint32_t)box[i]; es[i] = (cuuint32_t)elem_strides[i];
        TORCH_CHECK(box[i] >= 1 && box[i] <= 256 && elem_strides[i] >= 1 && elem_strides[i] <= 8, "tma_encode: box / element stride out of range");
    }
    for (int i = 0; i + 1 < rank; ++i) { gstr[i] = (cuuint64_t)strides_bytes[i]; TORCH_CHECK(strides_bytes[i] % 16 == 0, "tma_encode: strides must be multiples of 16 bytes"); }
    TORCH_CHECK(ptr % 16 == 0 && box[0] * 4 <= 128 && (box[0] * 4) % 16 == 0, "tma_encode: alignment");
    CUtensorMap m;
    CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, reinterpret_cast<void*>(ptr), gdim, gstr, b, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TORCH_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code ", (int)r);
    return py::bytes(reinterpret_cast<const char*>(&m), sizeof(m));
}

// TMA-fed launch of a plan (ops/conv_plan.py: tma_* fields + mapA / mapB descriptor bytes).  Returns the number of CTAs.
int64_t conv_tma(py::dict d) {
    mb::ConvTmaParams P;
    memset(&P, 0, sizeof(P));
    int mode, G, bn;
    mbhost::fill_conv_params(d, P.g, mode, G, bn);
    const std::string ma = d["mapA"].cast<std::string>(), mbs = d["mapB"].cast<std::string>();
    TORCH_CHECK(ma.size() == sizeof(CUtensorMap) && mbs.size() == sizeof(CUtensorMap), "conv_tma: tensor-map descriptors must be 128 bytes");
    memcpy(&P.mapA, ma.data(), sizeof(CUtensorMap)); memcpy(&P.mapB, mbs.data(), sizeof(CUtensorMap));
    using mbhost::dget;
    P.RH = dget<int>(d, "RH", 1); P.RW = dget<int>(d, "RW", 1);
    P.Bt = dget<int>(d, "Bt", 1); P.TH = dget<int>(d, "TH", 1); P.tpi = dget<int>(d, "tpi", 1); P.RT = dget<int>(d, "RT", 128);
    P.osc = dget<int>(d, "osc", 1); P.OPH = dget<int>(d, "OPH", P.RH); P.OPW = dget<int>(d, "OPW", P.RW); P.ystep = dget<int>(d, "ystep", 1);
    P.ncls = 1; P.mt_per_cls = dget<int>(d, "mtiles", 1);
    P.nB = P.g.M / std::max(1, P.OPH * P.OPW);
    if (mode != mb::kModeW) {
        auto classes = d["classes"].cast<std::vector<py::dict>>();
        TORCH_CHECK(!classes.empty() && classes.size() <= 4, "conv_tma: 1..4 tap classes");
        P.ncls = (int)classes.size();
        for (int c = 0; c < P.ncls; ++c) {
            auto taps = classes[c]["taps"].cast<std::vector<int>>();
            auto dx = classes[c]["dx"].cast<std::vector<int>>();
            auto dy = classes[c]["dy"].cast<std::vector<int>>();
            TORCH_CHECK(!taps.empty() && taps.size() <= 32 && dx.size() == taps.size() && dy.size() == taps.size(), "conv_tma: 1..32 taps per class");
            P.cls[c].ntaps = (int)taps.size(); P.cls[c].py = classes[c]["py"].cast<int>(); P.cls[c].px = classes[c]["px"].cast<int>();
            for (size_t i = 0; i < taps.size(); ++i) { P.cls[c].tap[i] = (unsigned char)taps[i]; P.cls[c].dx[i] = (signed char)dx[i]; P.cls[c].dy[i] = (signed char)dy[i]; }
        }
        TORCH_CHECK(P.ncls == 1 || P.g.splitk == 1, "conv_tma: class launches do not split K");
    }
    P.g_off = dget<int>(d, "g_off", 0);
    P.PK = dget<int>(d, "PK", 32); P.bh = dget<int>(d, "bh", 1); P.bb = dget<int>(d, "bb", 1); P.bpi = dget<int>(d, "bpi", 1);
    const mb::ConvGemmParams& p = P.g;
    TORCH_CHECK(p.row_tab == nullptr, "conv_tma: per-group row tables are not supported (use conv_gemm)");
    TORCH_CHECK(p.C % 32 == 0 && P.RT >= 1 && P.RT <= 128 && P.PK >= 1 && P.PK <= 32, "conv_tma: geometry");
    dim3 grid;
    if (mode == mb::kModeW) {
        const int Mreal = p.ntaps * p.C;
        P.prefill = (P.PK < 32 || p.ones_row) ? 1 : 0;
        grid = dim3((unsigned)((Mreal + (p.ones_row ? 1 : 0) + mb::kCgBM - 1) / mb::kCgBM), (unsigned)((p.N + bn - 1) / bn), (unsigned)(G * p.splitk));
    } else {
        grid = dim3((unsigned)(P.ncls * P.mt_per_cls), (unsigned)((p.N + bn - 1) / bn), (unsigned)(G * p.splitk));
    }
    TORCH_CHECK(grid.z <= 65535 && grid.y <= 65535, "conv_tma: grid too large");
    auto stream = at::cuda::getCurrentCUDAStream().stream();
    if (mode == mb::kModeF) { if (bn == 64) launch_tma<mb::kModeF, 64>(P, grid, stream); else launch_tma<mb::kModeF, 128>(P, grid, stream); }
    else if (mode == mb::kModeD) { if (bn == 64) launch_tma<mb::kModeD, 64>(P, grid, stream); else launch_tma<mb::kModeD, 128>(P, grid, stream); }
    else { if (bn == 64) launch_tma<mb::kModeW, 64>(P, grid, stream); else launch_tma<mb::kModeW, 128>(P, grid, stream); }
    return (int64_t)grid.x * grid.y * grid.z;
}
