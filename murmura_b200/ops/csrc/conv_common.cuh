// murmura_b200 — declarations shared by the cp.async (conv_tcgen05.cu) and TMA (conv_tma.cu) implicit-GEMM kernels:
// launch parameters, the TMEM → global epilogues and the plan-dict → parameter conversion.
#pragma once
#include <torch/extension.h>
#include <pybind11/pybind11.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "tc_common.cuh"

namespace mb {
using namespace mbtc;

constexpr int kCgBM = 128;                      // UMMA M
constexpr int kCgBK = 32;                       // fp32 per k-block = one 128-byte swizzle row
constexpr int kCgStages = 4;
constexpr int kCgLoaders = 128;                 // warps 0..3: operand gather, then epilogue
constexpr int kCgThreads = 160;                 // warp 4: TMEM allocator + MMA issuer
constexpr int kCgABytes = kCgBM * kCgBK * 4;    // 16 KiB

enum { kModeF = 0, kModeD = 1, kModeW = 2 };

struct ConvGemmParams {
    const float* X; long long x_gs;             // gather source: input activations (F, W) / output gradient (D); group stride (elements)
    float* Y; long long y_gs;                   // F, D: output [M][ldy];  W: the dY operand [P][ldy] (read only)
    const float* R; long long r_gs;             // epilogue operand with the layout of Y: rmode 1 = residual added before the activation (F),
    int rmode;                                  //   rmode 2 = ReLU mask, out *= (R > 0) (D: gradient through a ReLU fused into the producer)
    float* arena; long long arena_gs;           // arena rows: base of group g = row_tab ? row_tab[g] : arena + (gmap ? gmap[g] : g)·arena_gs
    const long long* row_tab; const int* gmap;
    long long w_off, bias_off, bn_mean_off, bn_var_off, bn_gamma_off, bn_beta_off;   // element offsets in a row; < 0 = absent
    const int* ptab;                            // packed (b << 16 | y << 8 | x) per GEMM row (F, D) / per reduction index (W)
    const float* ones;                          // ≥ 4 floats of 1.0 (bias-gradient row of W mode)
    int M, N, K;
    int splitk, kb_total, kb_per_split;
    int SH, SW, C, lds;                         // source plane; k-decode modulus (channels per tap); floats between source pixels
    int KW, stride, pad, ntaps;
    int Cw_real, wrow;                          // weight channels per tap, floats per output-channel row (KH·KW·Cin)
    int Ck_real;                                // D: real output channels (k-decode modulus C may be the padded count)
    int ldy;
    float alpha, eps;
    int relu, act, accumulate, vecB, ones_row;
    long long* dbg;                             // debug: per-CTA phase timestamps (%globaltimer, 8 slots) or null
    int mn_swap;                                // debug: swap the LBO / SBO roles of MN-major descriptors (ops/selfcheck.py probes it)
    unsigned char taps[64];
};


// ---- epilogue of modes F / D: one accumulator row per thread (TMEM lane), 16 columns per tcgen05.ld -------------------------------
// `grow` = global GEMM row of this thread (< 0: nothing to store).  All launch-invariant switches are hoisted out of the
// element loops: with one warp per SM sub-partition every dependent constant load / branch is exposed latency.
// Eval-mode BatchNorm of the output channels [n0, n0 + BN) as one scale / shift pair per column (call with the 128 epilogue
// threads before the accumulator is ready; `ss` = 2·BN floats of shared memory).
template <int BN>
__device__ __forceinline__ void epilogue_prepare_bn(const ConvGemmParams& p, const float* __restrict__ row, int n0, int tid, float* ss) {
    if (p.bn_mean_off < 0) return;
    for (int i = tid; i < BN; i += 128) {
        const int col = n0 + i;
        float sc = 0.f, sh = 0.f;
        if (col < p.N) {
            const float g_ = p.bn_gamma_off >= 0 ? row[p.bn_gamma_off + col] : 1.f;
            const float b_ = p.bn_beta_off >= 0 ? row[p.bn_beta_off + col] : 0.f;
            sc = rsqrtf(row[p.bn_var_off + col] + p.eps) * g_;
            sh = b_ - row[p.bn_mean_off + col] * sc;
        }
        ss[i] = sc; ss[BN + i] = sh;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");          // the four epilogue warps only
}

template <int BN>
__device__ __forceinline__ void epilogue_rows(const ConvGemmParams& p, uint32_t tmem_base, int warp, long long grow, int n0, int g,
                                              int split, float* __restrict__ row, float* __restrict__ Yg, const float* __restrict__ ss) {
    const int N = p.N, ldy = p.ldy;
    const float alpha = p.alpha;
    const bool rvalid = grow >= 0;
    float* yrow = Yg + grow * ldy;
    const bool atomic = p.accumulate || p.splitk > 1;
    const float* bias = (p.bias_off >= 0 && split == 0) ? row + p.bias_off : nullptr;
    const bool bn = p.bn_mean_off >= 0;
    const float* rrow = p.R ? p.R + (long long)g * p.r_gs + grow * ldy : nullptr;
    const bool mask = p.rmode == 2, relu = p.relu != 0, act2 = p.act == 2;
    const bool fused = bn || rrow != nullptr || relu || act2;
    const bool vec_ok = (ldy & 3) == 0;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        const int cbase = n0 + c0;
        if (!rvalid || cbase >= N) continue;
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) * alpha;
        const bool full = cbase + 16 <= N;
        if (bias) {
#pragma unroll
            for (int i = 0; i < 16; ++i) if (full || cbase + i < N) f[i] += bias[cbase + i];
        }
        if (fused) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int col = cbase + i;
                if (full || col < N) {
                    float x = f[i];
                    if (bn) x = fmaf(x, ss[c0 + i], ss[BN + c0 + i]);
                    if (rrow) { if (mask) x = rrow[col] > 0.f ? x : 0.f; else x += rrow[col]; }
                    if (relu) x = relu_keep_nan(x);
                    else if (act2) x = (x > 20.f ? x : log1pf(__expf(x))) + 1.f;       // Dirichlet head: softplus + 1
                    f[i] = x;
                }
            }
        }
        if (full && vec_ok) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                if (atomic) red_add_v4(yrow + cbase + i, f[i], f[i + 1], f[i + 2], f[i + 3]);
                else *reinterpret_cast<float4*>(yrow + cbase + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int col = cbase + i;
                if (col < N) { if (atomic) red_add_f32(yrow + col, f[i]); else yrow[col] = f[i]; }
            }
        }
    }
}

// ---- epilogue of mode W: rows = (live tap, ci) → W[col][tap][ci] += α·acc (32 lanes = 32 consecutive ci: coalesced reductions);
// the all-ones row (m == ntaps·C) lands in the bias -------------------------------------------------------------------------------
template <int BN>
__device__ __forceinline__ void epilogue_wgrad(const ConvGemmParams& p, uint32_t tmem_base, int warp, int r, int n0, float* __restrict__ row) {
    const int Mreal = p.ntaps * p.C, N = p.N;
    const float alpha = p.alpha;
    float* target = nullptr; long long cstride = 0;
    if (r < Mreal) {
        const int lt = r / p.C, cc = r - lt * p.C;
        if (cc < p.Cw_real) { target = row + p.w_off + (int)p.taps[lt] * p.Cw_real + cc; cstride = p.wrow; }
    } else if (r == Mreal && p.ones_row && p.bias_off >= 0) { target = row + p.bias_off; cstride = 1; }
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        const int cbase = n0 + c0;
        if (target == nullptr || cbase >= N) continue;
        float* t = target + (long long)cbase * cstride;
        if (cbase + 16 <= N) {
#pragma unroll
            for (int i = 0; i < 16; ++i) red_add_f32(t + (long long)i * cstride, __uint_as_float(v[i]) * alpha);
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) if (cbase + i < N) red_add_f32(t + (long long)i * cstride, __uint_as_float(v[i]) * alpha);
        }
    }
}

}  // namespace mb

// ---- plan dict → ConvGemmParams (host) -------------------------------------------------------------------------------------------------
namespace mbhost {
namespace py = pybind11;
template <typename T> inline T dget(const py::dict& d, const char* k, T def) { return d.contains(k) ? d[k].cast<T>() : def; }

inline void fill_conv_params(const py::dict& d, mb::ConvGemmParams& p, int& mode, int& G, int& bn) {
    memset(&p, 0, sizeof(p));
    mode = d["mode"].cast<int>();
    G = d["G"].cast<int>();
    bn = dget<int>(d, "BN", 64);
    p.X = reinterpret_cast<const float*>(d["X"].cast<int64_t>()); p.x_gs = d["x_gs"].cast<int64_t>();
    p.Y = reinterpret_cast<float*>(d["Y"].cast<int64_t>()); p.y_gs = d["y_gs"].cast<int64_t>();
    p.R = reinterpret_cast<const float*>(dget<int64_t>(d, "R", 0)); p.r_gs = dget<int64_t>(d, "r_gs", 0); p.rmode = dget<int>(d, "rmode", 1);
    p.arena = reinterpret_cast<float*>(d["arena"].cast<int64_t>()); p.arena_gs = d["arena_gs"].cast<int64_t>();
    p.row_tab = reinterpret_cast<const long long*>(dget<int64_t>(d, "row_tab", 0));
    p.gmap = reinterpret_cast<const int*>(dget<int64_t>(d, "gmap", 0));
    p.w_off = d["w_off"].cast<int64_t>(); p.bias_off = dget<int64_t>(d, "bias_off", -1);
    p.bn_mean_off = dget<int64_t>(d, "bn_mean_off", -1); p.bn_var_off = dget<int64_t>(d, "bn_var_off", -1);
    p.bn_gamma_off = dget<int64_t>(d, "bn_gamma_off", -1); p.bn_beta_off = dget<int64_t>(d, "bn_beta_off", -1);
    p.ptab = reinterpret_cast<const int*>(dget<int64_t>(d, "ptab", 0));
    p.ones = reinterpret_cast<const float*>(dget<int64_t>(d, "ones", 0));
    p.M = d["M"].cast<int>(); p.N = d["N"].cast<int>(); p.K = d["K"].cast<int>();
    p.splitk = dget<int>(d, "splitk", 1);
    p.SH = d["SH"].cast<int>(); p.SW = d["SW"].cast<int>(); p.C = d["C"].cast<int>(); p.lds = d["lds"].cast<int>();
    p.KW = d["KW"].cast<int>(); p.stride = d["stride"].cast<int>(); p.pad = d["pad"].cast<int>();
    p.Cw_real = d["Cw_real"].cast<int>(); p.wrow = d["wrow"].cast<int>(); p.Ck_real = dget<int>(d, "Ck_real", p.C);
    p.ldy = d["ldy"].cast<int>();
    p.alpha = dget<float>(d, "alpha", 1.f); p.eps = dget<float>(d, "eps", 1e-5f);
    p.relu = dget<int>(d, "relu", 0); p.act = dget<int>(d, "act", 0); p.accumulate = dget<int>(d, "accumulate", 0);
    p.vecB = dget<int>(d, "vecB", 4); p.ones_row = dget<int>(d, "ones_row", 0); p.mn_swap = dget<int>(d, "mn_swap", 0);
    p.dbg = reinterpret_cast<long long*>(dget<int64_t>(d, "dbg", 0));
    auto taps = d["taps"].cast<std::vector<int>>();
    TORCH_CHECK(!taps.empty() && taps.size() <= 64, "conv_gemm: 1..64 live taps");
    p.ntaps = (int)taps.size();
    for (size_t i = 0; i < taps.size(); ++i) p.taps[i] = (unsigned char)taps[i];
    TORCH_CHECK(mode >= 0 && mode <= 2 && (bn == 64 || bn == 128) && G >= 1 && p.M > 0 && p.N > 0 && p.K > 0, "conv_gemm: bad plan");
    TORCH_CHECK(p.C % 4 == 0 && p.lds % 4 == 0 && p.ldy % 4 == 0, "conv_gemm: channel counts / leading dimensions must be multiples of 4");
    const bool fused_act = p.relu || p.act || p.bn_mean_off >= 0 || (p.R && p.rmode == 1);
    TORCH_CHECK(!(fused_act || p.R) || p.splitk == 1, "conv_gemm: fused epilogues need the complete sum in one CTA (splitk = 1)");
    TORCH_CHECK(!(fused_act && p.accumulate), "conv_gemm: activations cannot be applied to an accumulating output");
    p.kb_total = dget<int>(d, "kb_total", (p.K + mb::kCgBK - 1) / mb::kCgBK);
    p.splitk = std::max(1, std::min(p.splitk, p.kb_total));
    p.kb_per_split = (p.kb_total + p.splitk - 1) / p.splitk;
    p.splitk = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;                 // no empty slice
}
}  // namespace mbhost
