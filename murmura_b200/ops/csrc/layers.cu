// murmura_b200 — grouped (all virtual nodes of a GPU in ONE launch) non-GEMM layers of the fused local-SGD step, sm_100a.
//
// Together with conv_tcgen05.cu these replace the per-node autograd graph of the reference's hot loop #1
// (murmura/core/node.py:59-109): mini-batch gather, training BatchNorm (+residual)(+ReLU)(+dropout) forward / backward with
// the SGD update of γ/β folded into the backward, max / average pooling, softmax-CE and evidential losses.  blockIdx.z (or
// .y) is the virtual node; parameters live in the nodes' arena rows (`arena + gmap[g]·arena_gs + offset`), activations in
// the trainer's workspace (`base + g·group_stride`).
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAException.h>
#include <pybind11/pybind11.h>
#include "common.cuh"
#include "pdl.cuh"

namespace cg = cooperative_groups;
namespace py = pybind11;

namespace mb {

__device__ __forceinline__ float4 gld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void gst4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// Dropout keep-mask for 4 consecutive channels: one Philox block per float4, keyed by (seed, step, layer, slot).
__device__ __forceinline__ float4 dropout_scale4(uint64_t seed, uint64_t stream, uint64_t idx4, float p_drop) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)idx4, (uint32_t)(idx4 >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const float s = 1.f / (1.f - p_drop);
    return make_float4(u32_to_unit(r.x) > p_drop ? s : 0.f, u32_to_unit(r.y) > p_drop ? s : 0.f,
                       u32_to_unit(r.z) > p_drop ? s : 0.f, u32_to_unit(r.w) > p_drop ? s : 0.f);
}

// =====================================================================================================================
// mini-batch gather (all nodes): xb[g][r][pix][Cdst] = X_g[perm[g][t·eb + r]][pix][Csrc] (zero padded channels), yb likewise
// =====================================================================================================================
struct GatherArgs {
    const long long* x_tab; const long long* y_tab;      // per-group device addresses of the resident shard (fp32 rows / int64 labels)
    const long long* perm; long long perm_ld;             // [G][perm_ld] sample indices of this round
    const int* gmap;                                      // group → row of perm / x_tab (node slot)
    float* xb; long long xb_gs; long long* yb; long long yb_gs;
    long long* rng_step; unsigned int* ticket;            // device step counter (dropout streams), advanced by the last CTA
    int t, eb, npix, Csrc, Cdst;
};

__global__ void __launch_bounds__(256) gather_grouped_kernel(GatherArgs a) {
    pdl_launch_dependents(); pdl_wait();        // PDL: see pdl.cuh (no prologue worth overlapping here, only the launch latency)
    const int g = blockIdx.z, r = blockIdx.y;
    const int slot = a.gmap ? a.gmap[g] : g;
    const long long src = a.perm[(long long)slot * a.perm_ld + (long long)a.t * a.eb + r];
    const float* in = reinterpret_cast<const float*>(a.x_tab[slot]) + src * (long long)a.npix * a.Csrc;
    float* out = a.xb + (long long)g * a.xb_gs + (long long)r * a.npix * a.Cdst;
    const long long total = (long long)a.npix * a.Cdst;
    if (a.Csrc == a.Cdst && (a.Csrc & 3) == 0) {
        const long long n4 = total >> 2;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
            reinterpret_cast<float4*>(out)[i] = reinterpret_cast<const float4*>(in)[i];
    } else {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
            const long long pix = i / a.Cdst; const int c = (int)(i - pix * a.Cdst);
            out[i] = c < a.Csrc ? in[pix * a.Csrc + c] : 0.f;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        a.yb[(long long)g * a.yb_gs + r] = reinterpret_cast<const long long*>(a.y_tab[slot])[src];
    if (a.rng_step != nullptr) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned int total_ctas = gridDim.x * gridDim.y * gridDim.z;
            if (atomicAdd(a.ticket, 1u) == total_ctas - 1) { *a.ticket = 0u; *a.rng_step += 1; }
        }
    }
}

// =====================================================================================================================
// first layer: mini-batch gather fused with im2col (+ per-step packing of the first layer's weights)
//   xcol[g][(r, oh, ow)][k] = X_g[perm[..r]][ih][iw][c],  k = (kh·KW + kw)·Cin + c, zero outside the image and for k >= Kreal
//   wpack[g][co][k] = W_slot[co][k] (k < Kreal, else 0), followed by the bias: rows of Kpad floats (a multiple of 32), so the
//   first convolution / linear layer is a plain TMA-fed GEMM although its real K (7·7·3, 5·5·1, 561 …) is unaligned.
// =====================================================================================================================
struct Im2colArgs {
    const long long* x_tab; const long long* y_tab; const long long* perm; long long perm_ld; const int* gmap;
    float* xcol; long long xcol_gs; long long* yb; long long yb_gs;
    long long* rng_step; unsigned int* ticket;
    int t, eb, IH, IW, Cin, KH, KW, stride, pad, OH, OW, Kreal, Kpad;
    const float* arena; long long arena_gs; const long long* row_tab; const int* wmap; long long w_off, bias_off;
    long long bn_off[4];                                   // eval-mode BatchNorm of the first layer: mean, var, γ, β (or < 0)
    float* wpack; long long wpack_gs; int Cout;
};

__global__ void __launch_bounds__(256) im2col_pack_kernel(Im2colArgs a) {
    pdl_launch_dependents(); pdl_wait();        // PDL: see pdl.cuh (no prologue worth overlapping here, only the launch latency)
    extern __shared__ int ktab[];                            // k → (kh << 24 | kw << 16 | c), −1 for the padding columns
    const int g = blockIdx.z, r = blockIdx.y;
    const int K4 = a.Kpad >> 2;
    if (r < a.eb) {
        for (int k = threadIdx.x; k < a.Kpad; k += blockDim.x) {
            int e = -1;
            if (k < a.Kreal) { const int tap = k / a.Cin, c = k - tap * a.Cin; const int kh = tap / a.KW; e = (kh << 24) | ((tap - kh * a.KW) << 16) | c; }
            ktab[k] = e;
        }
        __syncthreads();
        const int slot = a.gmap ? a.gmap[g] : g;
        const long long src = a.perm[(long long)slot * a.perm_ld + (long long)a.t * a.eb + r];
        const float* __restrict__ in = reinterpret_cast<const float*>(a.x_tab[slot]) + src * (long long)a.IH * a.IW * a.Cin;
        const int P = a.OH * a.OW;
        float* __restrict__ out = a.xcol + (long long)g * a.xcol_gs + (long long)r * P * a.Kpad;
        // a warp walks 32 consecutive k of one output pixel: coalesced 128-byte stores, image reads stay in L1
        const int warps = (gridDim.x * blockDim.x) >> 5, wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
        for (int pix = wid; pix < P; pix += warps) {
            const int oh = pix / a.OW, ow = pix - oh * a.OW;
            const int ih0 = oh * a.stride - a.pad, iw0 = ow * a.stride - a.pad;
            float* orow = out + (long long)pix * a.Kpad;
            for (int k = lane; k < a.Kpad; k += 32) {
                const int e = ktab[k];
                float x = 0.f;
                if (e >= 0) {
                    const int ih = ih0 + (e >> 24), iw = iw0 + ((e >> 16) & 255);
                    if (ih >= 0 && ih < a.IH && iw >= 0 && iw < a.IW) x = in[((long long)ih * a.IW + iw) * a.Cin + (e & 0xFFFF)];
                }
                orow[k] = x;
            }
        }
        if (blockIdx.x == 0 && threadIdx.x == 0)
            a.yb[(long long)g * a.yb_gs + r] = reinterpret_cast<const long long*>(a.y_tab[slot])[src];
    } else if (a.wpack != nullptr) {
        // weight packing block of this group
        const float* row = a.row_tab ? reinterpret_cast<const float*>(a.row_tab[g]) : a.arena + (long long)(a.wmap ? a.wmap[g] : g) * a.arena_gs;
        float* wp = a.wpack + (long long)g * a.wpack_gs;
        const int total = a.Cout * K4;
        for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
            const int co = q / K4, k4 = (q - co * K4) << 2;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = (k4 + j < a.Kreal) ? row[a.w_off + (long long)co * a.Kreal + k4 + j] : 0.f;
            *reinterpret_cast<float4*>(wp + (long long)co * a.Kpad + k4) = make_float4(v[0], v[1], v[2], v[3]);
        }
        // per-channel vectors after the weights: bias, then BatchNorm mean / var / γ / β, ceil4(Cout) floats each
        const int Cp = (a.Cout + 3) & ~3;
        float* vec = wp + (long long)a.Cout * a.Kpad;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.Cout; i += gridDim.x * blockDim.x) {
            if (a.bias_off >= 0) vec[i] = row[a.bias_off + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) if (a.bn_off[j] >= 0) vec[(j + 1) * Cp + i] = row[a.bn_off[j] + i];
        }
    }
    if (a.rng_step != nullptr) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned int total_ctas = gridDim.x * gridDim.y * gridDim.z;
            if (atomicAdd(a.ticket, 1u) == total_ctas - 1) { *a.ticket = 0u; *a.rng_step += 1; }
        }
    }
}

// =====================================================================================================================
// training BatchNorm (+residual) (+ReLU) (+dropout), grouped; the row splits of a channel tile form a thread-block cluster
// and merge their partial sums through distributed shared memory (same scheme as bn_train.cu).
// =====================================================================================================================
constexpr int kGbnThreads = 256;

struct GbnFwdArgs {
    const float* x; const float* res; float* y; long long x_gs, res_gs, y_gs;   // [G][M][C] activations, one group stride each
    float* save_mean; float* save_invstd;                            // [G][C]
    float* arena; long long arena_gs; const int* gmap;               // parameters: arena + gmap[g]·arena_gs + offsets
    long long gamma_off, beta_off, rmean_off, rvar_off;
    long long* nbt; long long nbt_gs; long long nbt_off;             // int64 table [S][nbt_gs], < 0 offset = absent
    const long long* rng_step; unsigned long long seed; int layer_id; float p_drop;
    int M, C; float eps, momentum; int relu;
};

struct GbnBwdArgs {
    const float* dy; const float* x; const float* y; long long dy_gs, x_gs, y_gs;
    float* dx; float* dres; long long dx_gs, dres_gs;                // dres: gradient of the residual branch (or null)
    const float* save_mean; const float* save_invstd;
    float* arena; long long arena_gs; const int* gmap;
    long long gamma_off, beta_off;
    const long long* rng_step; unsigned long long seed; int layer_id; float p_drop;
    int M, C; int relu; float lr;
};

template <int TC>
__device__ __forceinline__ void gbn_cta_reduce_pair(float4 a, float4 b, float (*wpart)[4][8], float* part) {
    constexpr int Q = TC / 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, quad = threadIdx.x & (Q - 1);
#pragma unroll
    for (int o = Q; o < 32; o <<= 1) {
        a.x += __shfl_xor_sync(0xffffffffu, a.x, o); a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
        a.z += __shfl_xor_sync(0xffffffffu, a.z, o); a.w += __shfl_xor_sync(0xffffffffu, a.w, o);
        b.x += __shfl_xor_sync(0xffffffffu, b.x, o); b.y += __shfl_xor_sync(0xffffffffu, b.y, o);
        b.z += __shfl_xor_sync(0xffffffffu, b.z, o); b.w += __shfl_xor_sync(0xffffffffu, b.w, o);
    }
    if (lane < Q) {
        float* w = wpart[warp][quad];
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int which = threadIdx.x >> 4, ch = threadIdx.x & 15;
        float s = 0.f;
        if (ch < TC) {
#pragma unroll
            for (int w = 0; w < kGbnThreads / 32; ++w) s += wpart[w][ch >> 2][which * 4 + (ch & 3)];
        }
        part[threadIdx.x] = s;
    }
}

__device__ __forceinline__ void gbn_cluster_fold(cg::cluster_group& cluster, float* part, float* tot) {
    cluster.sync();
    if (threadIdx.x < 32) {
        float s = 0.f;
        const unsigned n = cluster.num_blocks();
        for (unsigned r = 0; r < n; ++r) s += cluster.map_shared_rank(part, r)[threadIdx.x];
        tot[threadIdx.x] = s;
    }
    cluster.sync();
}

template <int TC>
__global__ void __launch_bounds__(kGbnThreads) gbn_fwd_kernel(GbnFwdArgs a) {
    pdl_launch_dependents(); pdl_wait();        // PDL: see pdl.cuh (no prologue worth overlapping here, only the launch latency)
    constexpr int Q = TC / 4, RL = kGbnThreads / Q;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ float wpart[kGbnThreads / 32][4][8];
    __shared__ float part[32], tot[32], stat[32];
    const int g = blockIdx.z;
    const int slot = a.gmap ? a.gmap[g] : g;
    const float* x = a.x + (long long)g * a.x_gs;
    const float* res = a.res ? a.res + (long long)g * a.res_gs : nullptr;
    float* y = a.y + (long long)g * a.y_gs;
    float* prow = a.arena + (long long)slot * a.arena_gs;
    const int quad = threadIdx.x & (Q - 1), rlane = threadIdx.x / Q;
    const int c = blockIdx.x * TC + quad * 4;
    const bool live = c < a.C;
    const int splits = gridDim.y;
    const int rows_per = (a.M + splits - 1) / splits;
    const int rb = blockIdx.y * rows_per, re = min(a.M, rb + rows_per);
    const float4 shift = live ? gld4(x + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    if (live) {
        int r = rb + rlane;
        for (; r + 3 * RL < re; r += 4 * RL) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = gld4(x + (size_t)(r + u * RL) * a.C + c);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float dx = v[u].x - shift.x, dy = v[u].y - shift.y, dz = v[u].z - shift.z, dw = v[u].w - shift.w;
                s.x += dx; s.y += dy; s.z += dz; s.w += dw;
                q.x = fmaf(dx, dx, q.x); q.y = fmaf(dy, dy, q.y); q.z = fmaf(dz, dz, q.z); q.w = fmaf(dw, dw, q.w);
            }
        }
        for (; r < re; r += RL) {
            const float4 v = gld4(x + (size_t)r * a.C + c);
            const float dx = v.x - shift.x, dy = v.y - shift.y, dz = v.z - shift.z, dw = v.w - shift.w;
            s.x += dx; s.y += dy; s.z += dz; s.w += dw;
            q.x = fmaf(dx, dx, q.x); q.y = fmaf(dy, dy, q.y); q.z = fmaf(dz, dz, q.z); q.w = fmaf(dw, dw, q.w);
        }
    }
    gbn_cta_reduce_pair<TC>(s, q, wpart, part);
    gbn_cluster_fold(cluster, part, tot);
    if (threadIdx.x < TC) {
        const int ch = blockIdx.x * TC + threadIdx.x;
        if (ch < a.C) {
            const float inv_m = 1.f / (float)a.M;
            const float ds = tot[threadIdx.x] * inv_m;
            const float var = fmaxf(tot[16 + threadIdx.x] * inv_m - ds * ds, 0.f);
            const float mean = x[ch] + ds;
            const float invstd = rsqrtf(var + a.eps);
            stat[threadIdx.x] = mean; stat[16 + threadIdx.x] = invstd;
            if (blockIdx.y == 0) {
                a.save_mean[(long long)g * a.C + ch] = mean; a.save_invstd[(long long)g * a.C + ch] = invstd;
                if (a.rmean_off >= 0) {
                    const float unbiased = var * ((float)a.M / fmaxf((float)a.M - 1.f, 1.f));
                    prow[a.rmean_off + ch] = (1.f - a.momentum) * prow[a.rmean_off + ch] + a.momentum * mean;
                    prow[a.rvar_off + ch] = (1.f - a.momentum) * prow[a.rvar_off + ch] + a.momentum * unbiased;
                }
                if (ch == 0 && a.nbt_off >= 0) a.nbt[(long long)slot * a.nbt_gs + a.nbt_off] += 1;
            }
        }
    }
    __syncthreads();
    if (!live) return;
    const int q4 = quad * 4;
    const float4 mean = make_float4(stat[q4], stat[q4 + 1], stat[q4 + 2], stat[q4 + 3]);
    const float4 gm = gld4(prow + a.gamma_off + c), bt = gld4(prow + a.beta_off + c);
    const float4 sc = make_float4(stat[16 + q4] * gm.x, stat[17 + q4] * gm.y, stat[18 + q4] * gm.z, stat[19 + q4] * gm.w);
    const float4 sh = make_float4(bt.x - mean.x * sc.x, bt.y - mean.y * sc.y, bt.z - mean.z * sc.z, bt.w - mean.w * sc.w);
    const bool drop = a.p_drop > 0.f;
    const uint64_t stream = drop ? ((uint64_t)(*a.rng_step) << 24) ^ ((uint64_t)a.layer_id << 12) ^ (uint64_t)slot : 0;
    for (int r = rb + rlane; r < re; r += RL) {
        const size_t o = (size_t)r * a.C + c;
        float4 v = gld4(x + o);
        v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
        if (res) { const float4 t = gld4(res + o); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
        if (a.relu) { v.x = relu_nan(v.x); v.y = relu_nan(v.y); v.z = relu_nan(v.z); v.w = relu_nan(v.w); }
        if (drop) { const float4 k = dropout_scale4(a.seed, stream, o >> 2, a.p_drop); v.x *= k.x; v.y *= k.y; v.z *= k.z; v.w *= k.w; }
        gst4(y + o, v);
    }
}

// dy → (dropout) → (ReLU mask from y) → dz;  dres = dz;  dx = γ·istd·(dz − mean(dz) − x̂·mean(dz·x̂));  γ −= lr·Σdz·x̂, β −= lr·Σdz
template <int TC>
__global__ void __launch_bounds__(kGbnThreads) gbn_bwd_kernel(GbnBwdArgs a) {
    pdl_launch_dependents(); pdl_wait();        // PDL: see pdl.cuh (no prologue worth overlapping here, only the launch latency)
    constexpr int Q = TC / 4, RL = kGbnThreads / Q;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ float wpart[kGbnThreads / 32][4][8];
    __shared__ float part[32], tot[32];
    const int g = blockIdx.z;
    const int slot = a.gmap ? a.gmap[g] : g;
    const float* dy = a.dy + (long long)g * a.dy_gs;
    const float* x = a.x + (long long)g * a.x_gs;
    const float* y = a.y + (long long)g * a.y_gs;
    float* dx = a.dx + (long long)g * a.dx_gs;
    float* dres = a.dres ? a.dres + (long long)g * a.dres_gs : nullptr;
    float* prow = a.arena + (long long)slot * a.arena_gs;
    const int quad = threadIdx.x & (Q - 1), rlane = threadIdx.x / Q;
    const int c = blockIdx.x * TC + quad * 4;
    const bool live = c < a.C;
    const int splits = gridDim.y;
    const int rows_per = (a.M + splits - 1) / splits;
    const int rb = blockIdx.y * rows_per, re = min(a.M, rb + rows_per);
    float4 mean = make_float4(0.f, 0.f, 0.f, 0.f), istd = mean, g4 = mean;
    if (live) {
        mean = gld4(a.save_mean + (long long)g * a.C + c); istd = gld4(a.save_invstd + (long long)g * a.C + c);
        g4 = gld4(prow + a.gamma_off + c);                    // γ is read before the fold; rank 0 updates it after
    }
    const bool drop = a.p_drop > 0.f;
    const uint64_t stream = drop ? ((uint64_t)(*a.rng_step) << 24) ^ ((uint64_t)a.layer_id << 12) ^ (uint64_t)slot : 0;
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
    if (live) {
        for (int r = rb + rlane; r < re; r += RL) {
            const size_t o = (size_t)r * a.C + c;
            float4 gq = gld4(dy + o);
            const float4 xv = gld4(x + o);
            if (drop) { const float4 k = dropout_scale4(a.seed, stream, o >> 2, a.p_drop); gq.x *= k.x; gq.y *= k.y; gq.z *= k.z; gq.w *= k.w; }
            if (a.relu) {
                const float4 yv = gld4(y + o);            // with dropout y = relu(·)·k: y > 0 ⇔ relu(·) > 0 and kept; dropped units have k = 0 already
                if (!drop) { gq.x = yv.x > 0.f ? gq.x : 0.f; gq.y = yv.y > 0.f ? gq.y : 0.f; gq.z = yv.z > 0.f ? gq.z : 0.f; gq.w = yv.w > 0.f ? gq.w : 0.f; }
                else {
                    // recompute the pre-dropout sign from the normalised input (dropped units carry no information in y)
                    const float4 bt = gld4(prow + a.beta_off + c);
                    gq.x = fmaf((xv.x - mean.x) * istd.x, g4.x, bt.x) > 0.f ? gq.x : 0.f; gq.y = fmaf((xv.y - mean.y) * istd.y, g4.y, bt.y) > 0.f ? gq.y : 0.f;
                    gq.z = fmaf((xv.z - mean.z) * istd.z, g4.z, bt.z) > 0.f ? gq.z : 0.f; gq.w = fmaf((xv.w - mean.w) * istd.w, g4.w, bt.w) > 0.f ? gq.w : 0.f;
                }
            }
            s1.x += gq.x; s1.y += gq.y; s1.z += gq.z; s1.w += gq.w;
            s2.x = fmaf(gq.x, (xv.x - mean.x) * istd.x, s2.x); s2.y = fmaf(gq.y, (xv.y - mean.y) * istd.y, s2.y);
            s2.z = fmaf(gq.z, (xv.z - mean.z) * istd.z, s2.z); s2.w = fmaf(gq.w, (xv.w - mean.w) * istd.w, s2.w);
        }
    }
    float4 bt4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live && drop && a.relu) bt4 = gld4(prow + a.beta_off + c);
    gbn_cta_reduce_pair<TC>(s1, s2, wpart, part);
    gbn_cluster_fold(cluster, part, tot);
    if (blockIdx.y == 0 && threadIdx.x < TC) {              // SGD step of the affine parameters (reference: plain SGD, core/node.py:74)
        const int ch = blockIdx.x * TC + threadIdx.x;
        if (ch < a.C) {
            prow[a.beta_off + ch] -= a.lr * tot[threadIdx.x];
            prow[a.gamma_off + ch] -= a.lr * tot[16 + threadIdx.x];
        }
    }
    if (!live) return;
    const int q4 = quad * 4;
    const float inv_m = 1.f / (float)a.M;
    const float4 k = make_float4(g4.x * istd.x, g4.y * istd.y, g4.z * istd.z, g4.w * istd.w);
    const float4 mb_ = make_float4(tot[q4] * inv_m, tot[q4 + 1] * inv_m, tot[q4 + 2] * inv_m, tot[q4 + 3] * inv_m);
    const float4 mg = make_float4(tot[16 + q4] * inv_m, tot[17 + q4] * inv_m, tot[18 + q4] * inv_m, tot[19 + q4] * inv_m);
    for (int r = rb + rlane; r < re; r += RL) {
        const size_t o = (size_t)r * a.C + c;
        float4 gq = gld4(dy + o);
        const float4 xv = gld4(x + o);
        if (drop) { const float4 kk = dropout_scale4(a.seed, stream, o >> 2, a.p_drop); gq.x *= kk.x; gq.y *= kk.y; gq.z *= kk.z; gq.w *= kk.w; }
        if (a.relu) {
            if (!drop) {
                const float4 yv = gld4(y + o);
                gq.x = yv.x > 0.f ? gq.x : 0.f; gq.y = yv.y > 0.f ? gq.y : 0.f; gq.z = yv.z > 0.f ? gq.z : 0.f; gq.w = yv.w > 0.f ? gq.w : 0.f;
            } else {
                gq.x = fmaf((xv.x - mean.x) * istd.x, g4.x, bt4.x) > 0.f ? gq.x : 0.f; gq.y = fmaf((xv.y - mean.y) * istd.y, g4.y, bt4.y) > 0.f ? gq.y : 0.f;
                gq.z = fmaf((xv.z - mean.z) * istd.z, g4.z, bt4.z) > 0.f ? gq.z : 0.f; gq.w = fmaf((xv.w - mean.w) * istd.w, g4.w, bt4.w) > 0.f ? gq.w : 0.f;
            }
        }
        if (dres) gst4(dres + o, gq);
        float4 d;
        d.x = k.x * (gq.x - mb_.x - (xv.x - mean.x) * istd.x * mg.x); d.y = k.y * (gq.y - mb_.y - (xv.y - mean.y) * istd.y * mg.y);
        d.z = k.z * (gq.z - mb_.z - (xv.z - mean.z) * istd.z * mg.z); d.w = k.w * (gq.w - mb_.w - (xv.w - mean.w) * istd.w * mg.w);
        gst4(dx + o, d);
    }
}

// =====================================================================================================================
// max pooling (NHWC in; NHWC or NCHW-flattened out), argmax kept as the tap index; backward gathers (no atomics)
// =====================================================================================================================
struct PoolArgs {
    const float* x; long long x_gs; float* y; long long y_gs; unsigned char* idx; long long idx_gs;
    int B, H, W, C, OH, OW, k, stride, pad, nchw_out;
};

__device__ __forceinline__ long long pool_out_index(const PoolArgs& a, int b, int oh, int ow, int c) {
    return a.nchw_out ? (((long long)b * a.C + c) * a.OH + oh) * a.OW + ow : (((long long)b * a.OH + oh) * a.OW + ow) * a.C + c;
}

// one thread = 4 consecutive channels of one output pixel (float4 loads / stores; no per-element div/mod chains)
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(PoolArgs a) {
    pdl_launch_dependents(); pdl_wait();        // PDL: see pdl.cuh (no prologue worth overlapping here, only the launch latency)
    const int g = blockIdx.y;
    const int C4 = a.C >> 2;
    const long long total = (long long)a.B * a.OH * a.OW * C4;
    const float* x = a.x + (long long)g * a.x_gs;
    float* y = a.y + (long long)g * a.y_gs;
    unsigned char* idx = a.idx + (long long)g * a.idx_gs;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) << 2; long long t = i / C4;
        const int ow = (int)(t % a.OW); t /= a.OW;
        const int oh = (int)(t % a.OH); const int b = (int)(t / a.OH);
        float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        int4 arg = make_int4(255, 255, 255, 255);
        for (int kh = 0; kh < a.k; ++kh) {
            const int ih = oh * a.stride + kh - a.pad;
            if (ih < 0 || ih >= a.H) continue;
            for (int kw = 0; kw < a.k; ++kw) {
                const int iw = ow * a.stride + kw - a.pad;
                if (iw < 0 || iw >= a.W) continue;
                const float4 v = *reinterpret_cast<const float4*>(x + (((long long)b * a.H + ih) * a.W + iw) * a.C + c);
                const int tap = kh * a.k + kw;                                   // first maximum in scan order (ATen's rule)
                if (v.x > best.x || v.x != v.x || arg.x == 255) { best.x = v.x; arg.x = tap; }
                if (v.y > best.y || v.y != v.y || arg.y == 255) { best.y = v.y; arg.y = tap; }
                if (v.z > best.z || v.z != v.z || arg.z == 255) { best.z = v.z; arg.z = tap; }
                if (v.w > best.w || v.w != v.w || arg.w == 255) { best.w = v.w; arg.w = tap; }
            }
        }
        const long long o = (((long long)b * a.OH + oh) * a.OW + ow) * a.C + c;     // NHWC position (also the index-map position)
        if (a.nchw_out) {
            const long long hw = (long long)a.OH * a.OW, base = ((long long)b * a.C + c) * hw + (long long)oh * a.OW + ow;
            y[base] = best.x; y[base + hw] = best.y; y[base + 2 * hw] = best.z; y[base + 3 * hw] = best.w;
        } else {
            *reinterpret_cast<float4*>(y + o) = best;
        }
        *reinterpret_cast<uchar4*>(idx + o) = make_uchar4((unsigned char)arg.x, (unsigned char)arg.y, (unsigned char)arg.z, (unsigned char)arg.w);
    }
}

// dx[b,ih,iw,c] = Σ_{windows containing (ih,iw) whose argmax is (ih,iw)} dy  (· [x > 0] when the pooled tensor is a ReLU output)
struct PoolBwdArgs {
    const float* dy; long long dy_gs; const unsigned char* idx; long long idx_gs; const float* x; long long x_gs;
    float* dx; long long dx_gs;
    int B, H, W, C, OH, OW, k, stride, pad, nchw_out, relu_mask;
};

__global__ void __launch_bounds__(256) maxpool_bwd_kernel(PoolBwdArgs a) {
    pdl_launch_dependents(); pdl_wait();        // PDL: see pdl.cuh (no prologue worth overlapping here, only the launch latency)
    const int g = blockIdx.y;
    const int C4 = a.C >> 2;
    const long long total = (long long)a.B * a.H * a.W * C4;
    const float* dy = a.dy + (long long)g * a.dy_gs;
    const unsigned char* idx = a.idx + (long long)g * a.idx_gs;
    const float* x = a.x + (long long)g * a.x_gs;
    float* dx = a.dx + (long long)g * a.dx_gs;
    const long long hw = (long long)a.OH * a.OW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) << 2; long long t = i / C4;
        const int iw = (int)(t % a.W); t /= a.W;
        const int ih = (int)(t % a.H); const int b = (int)(t / a.H);
        const long long xi = (((long long)b * a.H + ih) * a.W + iw) * a.C + c;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int kh = 0; kh < a.k; ++kh) {
            const int th = ih + a.pad - kh;
            if (th < 0 || th % a.stride != 0) continue;
            const int oh = th / a.stride;
            if (oh >= a.OH) continue;
            for (int kw = 0; kw < a.k; ++kw) {
                const int tw = iw + a.pad - kw;
                if (tw < 0 || tw % a.stride != 0) continue;
                const int ow = tw / a.stride;
                if (ow >= a.OW) continue;
                const long long o = (((long long)b * a.OH + oh) * a.OW + ow) * a.C + c;
                const uchar4 id = *reinterpret_cast<const uchar4*>(idx + o);
                const int tap = kh * a.k + kw;
                float4 d;
                if (a.nchw_out) {
                    const long long base = ((long long)b * a.C + c) * hw + (long long)oh * a.OW + ow;
                    d = make_float4(dy[base], dy[base + hw], dy[base + 2 * hw], dy[base + 3 * hw]);
                } else d = *reinterpret_cast<const float4*>(dy + o);
                if (id.x == tap) acc.x += d.x;
                if (id.y == tap) acc.y += d.y;
                if (id.z == tap) acc.z += d.z;
                if (id.w == tap) acc.w += d.w;
            }
        }
        if (a.relu_mask) {
            const float4 xv = *reinterpret_cast<const float4*>(x + xi);
            acc.x = xv.x > 0.f ? acc.x : 0.f; acc.y = xv.y > 0.f ? acc.y : 0.f; acc.z = xv.z > 0.f ? acc.z : 0.f; acc.w = xv.w > 0.f ? acc.w : 0.f;
        }
        *reinterpret_cast<float4*>(dx + xi) = acc;
    }
}

// ---- global average pooling [B][HW][C] ↔ [B][C] ------------------------------------------------------------------------
__global__ void __launch_bounds__(256) avgpool_fwd_kernel(const float* x, long long x_gs, float* y, long long y_gs, int B, int HW, int C) {
    pdl_launch_dependents(); pdl_wait();        // PDL: see pdl.cuh (no prologue worth overlapping here, only the launch latency)
    const int g = blockIdx.y;
    const float inv = 1.f / (float)HW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)B * C; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C); const long long b = i / C;
        const float* p = x + (long long)g * x_gs + b * HW * C + c;
        float s = 0.f;
        for (int h = 0; h < HW; ++h) s += p[(long long)h * C];
        y[(long long)g * y_gs + i] = s * inv;
    }
}
__global__ void __launch_bounds__(256) avgpool_bwd_kernel(const float* dy, long long dy_gs, float* dx, long long dx_gs, int B, int HW, int C) {
    pdl_launch_dependents(); pdl_wait();        // PDL: see pdl.cuh (no prologue worth overlapping here, only the launch latency)
    const int g = blockIdx.y;
    const float inv = 1.f / (float)HW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)B * HW * C; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C); const long long b = i / ((long long)HW * C);
        dx[(long long)g * dx_gs + i] = dy[(long long)g * dy_gs + b * C + c] * inv;
    }
}

// ---- stand-alone dropout (models that apply it after a plain ReLU) ------------------------------------------------------
__global__ void __launch_bounds__(256) dropout_kernel(const float* x, float* y, const float* mask, long long x_gs, long long y_gs, long long m_gs, long long n4, const int* gmap,
                                                      const long long* rng_step, unsigned long long seed, int layer_id, float p_drop) {
    pdl_launch_dependents(); pdl_wait();        // PDL: see pdl.cuh (no prologue worth overlapping here, only the launch latency)
    const int g = blockIdx.y;
    const int slot = gmap ? gmap[g] : g;
    const uint64_t stream = ((uint64_t)(*rng_step) << 24) ^ ((uint64_t)layer_id << 12) ^ (uint64_t)slot;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = reinterpret_cast<const float4*>(x + (long long)g * x_gs)[i];
        const float4 k = dropout_scale4(seed, stream, (uint64_t)i, p_drop);
        v.x *= k.x; v.y *= k.y; v.z *= k.z; v.w *= k.w;
        if (mask) {                                            // backward through a ReLU fused into the producer of the dropped tensor
            const float4 m = reinterpret_cast<const float4*>(mask + (long long)g * m_gs)[i];
            v.x = m.x > 0.f ? v.x : 0.f; v.y = m.y > 0.f ? v.y : 0.f; v.z = m.z > 0.f ? v.z : 0.f; v.w = m.w > 0.f ? v.w : 0.f;
        }
        reinterpret_cast<float4*>(y + (long long)g * y_gs)[i] = v;
    }
}

// =====================================================================================================================
// losses (one CTA per node): forward value accumulated into loss_acc[slot], gradient w.r.t. the layer's pre-activation
// =====================================================================================================================
__device__ __forceinline__ float gl_digamma(float x) {
    float r = 0.f;
    while (x < 6.f) { r -= 1.f / x; x += 1.f; }
    const float f = 1.f / (x * x);
    return r + logf(x) - 0.5f / x - f * (1.f / 12.f - f * (1.f / 120.f - f * (1.f / 252.f)));
}
__device__ __forceinline__ float gl_trigamma(float x) {
    float r = 0.f;
    while (x < 6.f) { r += 1.f / (x * x); x += 1.f; }
    const float f = 1.f / (x * x);
    return r + 1.f / x + 0.5f * f + (1.f / x) * f * (1.f / 6.f - f * (1.f / 30.f - f * (1.f / 42.f)));
}

// softmax cross-entropy, mean over the batch: grad[b][c] = (softmax − onehot)/B, padding columns [C, ld) are written as 0
__global__ void __launch_bounds__(256) ce_loss_grouped_kernel(const float* logits, long long gs, const long long* targets, long long t_gs,
                                                              float* grad, long long grad_gs, float* loss_acc, const int* gmap, int B, int C, int ld) {
    pdl_launch_dependents(); pdl_wait();        // PDL: see pdl.cuh (no prologue worth overlapping here, only the launch latency)
    __shared__ float wsum[8];
    const int g = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float invB = 1.f / (float)B;
    float local = 0.f;
    for (int r = warp; r < B; r += 8) {
        const float* z = logits + (long long)g * gs + (long long)r * ld;
        float* gr = grad + (long long)g * grad_gs + (long long)r * ld;
        const int t = (int)targets[(long long)g * t_gs + r];
        float mx = -INFINITY;
        for (int c = lane; c < C; c += 32) mx = fmaxf(mx, z[c]);
        mx = warp_max(mx);
        float se = 0.f;
        for (int c = lane; c < C; c += 32) se += __expf(z[c] - mx);
        se = warp_sum(se);
        const float lse = mx + __logf(se), inv = 1.f / se;
        for (int c = lane; c < ld; c += 32) gr[c] = c < C ? (__expf(z[c] - mx) * inv - (c == t ? 1.f : 0.f)) * invB : 0.f;
        if (lane == 0) local += lse - z[t];
    }
    if (lane == 0) wsum[warp] = local;
    __syncthreads();
    if (threadIdx.x == 0 && loss_acc) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) tot += wsum[w];
        loss_acc[gmap ? gmap[g] : g] += tot * invB;
    }
}

// evidential loss on α = softplus(z) + 1 (reference examples/wearables/models.py:89-179): grad is dL/dz = dL/dα · σ(z),
// σ(z) = 1 − exp(−(α − 1)); λ is read from the device (annealing schedule without re-capturing the graph)
__global__ void __launch_bounds__(256) evidential_loss_grouped_kernel(const float* alpha, long long gs, const long long* targets, long long t_gs,
                                                                      float* grad, long long grad_gs, float* loss_acc, const int* gmap, const float* lam_ptr,
                                                                      int B, int C, int ld) {
    pdl_launch_dependents(); pdl_wait();        // PDL: see pdl.cuh (no prologue worth overlapping here, only the launch latency)
    __shared__ float wsum[8];
    const int g = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float lam = *lam_ptr, invB = 1.f / (float)B;
    float local = 0.f;
    for (int r = warp; r < B; r += 8) {
        const float* a = alpha + (long long)g * gs + (long long)r * ld;
        float* gr = grad + (long long)g * grad_gs + (long long)r * ld;
        const int t = (int)targets[(long long)g * t_gs + r];
        float S = 0.f;
        for (int c = lane; c < C; c += 32) S += a[c];
        S = warp_sum(S);
        const float at = a[t];
        const float St = S - at + 1.f;
        float mse = 0.f, dot = 0.f, lg = 0.f, term = 0.f, sum_am1 = 0.f;
        const float psiSt = gl_digamma(St);
        for (int c = lane; c < C; c += 32) {
            const float p = a[c] / S, y = (c == t) ? 1.f : 0.f, d = p - y;
            mse = fmaf(d, d, mse); dot = fmaf(d, p, dot);
            const float at_c = (c == t) ? 1.f : a[c];
            lg += lgammaf(at_c);
            term += (at_c - 1.f) * (gl_digamma(at_c) - psiSt);
            sum_am1 += at_c - 1.f;
        }
        mse = warp_sum(mse); dot = warp_sum(dot); lg = warp_sum(lg); term = warp_sum(term); sum_am1 = warp_sum(sum_am1);
        const float kl = lgammaf(St) - lgammaf((float)C) - lg + term;
        const float tri_St = gl_trigamma(St);
        for (int c = lane; c < ld; c += 32) {
            float out = 0.f;
            if (c < C) {
                const float p = a[c] / S, y = (c == t) ? 1.f : 0.f;
                float d = (2.f / S) * ((p - y) - dot);
                if (c != t) d += lam * ((a[c] - 1.f) * gl_trigamma(a[c]) - tri_St * sum_am1);
                out = d * invB * (1.f - __expf(-(a[c] - 1.f)));
            }
            gr[c] = out;
        }
        if (lane == 0) local += mse + lam * kl;
    }
    if (lane == 0) wsum[warp] = local;
    __syncthreads();
    if (threadIdx.x == 0 && loss_acc) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) tot += wsum[w];
        loss_acc[gmap ? gmap[g] : g] += tot * invB;
    }
}

}  // namespace mb

// ---------------------------------------------------------------------------------------------------------------------
// host wrappers: plans are Python dicts with integer addresses (sub-buffers of the trainer workspace / arena)
// ---------------------------------------------------------------------------------------------------------------------
namespace {
template <typename T> T lget(const py::dict& d, const char* k, T def) { return d.contains(k) ? d[k].cast<T>() : def; }
template <typename T> T* lptr(const py::dict& d, const char* k) { return reinterpret_cast<T*>(d.contains(k) ? d[k].cast<int64_t>() : 0); }
cudaStream_t lstream() { return at::cuda::getCurrentCUDAStream().stream(); }
int gs_blocks(long long n, int G) { return (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, std::max(1, 148 * 8 / std::max(1, G)))); }

int gbn_row_splits(int M) { int s = 1; while (s < 8 && M > 256 * s) s <<= 1; return s; }
bool gbn_narrow(int C, int M, int G) { return ((C + 15) / 16) * gbn_row_splits(M) * G < 96; }

template <typename... P, typename... A>
void llaunch(void (*kernel)(P...), dim3 grid, size_t smem, A&&... args) {
    C10_CUDA_CHECK(mbhost::launch(kernel, grid, dim3(256), smem, lstream(), dim3(1, 1, 1), std::forward<A>(args)...));
}

template <typename Args>
void gbn_launch(void (*kernel)(Args), const Args& a, int C, int M, int tile, int G) {
    const int splits = gbn_row_splits(M);
    C10_CUDA_CHECK(mbhost::launch(kernel, dim3((C + tile - 1) / tile, splits, G), dim3(mb::kGbnThreads), 0, lstream(), dim3(1, splits, 1), a));
}
}  // namespace

void set_pdl(bool on) { mbhost::pdl_flag() = on; }      // A/B switch for programmatic dependent launch (default on)
bool get_pdl() { return mbhost::pdl_flag(); }

void gather_grouped(py::dict d) {
    mb::GatherArgs a;
    a.x_tab = lptr<const long long>(d, "x_tab"); a.y_tab = lptr<const long long>(d, "y_tab");
    a.perm = lptr<const long long>(d, "perm"); a.perm_ld = d["perm_ld"].cast<int64_t>();
    a.gmap = lptr<const int>(d, "gmap");
    a.xb = lptr<float>(d, "xb"); a.xb_gs = d["xb_gs"].cast<int64_t>();
    a.yb = lptr<long long>(d, "yb"); a.yb_gs = d["yb_gs"].cast<int64_t>();
    a.rng_step = lptr<long long>(d, "rng_step"); a.ticket = lptr<unsigned int>(d, "ticket");
    a.t = d["t"].cast<int>(); a.eb = d["eb"].cast<int>(); a.npix = d["npix"].cast<int>();
    a.Csrc = d["Csrc"].cast<int>(); a.Cdst = d["Cdst"].cast<int>();
    const int G = d["G"].cast<int>();
    TORCH_CHECK(G >= 1 && a.eb >= 1 && a.Cdst >= a.Csrc && (a.rng_step == nullptr || a.ticket != nullptr));
    const long long row = (long long)a.npix * a.Cdst;
    dim3 grid((unsigned)std::max<long long>(1, std::min<long long>(8, (row / 4 + 255) / 256)), (unsigned)a.eb, (unsigned)G);
    llaunch(mb::gather_grouped_kernel, grid, 0, a);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void im2col_pack(py::dict d) {
    mb::Im2colArgs a;
    a.x_tab = lptr<const long long>(d, "x_tab"); a.y_tab = lptr<const long long>(d, "y_tab");
    a.perm = lptr<const long long>(d, "perm"); a.perm_ld = d["perm_ld"].cast<int64_t>(); a.gmap = lptr<const int>(d, "gmap");
    a.xcol = lptr<float>(d, "xcol"); a.xcol_gs = d["xcol_gs"].cast<int64_t>();
    a.yb = lptr<long long>(d, "yb"); a.yb_gs = d["yb_gs"].cast<int64_t>();
    a.rng_step = lptr<long long>(d, "rng_step"); a.ticket = lptr<unsigned int>(d, "ticket");
    a.t = d["t"].cast<int>(); a.eb = d["eb"].cast<int>();
    a.IH = d["IH"].cast<int>(); a.IW = d["IW"].cast<int>(); a.Cin = d["Cin"].cast<int>(); a.KH = d["KH"].cast<int>(); a.KW = d["KW"].cast<int>();
    a.stride = d["stride"].cast<int>(); a.pad = d["pad"].cast<int>(); a.OH = d["OH"].cast<int>(); a.OW = d["OW"].cast<int>();
    a.Kreal = d["Kreal"].cast<int>(); a.Kpad = d["Kpad"].cast<int>();
    a.arena = lptr<const float>(d, "arena"); a.arena_gs = lget<int64_t>(d, "arena_gs", 0); a.row_tab = lptr<const long long>(d, "row_tab");
    a.wmap = lptr<const int>(d, "wmap"); a.w_off = lget<int64_t>(d, "w_off", 0); a.bias_off = lget<int64_t>(d, "bias_off", -1);
    a.bn_off[0] = lget<int64_t>(d, "bn_mean_off", -1); a.bn_off[1] = lget<int64_t>(d, "bn_var_off", -1);
    a.bn_off[2] = lget<int64_t>(d, "bn_gamma_off", -1); a.bn_off[3] = lget<int64_t>(d, "bn_beta_off", -1);
    a.wpack = lptr<float>(d, "wpack"); a.wpack_gs = lget<int64_t>(d, "wpack_gs", 0); a.Cout = lget<int>(d, "Cout", 0);
    const int G = d["G"].cast<int>();
    TORCH_CHECK(G >= 1 && a.eb >= 1 && a.Kpad % 4 == 0 && a.Kpad >= a.Kreal && (a.rng_step == nullptr || a.ticket != nullptr));
    TORCH_CHECK(a.Cin < 65536 && a.KW < 256 && a.KH < 128 && a.Kpad * 4 <= 48 * 1024, "im2col_pack: first-layer geometry out of range");
    const long long pix = (long long)a.OH * a.OW;                          // 8 warps per block, one output pixel per warp pass
    dim3 grid((unsigned)std::max<long long>(1, std::min<long long>(16, (pix + 7) / 8)), (unsigned)(a.eb + (a.wpack ? 1 : 0)), (unsigned)G);
    llaunch(mb::im2col_pack_kernel, grid, a.Kpad * sizeof(int), a);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void bn_fwd_grouped(py::dict d) {
    mb::GbnFwdArgs a;
    a.x = lptr<const float>(d, "x"); a.res = lptr<const float>(d, "res"); a.y = lptr<float>(d, "y");
    a.x_gs = d["x_gs"].cast<int64_t>(); a.y_gs = d["y_gs"].cast<int64_t>(); a.res_gs = lget<int64_t>(d, "res_gs", 0);
    a.save_mean = lptr<float>(d, "save_mean"); a.save_invstd = lptr<float>(d, "save_invstd");
    a.arena = lptr<float>(d, "arena"); a.arena_gs = d["arena_gs"].cast<int64_t>(); a.gmap = lptr<const int>(d, "gmap");
    a.gamma_off = d["gamma_off"].cast<int64_t>(); a.beta_off = d["beta_off"].cast<int64_t>();
    a.rmean_off = lget<int64_t>(d, "rmean_off", -1); a.rvar_off = lget<int64_t>(d, "rvar_off", -1);
    a.nbt = lptr<long long>(d, "nbt"); a.nbt_gs = lget<int64_t>(d, "nbt_gs", 0); a.nbt_off = lget<int64_t>(d, "nbt_off", -1);
    a.rng_step = lptr<const long long>(d, "rng_step"); a.seed = (unsigned long long)lget<int64_t>(d, "seed", 0);
    a.layer_id = lget<int>(d, "layer_id", 0); a.p_drop = lget<float>(d, "p_drop", 0.f);
    a.M = d["M"].cast<int>(); a.C = d["C"].cast<int>(); a.eps = lget<float>(d, "eps", 1e-5f); a.momentum = lget<float>(d, "momentum", 0.1f);
    a.relu = lget<int>(d, "relu", 1);
    const int G = d["G"].cast<int>();
    TORCH_CHECK((a.C & 3) == 0 && a.M >= 2 && G >= 1, "bn_fwd_grouped: C % 4 == 0 and M >= 2 required");
    TORCH_CHECK(a.p_drop <= 0.f || a.rng_step != nullptr, "bn_fwd_grouped: dropout needs the device step counter");
    if (a.nbt == nullptr) a.nbt_off = -1;
    if (gbn_narrow(a.C, a.M, G)) gbn_launch(mb::gbn_fwd_kernel<8>, a, a.C, a.M, 8, G);
    else gbn_launch(mb::gbn_fwd_kernel<16>, a, a.C, a.M, 16, G);
}

void bn_bwd_grouped(py::dict d) {
    mb::GbnBwdArgs a;
    a.dy = lptr<const float>(d, "dy"); a.x = lptr<const float>(d, "x"); a.y = lptr<const float>(d, "y");
    a.dy_gs = d["dy_gs"].cast<int64_t>(); a.x_gs = d["x_gs"].cast<int64_t>(); a.y_gs = d["y_gs"].cast<int64_t>();
    a.dx = lptr<float>(d, "dx"); a.dres = lptr<float>(d, "dres"); a.dx_gs = d["dx_gs"].cast<int64_t>(); a.dres_gs = lget<int64_t>(d, "dres_gs", 0);
    a.save_mean = lptr<const float>(d, "save_mean"); a.save_invstd = lptr<const float>(d, "save_invstd");
    a.arena = lptr<float>(d, "arena"); a.arena_gs = d["arena_gs"].cast<int64_t>(); a.gmap = lptr<const int>(d, "gmap");
    a.gamma_off = d["gamma_off"].cast<int64_t>(); a.beta_off = d["beta_off"].cast<int64_t>();
    a.rng_step = lptr<const long long>(d, "rng_step"); a.seed = (unsigned long long)lget<int64_t>(d, "seed", 0);
    a.layer_id = lget<int>(d, "layer_id", 0); a.p_drop = lget<float>(d, "p_drop", 0.f);
    a.M = d["M"].cast<int>(); a.C = d["C"].cast<int>(); a.relu = lget<int>(d, "relu", 1); a.lr = d["lr"].cast<float>();
    const int G = d["G"].cast<int>();
    TORCH_CHECK((a.C & 3) == 0 && G >= 1);
    if (gbn_narrow(a.C, a.M, G)) gbn_launch(mb::gbn_bwd_kernel<8>, a, a.C, a.M, 8, G);
    else gbn_launch(mb::gbn_bwd_kernel<16>, a, a.C, a.M, 16, G);
}

void maxpool_fwd_grouped(py::dict d) {
    mb::PoolArgs a;
    a.x = lptr<const float>(d, "x"); a.x_gs = d["x_gs"].cast<int64_t>(); a.y = lptr<float>(d, "y"); a.y_gs = d["y_gs"].cast<int64_t>();
    a.idx = lptr<unsigned char>(d, "idx"); a.idx_gs = d["idx_gs"].cast<int64_t>();
    a.B = d["B"].cast<int>(); a.H = d["H"].cast<int>(); a.W = d["W"].cast<int>(); a.C = d["C"].cast<int>();
    a.OH = d["OH"].cast<int>(); a.OW = d["OW"].cast<int>(); a.k = d["k"].cast<int>(); a.stride = d["stride"].cast<int>(); a.pad = d["pad"].cast<int>();
    a.nchw_out = lget<int>(d, "nchw_out", 0);
    const int G = d["G"].cast<int>();
    TORCH_CHECK(a.k * a.k < 255 && a.C % 4 == 0, "maxpool: C must be a multiple of 4");
    dim3 grid(gs_blocks((long long)a.B * a.OH * a.OW * (a.C / 4), G), G);
    llaunch(mb::maxpool_fwd_kernel, grid, 0, a);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void maxpool_bwd_grouped(py::dict d) {
    mb::PoolBwdArgs a;
    a.dy = lptr<const float>(d, "dy"); a.dy_gs = d["dy_gs"].cast<int64_t>(); a.idx = lptr<const unsigned char>(d, "idx"); a.idx_gs = d["idx_gs"].cast<int64_t>();
    a.x = lptr<const float>(d, "x"); a.x_gs = d["x_gs"].cast<int64_t>(); a.dx = lptr<float>(d, "dx"); a.dx_gs = d["dx_gs"].cast<int64_t>();
    a.B = d["B"].cast<int>(); a.H = d["H"].cast<int>(); a.W = d["W"].cast<int>(); a.C = d["C"].cast<int>();
    a.OH = d["OH"].cast<int>(); a.OW = d["OW"].cast<int>(); a.k = d["k"].cast<int>(); a.stride = d["stride"].cast<int>(); a.pad = d["pad"].cast<int>();
    a.nchw_out = lget<int>(d, "nchw_out", 0); a.relu_mask = lget<int>(d, "relu_mask", 0);
    const int G = d["G"].cast<int>();
    TORCH_CHECK(a.C % 4 == 0);
    dim3 grid(gs_blocks((long long)a.B * a.H * a.W * (a.C / 4), G), G);
    llaunch(mb::maxpool_bwd_kernel, grid, 0, a);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void avgpool_grouped(py::dict d) {
    const int G = d["G"].cast<int>(), B = d["B"].cast<int>(), HW = d["HW"].cast<int>(), C = d["C"].cast<int>();
    const bool bwd = lget<int>(d, "backward", 0) != 0;
    if (!bwd) {
        dim3 grid(gs_blocks((long long)B * C, G), G);
        llaunch(mb::avgpool_fwd_kernel, grid, 0, lptr<const float>(d, "x"), d["x_gs"].cast<int64_t>(), lptr<float>(d, "y"), d["y_gs"].cast<int64_t>(), B, HW, C);
    } else {
        dim3 grid(gs_blocks((long long)B * HW * C, G), G);
        llaunch(mb::avgpool_bwd_kernel, grid, 0, lptr<const float>(d, "dy"), d["dy_gs"].cast<int64_t>(), lptr<float>(d, "dx"), d["dx_gs"].cast<int64_t>(), B, HW, C);
    }
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void dropout_grouped(py::dict d) {
    const int G = d["G"].cast<int>();
    const long long n = d["n"].cast<int64_t>();
    TORCH_CHECK(n % 4 == 0);
    dim3 grid(gs_blocks(n / 4, G), G);
    llaunch(mb::dropout_kernel, grid, 0, lptr<const float>(d, "x"), lptr<float>(d, "y"), lptr<const float>(d, "mask"), d["x_gs"].cast<int64_t>(), d["y_gs"].cast<int64_t>(), lget<int64_t>(d, "m_gs", 0), n / 4, lptr<const int>(d, "gmap"),
                                                    lptr<const long long>(d, "rng_step"), (unsigned long long)lget<int64_t>(d, "seed", 0),
                                                    lget<int>(d, "layer_id", 0), d["p_drop"].cast<float>());
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void loss_grouped(py::dict d) {
    const int G = d["G"].cast<int>(), B = d["B"].cast<int>(), C = d["C"].cast<int>(), ld = d["ld"].cast<int>();
    const bool evidential = lget<int>(d, "evidential", 0) != 0;
    TORCH_CHECK(B > 0 && C > 0 && ld >= C);
    if (evidential)
        llaunch(mb::evidential_loss_grouped_kernel, G, 0, lptr<const float>(d, "out"), d["gs"].cast<int64_t>(), lptr<const long long>(d, "targets"),
            d["t_gs"].cast<int64_t>(), lptr<float>(d, "grad"), lget<int64_t>(d, "grad_gs", d["gs"].cast<int64_t>()), lptr<float>(d, "loss_acc"), lptr<const int>(d, "gmap"), lptr<const float>(d, "lam"), B, C, ld);
    else
        llaunch(mb::ce_loss_grouped_kernel, G, 0, lptr<const float>(d, "out"), d["gs"].cast<int64_t>(), lptr<const long long>(d, "targets"),
            d["t_gs"].cast<int64_t>(), lptr<float>(d, "grad"), lget<int64_t>(d, "grad_gs", d["gs"].cast<int64_t>()), lptr<float>(d, "loss_acc"), lptr<const int>(d, "gmap"), B, C, ld);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}
