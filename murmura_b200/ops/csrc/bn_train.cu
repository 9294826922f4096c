// Fused training-mode BatchNorm (+ residual add) (+ ReLU), forward and backward, for NHWC / [B,C] fp32 activations.
//
// Replaces the cuDNN batch-norm kernel + the separate `add` / `relu` / `threshold_backward` / `num_batches_tracked += 1`
// element-wise launches of one local SGD step (reference: `core/node.py:59-109` runs stock `nn.BatchNorm2d` + `nn.ReLU`).
// The activations of the federated models are tiny (≤ 4 MB, L2 resident), so a step is launch-latency bound: what matters
// is the NUMBER of launches and the latency of each.  One launch here does the whole layer:
//
//   grid = (channel tiles of 16 or 8, row splits);  the row splits of one channel tile form a THREAD-BLOCK CLUSTER (≤ 8 CTAs):
//   pass 1  each CTA reduces its rows to per-channel partial sums (shifted by the first row for stability),
//   merge   partials are exchanged through distributed shared memory (`cluster.map_shared_rank`) — no global scratch,
//           no second kernel, no atomics; every CTA folds them in rank order, so the result is deterministic,
//   pass 2  normalise (+ residual) (+ ReLU) and store; rank 0 updates running stats / saved stats / num_batches_tracked.
//
// Backward is the same shape: pass 1 → Σdz, Σdz·x̂ ; merge ; pass 2 → dx (and the residual branch's gradient dz).
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAException.h>

namespace cg = cooperative_groups;
using torch::Tensor;

#include "common.cuh"

namespace mb {

constexpr int kBnThreads = 256;      // (TC/4) channel quads × (1024/TC) row lanes; TC = channels per CTA:
                                     // 16 (64 B of every row) normally, 8 (one 32 B sector) when that is needed to fill the SMs

struct BnFwdArgs {
    const float* x; const float* res; float* y;
    const float* gamma; const float* beta;
    float* rmean; float* rvar; long long* nbt;
    float* save_mean; float* save_invstd;
    int M, C; float eps, momentum; int relu;
};

struct BnBwdArgs {
    const float* dy; const float* x; const float* y; const float* gamma;
    const float* save_mean; const float* save_invstd;
    float* dx; float* dres; float* dgamma; float* dbeta;
    int M, C; int relu;
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// Reduce two float4 accumulators over the row lanes of the CTA → part[0..TC-1] (first) and part[16..16+TC-1] (second).
template <int TC>
__device__ __forceinline__ void cta_reduce_pair(float4 a, float4 b, float (*wpart)[4][8], float* part) {
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
    if (threadIdx.x < 32) {                                  // thread t: accumulator (t / 16), channel (t % 16)
        const int which = threadIdx.x >> 4, ch = threadIdx.x & 15;
        float s = 0.f;
        if (ch < TC) {
#pragma unroll
            for (int w = 0; w < kBnThreads / 32; ++w) s += wpart[w][ch >> 2][which * 4 + (ch & 3)];
        }
        part[threadIdx.x] = s;
    }
}

// Sum part[] over the CTAs of the cluster (rank order) → tot[0..31] in every CTA.
__device__ __forceinline__ void cluster_fold(cg::cluster_group& cluster, float* part, float* tot) {
    cluster.sync();                                          // partials of every CTA are visible
    if (threadIdx.x < 32) {
        float s = 0.f;
        const unsigned n = cluster.num_blocks();
        for (unsigned r = 0; r < n; ++r) s += cluster.map_shared_rank(part, r)[threadIdx.x];
        tot[threadIdx.x] = s;
    }
    cluster.sync();                                          // nobody leaves (or reuses part[]) while peers still read it
}

template <int TC>
__global__ void __launch_bounds__(kBnThreads) bn_act_fwd_kernel(BnFwdArgs a) {
    constexpr int kBnTile = TC, Q = TC / 4, kBnRowLanes = kBnThreads / Q;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ float wpart[kBnThreads / 32][4][8];
    __shared__ float part[32], tot[32], stat[32];
    const int quad = threadIdx.x & (Q - 1), rlane = threadIdx.x / Q;
    const int c = blockIdx.x * kBnTile + quad * 4;
    const bool live = c < a.C;
    const int splits = gridDim.y;
    const int rows_per = (a.M + splits - 1) / splits;
    const int rb = blockIdx.y * rows_per, re = min(a.M, rb + rows_per);
    const float4 shift = live ? ld4(a.x + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    if (live) {
        int r = rb + rlane;
        for (; r + 3 * kBnRowLanes < re; r += 4 * kBnRowLanes) {                     // 4 independent 128-bit loads in flight
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = ld4(a.x + (size_t)(r + u * kBnRowLanes) * a.C + c);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float dx = v[u].x - shift.x, dy = v[u].y - shift.y, dz = v[u].z - shift.z, dw = v[u].w - shift.w;
                s.x += dx; s.y += dy; s.z += dz; s.w += dw;
                q.x = fmaf(dx, dx, q.x); q.y = fmaf(dy, dy, q.y); q.z = fmaf(dz, dz, q.z); q.w = fmaf(dw, dw, q.w);
            }
        }
        for (; r < re; r += kBnRowLanes) {
            const float4 v = ld4(a.x + (size_t)r * a.C + c);
            const float dx = v.x - shift.x, dy = v.y - shift.y, dz = v.z - shift.z, dw = v.w - shift.w;
            s.x += dx; s.y += dy; s.z += dz; s.w += dw;
            q.x = fmaf(dx, dx, q.x); q.y = fmaf(dy, dy, q.y); q.z = fmaf(dz, dz, q.z); q.w = fmaf(dw, dw, q.w);
        }
    }
    cta_reduce_pair<TC>(s, q, wpart, part);
    cluster_fold(cluster, part, tot);
    if (threadIdx.x < kBnTile) {
        const int ch = blockIdx.x * kBnTile + threadIdx.x;
        if (ch < a.C) {
            const float inv_m = 1.f / (float)a.M;
            const float ds = tot[threadIdx.x] * inv_m;
            const float var = fmaxf(tot[16 + threadIdx.x] * inv_m - ds * ds, 0.f);
            const float mean = a.x[ch] + ds;                 // shift = first row of this channel
            const float invstd = rsqrtf(var + a.eps);
            stat[threadIdx.x] = mean; stat[16 + threadIdx.x] = invstd;
            if (blockIdx.y == 0) {
                a.save_mean[ch] = mean; a.save_invstd[ch] = invstd;
                if (a.rmean) {
                    const float unbiased = var * ((float)a.M / fmaxf((float)a.M - 1.f, 1.f));
                    a.rmean[ch] = (1.f - a.momentum) * a.rmean[ch] + a.momentum * mean;
                    a.rvar[ch] = (1.f - a.momentum) * a.rvar[ch] + a.momentum * unbiased;
                }
                if (ch == 0 && a.nbt) *a.nbt += 1;
            }
        }
    }
    __syncthreads();
    if (!live) return;
    const int q4 = quad * 4;
    const float4 mean = make_float4(stat[q4], stat[q4 + 1], stat[q4 + 2], stat[q4 + 3]);
    const float4 g = ld4(a.gamma + c), b = ld4(a.beta + c);
    float4 sc = make_float4(stat[16 + q4] * g.x, stat[17 + q4] * g.y, stat[18 + q4] * g.z, stat[19 + q4] * g.w);
    const float4 sh = make_float4(b.x - mean.x * sc.x, b.y - mean.y * sc.y, b.z - mean.z * sc.z, b.w - mean.w * sc.w);
    for (int r = rb + rlane; r < re; r += kBnRowLanes) {
        const size_t o = (size_t)r * a.C + c;
        float4 v = ld4(a.x + o);
        v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
        if (a.res) { const float4 t = ld4(a.res + o); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
        if (a.relu) { v.x = relu_nan(v.x); v.y = relu_nan(v.y); v.z = relu_nan(v.z); v.w = relu_nan(v.w); }
        st4(a.y + o, v);
    }
}

template <int TC>
__global__ void __launch_bounds__(kBnThreads) bn_act_bwd_kernel(BnBwdArgs a) {
    constexpr int kBnTile = TC, Q = TC / 4, kBnRowLanes = kBnThreads / Q;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ float wpart[kBnThreads / 32][4][8];
    __shared__ float part[32], tot[32];
    const int quad = threadIdx.x & (Q - 1), rlane = threadIdx.x / Q;
    const int c = blockIdx.x * kBnTile + quad * 4;
    const bool live = c < a.C;
    const int splits = gridDim.y;
    const int rows_per = (a.M + splits - 1) / splits;
    const int rb = blockIdx.y * rows_per, re = min(a.M, rb + rows_per);
    float4 mean = make_float4(0.f, 0.f, 0.f, 0.f), istd = mean;
    if (live) { mean = ld4(a.save_mean + c); istd = ld4(a.save_invstd + c); }
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
    if (live) {
        for (int r = rb + rlane; r < re; r += kBnRowLanes) {
            const size_t o = (size_t)r * a.C + c;
            float4 g = ld4(a.dy + o);
            const float4 xv = ld4(a.x + o);
            if (a.relu) {
                const float4 yv = ld4(a.y + o);
                g.x = yv.x > 0.f ? g.x : 0.f; g.y = yv.y > 0.f ? g.y : 0.f; g.z = yv.z > 0.f ? g.z : 0.f; g.w = yv.w > 0.f ? g.w : 0.f;
            }
            s1.x += g.x; s1.y += g.y; s1.z += g.z; s1.w += g.w;
            s2.x = fmaf(g.x, (xv.x - mean.x) * istd.x, s2.x); s2.y = fmaf(g.y, (xv.y - mean.y) * istd.y, s2.y);
            s2.z = fmaf(g.z, (xv.z - mean.z) * istd.z, s2.z); s2.w = fmaf(g.w, (xv.w - mean.w) * istd.w, s2.w);
        }
    }
    cta_reduce_pair<TC>(s1, s2, wpart, part);
    cluster_fold(cluster, part, tot);
    if (blockIdx.y == 0 && threadIdx.x < kBnTile) {
        const int ch = blockIdx.x * kBnTile + threadIdx.x;
        if (ch < a.C) { a.dbeta[ch] = tot[threadIdx.x]; a.dgamma[ch] = tot[16 + threadIdx.x]; }
    }
    if (!live) return;
    const int q4 = quad * 4;
    const float inv_m = 1.f / (float)a.M;
    const float4 g4 = ld4(a.gamma + c);
    const float4 k = make_float4(g4.x * istd.x, g4.y * istd.y, g4.z * istd.z, g4.w * istd.w);
    const float4 mb_ = make_float4(tot[q4] * inv_m, tot[q4 + 1] * inv_m, tot[q4 + 2] * inv_m, tot[q4 + 3] * inv_m);
    const float4 mg = make_float4(tot[16 + q4] * inv_m, tot[17 + q4] * inv_m, tot[18 + q4] * inv_m, tot[19 + q4] * inv_m);
    for (int r = rb + rlane; r < re; r += kBnRowLanes) {
        const size_t o = (size_t)r * a.C + c;
        float4 g = ld4(a.dy + o);
        const float4 xv = ld4(a.x + o);
        if (a.relu) {
            const float4 yv = ld4(a.y + o);
            g.x = yv.x > 0.f ? g.x : 0.f; g.y = yv.y > 0.f ? g.y : 0.f; g.z = yv.z > 0.f ? g.z : 0.f; g.w = yv.w > 0.f ? g.w : 0.f;
        }
        if (a.dres) st4(a.dres + o, g);
        float4 d;
        d.x = k.x * (g.x - mb_.x - (xv.x - mean.x) * istd.x * mg.x); d.y = k.y * (g.y - mb_.y - (xv.y - mean.y) * istd.y * mg.y);
        d.z = k.z * (g.z - mb_.z - (xv.z - mean.z) * istd.z * mg.z); d.w = k.w * (g.w - mb_.w - (xv.w - mean.w) * istd.w * mg.w);
        st4(a.dx + o, d);
    }
}

}  // namespace mb

namespace {

cudaStream_t bn_stream() { return at::cuda::getCurrentCUDAStream().stream(); }

// [M, C] view of an activation: 2-D contiguous, or 4-D channels_last (any 4-D tensor with H = W = 1 qualifies).
bool rows_channels(const Tensor& x, int& M, int& C) {
    if (x.dim() == 2 && x.is_contiguous()) { M = (int)x.size(0); C = (int)x.size(1); return true; }
    if (x.dim() == 4 && x.is_contiguous(at::MemoryFormat::ChannelsLast)) {
        M = (int)(x.size(0) * x.size(2) * x.size(3)); C = (int)x.size(1); return true;
    }
    return false;
}

int row_splits(int M) {                       // ≈256 rows per CTA, power of two, portable cluster size ≤ 8
    int s = 1;
    while (s < 8 && M > 256 * s) s <<= 1;
    return s;
}

// 8-channel tiles when 16-channel tiles would leave more than a third of the 148 SMs idle
bool narrow_tiles(int C, int M) { return ((C + 15) / 16) * row_splits(M) < 96; }

template <typename Args>
void launch_clustered(void (*kernel)(Args), const Args& a, int C, int M, int tile) {
    cudaLaunchConfig_t cfg = {};
    const int splits = row_splits(M);
    cfg.gridDim = dim3((C + tile - 1) / tile, splits, 1);
    cfg.blockDim = dim3(mb::kBnThreads, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = bn_stream();
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = splits; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    C10_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, a));
}

}  // namespace

// y, save_mean, save_invstd
std::vector<Tensor> bn_act_fwd(Tensor x, c10::optional<Tensor> residual, Tensor gamma, Tensor beta, c10::optional<Tensor> rmean,
                               c10::optional<Tensor> rvar, c10::optional<Tensor> nbt, double momentum, double eps, bool relu) {
    c10::cuda::CUDAGuard guard(x.device());
    int M = 0, C = 0;
    TORCH_CHECK(x.dtype() == torch::kFloat32 && rows_channels(x, M, C), "bn_act_fwd: need fp32 [B,C] or channels_last [B,C,H,W]");
    TORCH_CHECK((C & 3) == 0 && M >= 2, "bn_act_fwd: C must be a multiple of 4 and there must be >1 value per channel");
    TORCH_CHECK(gamma.numel() == C && beta.numel() == C && gamma.is_contiguous() && beta.is_contiguous());
    Tensor y = torch::empty_like(x);
    Tensor save_mean = torch::empty({C}, x.options()), save_invstd = torch::empty({C}, x.options());
    mb::BnFwdArgs a;
    a.x = x.data_ptr<float>(); a.y = y.data_ptr<float>(); a.res = nullptr;
    if (residual.has_value() && residual->defined()) {
        int Mr = 0, Cr = 0;
        TORCH_CHECK(residual->dtype() == torch::kFloat32 && rows_channels(*residual, Mr, Cr) && Mr == M && Cr == C && residual->dim() == x.dim(),
                    "bn_act_fwd: residual layout must match x");
        a.res = residual->data_ptr<float>();
    }
    a.gamma = gamma.data_ptr<float>(); a.beta = beta.data_ptr<float>();
    const bool track = rmean.has_value() && rmean->defined();
    a.rmean = track ? rmean->data_ptr<float>() : nullptr; a.rvar = track ? rvar->data_ptr<float>() : nullptr;
    a.nbt = (nbt.has_value() && nbt->defined()) ? reinterpret_cast<long long*>(nbt->data_ptr<int64_t>()) : nullptr;
    a.save_mean = save_mean.data_ptr<float>(); a.save_invstd = save_invstd.data_ptr<float>();
    a.M = M; a.C = C; a.eps = (float)eps; a.momentum = (float)momentum; a.relu = relu ? 1 : 0;
    if (narrow_tiles(C, M)) launch_clustered(mb::bn_act_fwd_kernel<8>, a, C, M, 8);
    else launch_clustered(mb::bn_act_fwd_kernel<16>, a, C, M, 16);
    return {y, save_mean, save_invstd};
}

// dx, dgamma, dbeta, dres (undefined tensor when !want_dres)
std::vector<Tensor> bn_act_bwd(Tensor dy, Tensor x, Tensor y, Tensor gamma, Tensor save_mean, Tensor save_invstd, bool relu, bool want_dres) {
    c10::cuda::CUDAGuard guard(x.device());
    int M = 0, C = 0, Md = 0, Cd = 0;
    TORCH_CHECK(rows_channels(x, M, C) && rows_channels(dy, Md, Cd) && Md == M && Cd == C && dy.dtype() == torch::kFloat32,
                "bn_act_bwd: dy must have the layout of x");
    Tensor dx = torch::empty_like(x), dgamma = torch::empty({C}, x.options()), dbeta = torch::empty({C}, x.options());
    Tensor dres = want_dres ? torch::empty_like(x) : Tensor();
    mb::BnBwdArgs a;
    a.dy = dy.data_ptr<float>(); a.x = x.data_ptr<float>(); a.y = y.data_ptr<float>(); a.gamma = gamma.data_ptr<float>();
    a.save_mean = save_mean.data_ptr<float>(); a.save_invstd = save_invstd.data_ptr<float>();
    a.dx = dx.data_ptr<float>(); a.dres = want_dres ? dres.data_ptr<float>() : nullptr;
    a.dgamma = dgamma.data_ptr<float>(); a.dbeta = dbeta.data_ptr<float>();
    a.M = M; a.C = C; a.relu = relu ? 1 : 0;
    if (narrow_tiles(C, M)) launch_clustered(mb::bn_act_bwd_kernel<8>, a, C, M, 8);
    else launch_clustered(mb::bn_act_bwd_kernel<16>, a, C, M, 16);
    return {dx, dgamma, dbeta, dres};
}
