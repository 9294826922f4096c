// murmura_b200 — local-training and evaluation kernels (sm_100a).
//
// Reference call sites replaced: K8 local SGD step (murmura/core/node.py:74-99 —
// fresh plain SGD, θ -= lr·g), K14 evaluation accumulators (murmura/core/node.py:134-196,
// murmura/utils/metrics.py:9-47, one host sync per batch in the reference → none here), the
// evidential loss (murmura/examples/wearables/models.py:89-179) fused forward+backward, and the
// Dirichlet epilogue used by EvidentialTrust / DMTT scoring (aggregation/evidential_trust.py:236-281).
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include "common.cuh"

namespace mb {

// ---- fused multi-slot SGD: θ -= lr·g ; g = 0 (one launch for a whole slot range) -------------------
__global__ void __launch_bounds__(256) sgd_step_kernel(float* __restrict__ live, size_t stride, float* __restrict__ grad,
                                                       size_t gstride, int n4, float lr) {
    const int v = blockIdx.y;
    float4* p = reinterpret_cast<float4*>(live + (size_t)v * stride);
    float4* g = reinterpret_cast<float4*>(grad + (size_t)v * gstride);
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        float4 w = p[i];
        const float4 d = g[i];
        w.x = fmaf(-lr, d.x, w.x); w.y = fmaf(-lr, d.y, w.y); w.z = fmaf(-lr, d.z, w.z); w.w = fmaf(-lr, d.w, w.w);
        p[i] = w;
        g[i] = zero;
    }
}

// ---- multi-tensor SGD: θ_t -= lr·g_t for up to 64 (param, grad) pairs in ONE launch ----------------------
// Gradients stay wherever autograd produced them (no accumulate-into-arena pass); pointers travel in the kernel
// parameter block, so the launch is CUDA-graph capturable with zero device-side tables.
constexpr int kMtTensors = 64;
constexpr int kMtBlocks = 1024;                   // (kernel parameter block: ~6.4 KiB, limit 32 KiB since CUDA 12.1)
constexpr int kMtChunk = 16384;                   // floats per block → ResNet-18 = 690 blocks ≈ 4.7 per SM
struct MultiSgdArgs {
    float* p[kMtTensors];
    const float* g[kMtTensors];
    int n[kMtTensors];
    unsigned char block_tensor[kMtBlocks];
    int block_chunk[kMtBlocks];
};
__global__ void __launch_bounds__(256) sgd_multi_kernel(const __grid_constant__ MultiSgdArgs a, float lr) {
    const int t = a.block_tensor[blockIdx.x];
    const int base = a.block_chunk[blockIdx.x] * kMtChunk;
    const int n = min(a.n[t] - base, kMtChunk);
    float* __restrict__ p = a.p[t] + base;
    const float* __restrict__ g = a.g[t] + base;
    if ((((uintptr_t)p | (uintptr_t)g) & 15) == 0) {
        const int n4 = n >> 2;
        for (int i = threadIdx.x; i < n4; i += blockDim.x) {
            float4 w = reinterpret_cast<float4*>(p)[i];
            const float4 d = reinterpret_cast<const float4*>(g)[i];
            w.x = fmaf(-lr, d.x, w.x); w.y = fmaf(-lr, d.y, w.y); w.z = fmaf(-lr, d.z, w.z); w.w = fmaf(-lr, d.w, w.w);
            reinterpret_cast<float4*>(p)[i] = w;
        }
        for (int i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) p[i] = fmaf(-lr, g[i], p[i]);
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = fmaf(-lr, g[i], p[i]);
    }
}

// ---- BatchNorm inference (+ optional ReLU) as one per-channel affine pass ---------------------------------------
// y = (x - mean[c]) * rsqrt(var[c] + eps) * gamma[c] + beta[c].  cuDNN's fp32 inference kernel degenerates on the tiny
// late-stage feature maps of ResNet-18@32x32 (37 us per call at [B,512,1,1]); this is a plain bandwidth-bound pass.
// `inner` = elements between consecutive channel indices: 1 for channels_last (NHWC) tensors, H*W for NCHW / [B,C].
__global__ void __launch_bounds__(256) bn_eval_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ mean,
                                                      const float* __restrict__ var, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, float eps, long long n, int C, int inner, int relu,
                                                      const float* __restrict__ res) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    if (inner == 1 && (C & 3) == 0) {                        // NHWC, 4 consecutive channels per thread
        const long long n4 = n >> 2;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            const int c = (int)((i << 2) % C);
            float4 v = reinterpret_cast<const float4*>(x)[i];
            const float4 m = *reinterpret_cast<const float4*>(mean + c), s2 = *reinterpret_cast<const float4*>(var + c);
            const float4 g = gamma ? *reinterpret_cast<const float4*>(gamma + c) : make_float4(1.f, 1.f, 1.f, 1.f);
            const float4 b = beta ? *reinterpret_cast<const float4*>(beta + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            v.x = (v.x - m.x) * rsqrtf(s2.x + eps) * g.x + b.x; v.y = (v.y - m.y) * rsqrtf(s2.y + eps) * g.y + b.y;
            v.z = (v.z - m.z) * rsqrtf(s2.z + eps) * g.z + b.z; v.w = (v.w - m.w) * rsqrtf(s2.w + eps) * g.w + b.w;
            if (res) { const float4 t = reinterpret_cast<const float4*>(res)[i]; v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
            if (relu) { v.x = relu_nan(v.x); v.y = relu_nan(v.y); v.z = relu_nan(v.z); v.w = relu_nan(v.w); }
            reinterpret_cast<float4*>(y)[i] = v;
        }
        return;
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int c = (int)((i / inner) % C);
        float v = (x[i] - mean[c]) * rsqrtf(var[c] + eps) * (gamma ? gamma[c] : 1.f) + (beta ? beta[c] : 0.f);
        if (res) v += res[i];
        y[i] = relu ? relu_nan(v) : v;
    }
}

// ---- softmax cross-entropy evaluation: one warp per row, device-side accumulators ----------------
// stats[0] += Σ CE, stats[1] += #correct, stats[2] += #rows  (rows >= n_valid are padding)
__global__ void ce_eval_kernel(const float* __restrict__ logits, const long long* __restrict__ targets,
                               const int* __restrict__ n_valid_ptr, int B, int C, float* __restrict__ stats) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int n_valid = n_valid_ptr ? min(*n_valid_ptr, B) : B;
    float loss = 0.f, correct = 0.f, rows = 0.f;
    if (warp < n_valid) {
        const float* z = logits + (size_t)warp * C;
        float mx = -INFINITY; int arg = 0;
        for (int c = lane; c < C; c += 32) { const float x = z[c]; if (x > mx) { mx = x; arg = c; } }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, mx, o); const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
        }
        float se = 0.f;
        for (int c = lane; c < C; c += 32) se += __expf(z[c] - mx);
        se = warp_sum(se);
        const int t = (int)targets[warp];
        loss = __logf(se) + mx - z[t];
        correct = (arg == t) ? 1.f : 0.f;
        rows = 1.f;
    }
    if (lane == 0 && rows != 0.f) { atomicAdd(stats + 0, loss); atomicAdd(stats + 1, correct); atomicAdd(stats + 2, rows); }
}

// ---- Dirichlet (evidential) evaluation: alpha rows → correct, vacuity, entropy, strength, sq-err ----
// stats[0] += Σ‖y - α/S‖², [1] += #correct, [2] += #rows, [3] += Σ K/S, [4] += Σ H(α/S), [5] += Σ S
__global__ void dirichlet_eval_kernel(const float* __restrict__ alpha, const long long* __restrict__ targets,
                                      const int* __restrict__ n_valid_ptr, int B, int C, float* __restrict__ stats) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int n_valid = n_valid_ptr ? min(*n_valid_ptr, B) : B;
    if (warp >= n_valid) return;
    const float* a = alpha + (size_t)warp * C;
    float S = 0.f, mx = -INFINITY; int arg = 0;
    for (int c = lane; c < C; c += 32) { const float x = a[c]; S += x; if (x > mx) { mx = x; arg = c; } }
    S = warp_sum(S);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, mx, o); const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
    }
    const int t = (int)targets[warp];
    float ent = 0.f, sq = 0.f;
    for (int c = lane; c < C; c += 32) {
        const float p = a[c] / S;
        ent -= p * logf(p + 1e-10f);
        const float d = ((c == t) ? 1.f : 0.f) - p;
        sq = fmaf(d, d, sq);
    }
    ent = warp_sum(ent); sq = warp_sum(sq);
    if (lane == 0) {
        atomicAdd(stats + 0, sq); atomicAdd(stats + 1, arg == t ? 1.f : 0.f); atomicAdd(stats + 2, 1.f);
        atomicAdd(stats + 3, (float)C / S); atomicAdd(stats + 4, ent); atomicAdd(stats + 5, S);
    }
}

// ---- evidential loss, forward + gradient in one pass ----------------------------------------------
__device__ __forceinline__ float digammaf_pos(float x) {            // x > 0
    float r = 0.f;
    while (x < 6.f) { r -= 1.f / x; x += 1.f; }
    const float f = 1.f / (x * x);
    return r + logf(x) - 0.5f / x - f * (1.f / 12.f - f * (1.f / 120.f - f * (1.f / 252.f)));
}
__device__ __forceinline__ float trigammaf_pos(float x) {           // x > 0
    float r = 0.f;
    while (x < 6.f) { r += 1.f / (x * x); x += 1.f; }
    const float f = 1.f / (x * x);
    return r + 1.f / x + 0.5f * f + (1.f / x) * f * (1.f / 6.f - f * (1.f / 30.f - f * (1.f / 42.f)));
}

// loss = mean_b [ Σ_k (y_k - α_k/S)² + λ·KL(Dir(α̃)‖Dir(1)) ],  α̃ = y + (1-y)·α
// One warp per row; writes dL/dα (already divided by B) and accumulates the mean loss.
__global__ void evidential_loss_kernel(const float* __restrict__ alpha, const long long* __restrict__ targets, int B, int C,
                                       float lam, const float* __restrict__ lam_ptr, float* __restrict__ loss_out,
                                       float* __restrict__ grad_out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    if (lam_ptr != nullptr) lam = *lam_ptr;                       // annealing coefficient kept on the device (CUDA-graph friendly)
    const float* a = alpha + (size_t)warp * C;
    float* g = grad_out + (size_t)warp * C;
    const int t = (int)targets[warp];
    float S = 0.f;
    for (int c = lane; c < C; c += 32) S += a[c];
    S = warp_sum(S);
    const float at = a[t];
    const float St = S - at + 1.f;                               // Σ α̃
    float mse = 0.f, dot = 0.f, lg = 0.f, term = 0.f, sum_am1 = 0.f;
    const float psiSt = digammaf_pos(St);
    for (int c = lane; c < C; c += 32) {
        const float p = a[c] / S, y = (c == t) ? 1.f : 0.f, d = p - y;
        mse = fmaf(d, d, mse); dot = fmaf(d, p, dot);
        const float at_c = (c == t) ? 1.f : a[c];                // α̃_c
        lg += lgammaf(at_c);
        term += (at_c - 1.f) * (digammaf_pos(at_c) - psiSt);
        sum_am1 += at_c - 1.f;
    }
    mse = warp_sum(mse); dot = warp_sum(dot); lg = warp_sum(lg); term = warp_sum(term); sum_am1 = warp_sum(sum_am1);
    const float kl = lgammaf(St) - lgammaf((float)C) - lg + term;
    const float invB = 1.f / (float)B, tri_St = trigammaf_pos(St);
    for (int c = lane; c < C; c += 32) {
        const float p = a[c] / S, y = (c == t) ? 1.f : 0.f;
        float gr = (2.f / S) * ((p - y) - dot);                   // d mse / d α_c
        if (c != t) gr += lam * ((a[c] - 1.f) * trigammaf_pos(a[c]) - tri_St * sum_am1);
        g[c] = gr * invB;
    }
    if (lane == 0) atomicAdd(loss_out, (mse + lam * kl) * invB);
}

// ---- fused softmax cross-entropy: forward + backward + running-loss accumulator, ONE launch ---------------------------
// Replaces log_softmax / nll_loss / their two backward kernels / the `loss_sum += loss` add of a training step.
// One CTA (the batch of a federated client is tens of rows): warp per row, block reduction, no zero-init launch, no atomics.
__global__ void __launch_bounds__(256) ce_loss_kernel(const float* __restrict__ logits, const long long* __restrict__ targets, int B, int C,
                                                      float* __restrict__ loss_out, float* __restrict__ loss_acc, float* __restrict__ grad) {
    __shared__ float wsum[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float invB = 1.f / (float)B;
    float local = 0.f;
    for (int r = warp; r < B; r += 8) {
        const float* z = logits + (size_t)r * C;
        float* g = grad + (size_t)r * C;
        const int t = (int)targets[r];
        float mx = -INFINITY;
        for (int c = lane; c < C; c += 32) mx = fmaxf(mx, z[c]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float se = 0.f;
        for (int c = lane; c < C; c += 32) se += __expf(z[c] - mx);
        se = warp_sum(se);
        const float lse = mx + __logf(se), inv = 1.f / se;
        for (int c = lane; c < C; c += 32) g[c] = (__expf(z[c] - mx) * inv - (c == t ? 1.f : 0.f)) * invB;
        if (lane == 0) local += lse - z[t];
    }
    if (lane == 0) wsum[warp] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) tot += wsum[w];
        tot *= invB;
        *loss_out = tot;
        if (loss_acc) *loss_acc += tot;
    }
}

// ---- mini-batch gather: xb[r] = X[perm[step·eb + r]], yb likewise; the last CTA advances the device-side step counter ----
// Replaces 2 index arithmetic kernels + 3 index_select + the counter increment of a training step.
__global__ void __launch_bounds__(256) gather_batch_kernel(const float* __restrict__ X, const long long* __restrict__ y,
                                                           const long long* __restrict__ perm, long long* __restrict__ step,
                                                           unsigned int* __restrict__ ticket, int eb, long long row_len,
                                                           float* __restrict__ xb, long long* __restrict__ yb) {
    const int r = blockIdx.y;
    const long long src = perm[*step * eb + r];
    const float* in = X + src * row_len;
    float* out = xb + (long long)r * row_len;
    if ((row_len & 3) == 0) {
        const long long n4 = row_len >> 2;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
            reinterpret_cast<float4*>(out)[i] = reinterpret_cast<const float4*>(in)[i];
    } else {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < row_len; i += (long long)gridDim.x * blockDim.x) out[i] = in[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) yb[r] = y[src];
    __syncthreads();                                          // every thread of this CTA has read *step
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int total = gridDim.x * gridDim.y;
        if (atomicAdd(ticket, 1u) == total - 1) { *ticket = 0u; *step += 1; }      // last CTA: all others already consumed *step
    }
}

}  // namespace mb

using torch::Tensor;
static inline cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream(); }

void sgd_step(Tensor live, int64_t stride, Tensor grad, int64_t gstride, int64_t slot0, int64_t nslots, int64_t n, double lr) {
    if (nslots == 0) return;
    c10::cuda::CUDAGuard guard(live.device());
    TORCH_CHECK(n % 4 == 0, "parameter region must be padded to a multiple of 4 floats");
    const int n4 = (int)(n / 4);
    dim3 grid(std::max(1, std::min((n4 + 255) / 256, (148 * 8) / (int)std::max<int64_t>(1, nslots))), (unsigned)nslots);
    mb::sgd_step_kernel<<<grid, 256, 0, cur_stream()>>>(live.data_ptr<float>() + (size_t)slot0 * stride, (size_t)stride,
                                                         grad.data_ptr<float>() + (size_t)slot0 * gstride, (size_t)gstride, n4, (float)lr);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// params[i] / grads[i]: dense tensors with identical physical layout (checked by the caller)
void sgd_multi(std::vector<Tensor> params, std::vector<Tensor> grads, double lr) {
    TORCH_CHECK(params.size() == grads.size());
    if (params.empty()) return;
    c10::cuda::CUDAGuard guard(params[0].device());
    struct Piece { float* p; const float* g; int64_t n; };
    std::vector<Piece> pieces;
    const int64_t max_piece = (int64_t)mb::kMtBlocks * mb::kMtChunk;          // one launch worth of one tensor
    for (size_t i = 0; i < params.size(); ++i) {
        const int64_t n = params[i].numel();
        TORCH_CHECK(n == grads[i].numel(), "sgd_multi: size mismatch");
        float* p = params[i].data_ptr<float>(); const float* g = grads[i].data_ptr<float>();
        for (int64_t off = 0; off < n; off += max_piece) pieces.push_back({p + off, g + off, std::min(max_piece, n - off)});
    }
    size_t i = 0;
    while (i < pieces.size()) {
        mb::MultiSgdArgs a;
        int nt = 0, nb = 0;
        while (i < pieces.size() && nt < mb::kMtTensors) {
            const int chunks = (int)((pieces[i].n + mb::kMtChunk - 1) / mb::kMtChunk);
            if (nb + chunks > mb::kMtBlocks) break;
            a.p[nt] = pieces[i].p; a.g[nt] = pieces[i].g; a.n[nt] = (int)pieces[i].n;
            for (int c = 0; c < chunks; ++c) { a.block_tensor[nb] = (unsigned char)nt; a.block_chunk[nb] = c; ++nb; }
            ++nt; ++i;
        }
        mb::sgd_multi_kernel<<<nb, 256, 0, cur_stream()>>>(a, (float)lr);
        C10_CUDA_KERNEL_LAUNCH_CHECK();
    }
}

Tensor bn_eval(Tensor x, Tensor mean, Tensor var, c10::optional<Tensor> gamma, c10::optional<Tensor> beta, double eps, bool relu,
               c10::optional<Tensor> residual) {
    c10::cuda::CUDAGuard guard(x.device());
    TORCH_CHECK(x.dtype() == torch::kFloat32 && x.dim() >= 2);
    const int C = (int)x.size(1);
    int inner;
    Tensor xin = x;
    if (x.dim() == 4 && x.is_contiguous(at::MemoryFormat::ChannelsLast) && !(x.size(2) == 1 && x.size(3) == 1)) inner = 1;
    else if (x.dim() == 4 && x.size(2) == 1 && x.size(3) == 1) { inner = 1; if (!x.is_contiguous() && !x.is_contiguous(at::MemoryFormat::ChannelsLast)) xin = x.contiguous(); }
    else { if (!x.is_contiguous()) xin = x.contiguous(); inner = 1; for (int d = 2; d < xin.dim(); ++d) inner *= (int)xin.size(d); }
    Tensor y = torch::empty_like(xin);
    const long long n = xin.numel();
    if (n == 0) return y;
    const float* res = nullptr;
    Tensor rin;
    if (residual.has_value() && residual->defined()) {              // must share xin's physical layout
        rin = residual->strides() == xin.strides() ? *residual : torch::empty_like(xin).copy_(*residual);
        res = rin.data_ptr<float>();
    }
    const int blocks = (int)std::min<long long>(148 * 8, (n / 4 + 255) / 256 + 1);
    mb::bn_eval_kernel<<<blocks, 256, 0, cur_stream()>>>(xin.data_ptr<float>(), y.data_ptr<float>(), mean.data_ptr<float>(), var.data_ptr<float>(),
        gamma.has_value() && gamma->defined() ? gamma->data_ptr<float>() : nullptr, beta.has_value() && beta->defined() ? beta->data_ptr<float>() : nullptr,
        (float)eps, n, C, inner, relu ? 1 : 0, res);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    return y;
}

void ce_eval(Tensor logits, Tensor targets, c10::optional<Tensor> n_valid, Tensor stats) {
    c10::cuda::CUDAGuard guard(logits.device());
    TORCH_CHECK(logits.is_contiguous() && logits.dtype() == torch::kFloat32 && targets.dtype() == torch::kInt64);
    const int B = (int)logits.size(0), C = (int)logits.size(1);
    if (B == 0) return;
    mb::ce_eval_kernel<<<(B * 32 + 255) / 256, 256, 0, cur_stream()>>>(logits.data_ptr<float>(), reinterpret_cast<const long long*>(targets.data_ptr<int64_t>()),
        n_valid.has_value() ? n_valid->data_ptr<int>() : nullptr, B, C, stats.data_ptr<float>());
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void dirichlet_eval(Tensor alpha, Tensor targets, c10::optional<Tensor> n_valid, Tensor stats) {
    c10::cuda::CUDAGuard guard(alpha.device());
    TORCH_CHECK(alpha.is_contiguous() && alpha.dtype() == torch::kFloat32 && targets.dtype() == torch::kInt64);
    const int B = (int)alpha.size(0), C = (int)alpha.size(1);
    if (B == 0) return;
    mb::dirichlet_eval_kernel<<<(B * 32 + 255) / 256, 256, 0, cur_stream()>>>(alpha.data_ptr<float>(), reinterpret_cast<const long long*>(targets.data_ptr<int64_t>()),
        n_valid.has_value() ? n_valid->data_ptr<int>() : nullptr, B, C, stats.data_ptr<float>());
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

std::vector<Tensor> evidential_loss_fwd_bwd(Tensor alpha, Tensor targets, double lam, c10::optional<Tensor> lam_t) {
    c10::cuda::CUDAGuard guard(alpha.device());
    TORCH_CHECK(alpha.is_contiguous() && alpha.dtype() == torch::kFloat32 && targets.dtype() == torch::kInt64);
    const int B = (int)alpha.size(0), C = (int)alpha.size(1);
    Tensor loss = torch::zeros({}, alpha.options());
    Tensor grad = torch::empty_like(alpha);
    if (B > 0)
        mb::evidential_loss_kernel<<<(B * 32 + 255) / 256, 256, 0, cur_stream()>>>(alpha.data_ptr<float>(), reinterpret_cast<const long long*>(targets.data_ptr<int64_t>()),
            B, C, (float)lam, lam_t.has_value() ? lam_t->data_ptr<float>() : nullptr, loss.data_ptr<float>(), grad.data_ptr<float>());
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    return {loss, grad};
}

// loss (0-dim), grad [B,C] (already divided by B).  `loss_acc` (optional 0-dim fp32) += loss inside the kernel.
std::vector<Tensor> ce_loss_fwd_bwd(Tensor logits, Tensor targets, c10::optional<Tensor> loss_acc) {
    c10::cuda::CUDAGuard guard(logits.device());
    TORCH_CHECK(logits.dim() == 2 && logits.is_contiguous() && logits.dtype() == torch::kFloat32 && targets.dtype() == torch::kInt64);
    const int B = (int)logits.size(0), C = (int)logits.size(1);
    TORCH_CHECK(B > 0 && targets.numel() == B);
    Tensor loss = torch::empty({}, logits.options());
    Tensor grad = torch::empty_like(logits);
    mb::ce_loss_kernel<<<1, 256, 0, cur_stream()>>>(logits.data_ptr<float>(), reinterpret_cast<const long long*>(targets.data_ptr<int64_t>()), B, C,
        loss.data_ptr<float>(), loss_acc.has_value() && loss_acc->defined() ? loss_acc->data_ptr<float>() : nullptr, grad.data_ptr<float>());
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    return {loss, grad};
}

// xb [eb, *X.shape[1:]], yb [eb]; advances `step` (0-dim int64) by one; `ticket` is a zero-initialised int32 scratch word.
std::vector<Tensor> gather_batch(Tensor X, Tensor y, Tensor perm, Tensor step, Tensor ticket, int64_t eb) {
    c10::cuda::CUDAGuard guard(X.device());
    TORCH_CHECK(X.is_contiguous() && X.dtype() == torch::kFloat32 && y.dtype() == torch::kInt64 && perm.dtype() == torch::kInt64 &&
                step.dtype() == torch::kInt64 && ticket.dtype() == torch::kInt32 && eb > 0);
    auto shape = X.sizes().vec();
    shape[0] = eb;
    Tensor xb = torch::empty(shape, X.options());
    Tensor yb = torch::empty({eb}, y.options());
    const int64_t row_len = X.numel() / std::max<int64_t>(1, X.size(0));
    const int chunks = (int)std::max<int64_t>(1, std::min<int64_t>(8, (row_len / 4 + 255) / 256));
    dim3 grid(chunks, (unsigned)eb);
    mb::gather_batch_kernel<<<grid, 256, 0, cur_stream()>>>(X.data_ptr<float>(), reinterpret_cast<const long long*>(y.data_ptr<int64_t>()),
        reinterpret_cast<const long long*>(perm.data_ptr<int64_t>()), reinterpret_cast<long long*>(step.data_ptr<int64_t>()),
        reinterpret_cast<unsigned int*>(ticket.data_ptr<int>()), (int)eb, (long long)row_len, xb.data_ptr<float>(),
        reinterpret_cast<long long*>(yb.data_ptr<int64_t>()));
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    return {xb, yb};
}
