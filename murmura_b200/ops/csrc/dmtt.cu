// murmura_b200 — dynamic-topology and DMTT trust kernels (sm_100a).
//
// Reference call sites replaced: K13 mobility adjacency (murmura/topology/dynamic.py:63-105, an
// O(N²) Python loop re-run in every process), K11 falsified claims (murmura/attacks/topology_liar.py:78-102),
// K12 trust-state update + Top-B (murmura/dmtt/state.py:53-142, murmura/dmtt/node_process.py:215-241,369-395).
// Positions are generated on the host with NumPy's PCG64 (bit-exact with the reference) and uploaded once
// for all rounds; everything per-round happens on the device so the round loop never returns to the host.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

namespace mb {

__device__ __forceinline__ double torus_dist(const double* pos, int i, int j, double area) {
    double dx = fabs(pos[2 * i] - pos[2 * j]), dy = fabs(pos[2 * i + 1] - pos[2 * j + 1]);
    dx = fmin(dx, area - dx); dy = fmin(dy, area - dy);
    return sqrt(dx * dx + dy * dy);
}

// adj[i][j] = dist(i,j) < range (float64 so boundary decisions match NumPy); isolated nodes are then
// linked to their nearest peer, sequentially in node order (later nodes see earlier repairs).
__global__ void mobility_adjacency_kernel(const double* __restrict__ pos_all, int N, int round, double area, double range,
                                          int ensure_connected, uint8_t* __restrict__ adj) {
    const double* pos = pos_all + (size_t)round * N * 2;
    for (int p = threadIdx.x; p < N * N; p += blockDim.x) {
        const int i = p / N, j = p % N;
        adj[p] = (i != j && torus_dist(pos, i, j, area) < range) ? 1 : 0;
    }
    __syncthreads();
    if (ensure_connected && threadIdx.x == 0) {
        for (int i = 0; i < N; ++i) {
            bool any = false;
            for (int j = 0; j < N && !any; ++j) any = adj[i * N + j] != 0;
            if (any) continue;
            int best = -1; double bd = 1e300;
            for (int j = 0; j < N; ++j) if (j != i) { const double d = torus_dist(pos, i, j, area); if (d < bd) { bd = d; best = j; } }
            if (best >= 0) { adj[i * N + best] = 1; adj[best * N + i] = 1; }
        }
    }
}

__global__ void liar_claims_kernel(const uint8_t* __restrict__ adj, const uint8_t* __restrict__ is_liar, int N,
                                   uint8_t* __restrict__ claims) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N * N) return;
    const int i = p / N, j = p % N;
    claims[p] = (adj[p] || (i != j && is_liar[i] && is_liar[j])) ? 1 : 0;
}

// One block per local node i (global id gids[blockIdx.x]).
//   gids    [V]     global ids of the local nodes (any placement)
//   collab  [N][N]  C^{t-1}: row i = peers i sent to / expected from (all ranks' rows, gathered)
//   received[V][N]  states/claims that actually arrived at i this round
//   claims  [N][N]  claimed neighbourhoods; adj [N][N] ground-truth G^t
//   model_score / score_valid [V][N]  s_ij^{model} of received models
// Updates ĉ, α, β rows of i and emits C_i^t = TopB_j q_ij over the true G^t neighbours.
__global__ void dmtt_update_kernel(const uint8_t* __restrict__ adj, const uint8_t* __restrict__ claims,
                                   const uint8_t* __restrict__ collab, const uint8_t* __restrict__ received,
                                   const float* __restrict__ model_score, const uint8_t* __restrict__ score_valid,
                                   float* __restrict__ c_hat, float* __restrict__ alpha, float* __restrict__ beta,
                                   uint8_t* __restrict__ next_collab, float* __restrict__ q_out, int N, const int* __restrict__ gids,
                                   float rho, float lam, float w_d, float w_x, float tau_U, float eta,
                                   float l1, float l2, float l3, int B) {
    extern __shared__ float s_q[];                  // [N]
    const int vi = blockIdx.x, i = gids[vi];
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        const size_t ij = (size_t)vi * N + j;
        float c = c_hat[ij], a = alpha[ij], b = beta[ij];
        if (collab[(size_t)i * N + j]) c = (1.f - rho) * c + rho * (received[ij] ? 1.f : 0.f);      // Algorithm 1: link EMA
        if (received[ij]) {                                                                         // Algorithm 4: Beta evidence
            float d = 0.f, x = 0.f;
            for (int u = 0; u < N; ++u) if (claims[(size_t)j * N + u]) { if (adj[(size_t)j * N + u]) d += 1.f; else x += 1.f; }
            a = fmaxf(0.01f, lam * a + w_d * d);
            b = fmaxf(0.01f, lam * b + w_x * x);
        }
        c_hat[ij] = c; alpha[ij] = a; beta[ij] = b;
        const float s = a + b, R = a / s, U = sqrtf(fmaxf(0.f, a * b / (s * s * (s + 1.f))));
        const float T = R * __expf(-eta * fmaxf(0.f, U - tau_U));
        const float sm = score_valid[ij] ? model_score[ij] : 0.5f;
        const float q = l1 * sm + l2 * T + l3 * c;
        q_out[ij] = q;
        s_q[j] = adj[(size_t)i * N + j] ? q : -INFINITY;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < N; j += blockDim.x) {                  // Top-B, descending, ties keep candidate order
        uint8_t pick = 0;
        if (s_q[j] != -INFINITY) {
            int rank = 0;
            for (int u = 0; u < N; ++u) if (u != j && s_q[u] != -INFINITY && (s_q[u] > s_q[j] || (s_q[u] == s_q[j] && u < j))) ++rank;
            pick = rank < B ? 1 : 0;
        }
        next_collab[(size_t)vi * N + j] = pick;
    }
}

}  // namespace mb

using torch::Tensor;
static inline cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream(); }

void mobility_adjacency(Tensor pos, int64_t round, double area, double range, bool ensure_connected, Tensor adj) {
    c10::cuda::CUDAGuard guard(pos.device());
    TORCH_CHECK(pos.dtype() == torch::kFloat64 && pos.dim() == 3 && adj.dtype() == torch::kUInt8);
    const int N = (int)pos.size(1);
    mb::mobility_adjacency_kernel<<<1, 256, 0, cur_stream()>>>(pos.data_ptr<double>(), N, (int)round, area, range,
                                                               ensure_connected ? 1 : 0, adj.data_ptr<uint8_t>());
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void liar_claims(Tensor adj, Tensor is_liar, Tensor claims) {
    c10::cuda::CUDAGuard guard(adj.device());
    const int N = (int)adj.size(0);
    mb::liar_claims_kernel<<<(N * N + 255) / 256, 256, 0, cur_stream()>>>(adj.data_ptr<uint8_t>(), is_liar.data_ptr<uint8_t>(), N,
                                                                          claims.data_ptr<uint8_t>());
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void dmtt_update(Tensor adj, Tensor claims, Tensor collab, Tensor received, Tensor model_score, Tensor score_valid, Tensor c_hat,
                 Tensor alpha, Tensor beta, Tensor next_collab, Tensor q_out, double rho, double lam, double w_d, double w_x,
                 double tau_U, double eta, double l1, double l2, double l3, int64_t B, Tensor gids) {
    c10::cuda::CUDAGuard guard(adj.device());
    const int N = (int)adj.size(0), V = (int)c_hat.size(0);
    if (V == 0) return;
    mb::dmtt_update_kernel<<<V, 64, N * sizeof(float), cur_stream()>>>(
        adj.data_ptr<uint8_t>(), claims.data_ptr<uint8_t>(), collab.data_ptr<uint8_t>(), received.data_ptr<uint8_t>(),
        model_score.data_ptr<float>(), score_valid.data_ptr<uint8_t>(), c_hat.data_ptr<float>(), alpha.data_ptr<float>(),
        beta.data_ptr<float>(), next_collab.data_ptr<uint8_t>(), q_out.data_ptr<float>(), N, gids.data_ptr<int>(), (float)rho, (float)lam,
        (float)w_d, (float)w_x, (float)tau_U, (float)eta, (float)l1, (float)l2, (float)l3, (int)B);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}
