// murmura_b200 — pairwise-distance Gram matrix on 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Replaces the all-pairs L2 loop of Krum (reference murmura/aggregation/krum.py:55-62 via
// murmura/aggregation/base.py:118-135: m² per-key torch.norm calls with .item() host syncs).
//
//   G = X·Xᵀ,  X ∈ R^{R×P}  (R ≤ 128 model states, P up to 60 M parameters), dist²_ab = G_aa + G_bb − 2·G_ab
//
// This is a skinny-M/N, enormous-K GEMM, i.e. purely bandwidth bound: every fp32 parameter tile is pulled
// ONCE (from local HBM or from a peer GPU's published buffer over NVLink — the tensor maps point at
// peer-mapped addresses) by TMA straight into 128B-swizzled shared memory and consumed in place by
// `tcgen05.mma.kind::tf32` as BOTH the A and the B operand (same smem descriptor), accumulating a
// 128×N fp32 tile in TMEM.  Split-K over a persistent grid (one CTA per SM), partial tiles are merged
// with fp32 reductions into the R×R result.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected
// lane), warps 2..5 = epilogue (tcgen05.ld → red.global.add).
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mb {

constexpr int kGramRows = 128;                 // UMMA M (tile rows; unused rows are ignored)
constexpr int kGramKB = 32;                    // fp32 elements per k-block = one 128-byte swizzle row
constexpr int kMaxStages = 8;
constexpr int kMaxGroups = 16;                 // 8-row groups per tile (16 × 8 = 128 rows)
constexpr int kMaxMaps = 16;
constexpr int kGramThreads = 192;
constexpr int kSmemBudget = 200 * 1024;

// The tile is assembled from 8-row groups; group g = rows [8g, 8g+8) comes from tensor map `map[g]` at row `y[g]`.
// One TMA instruction fetches ONE group for `kb_per_stage` consecutive k-blocks (3-D box {32 floats, 8 rows, KB}),
// landing as [KB][8 rows][128 B]; the stage is laid out [group][KB][8][128 B], so for a fixed k-block the 8-row
// groups are a constant KB·1024 B apart — exactly the SBO of the UMMA shared-memory descriptor.
struct GramMaps { CUtensorMap m[kMaxMaps]; };
struct GramGroups { int n; int kb_per_stage; int stages; int map[kMaxGroups]; int y[kMaxGroups]; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    uint32_t spins = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) break;
        if (++spins > (1u << 22)) __trap();        // never hang the GPU: a lost arrival aborts the kernel instead
    }
}
// Long wait (epilogue warps idle for the whole main loop): back off so the pollers do not steal issue slots
// from the TMA / MMA warps that share their SM sub-partitions.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(256);
        if (++spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

// K-major operand, 128-byte swizzle: rows are 128 B, 8-row groups are `sbo_bytes` apart, descriptor version 1.
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t smem_addr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);            // start address  [0,14)
    d |= (uint64_t)0 << 16;                                  // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;       // stride byte offset [32,46)
    d |= (uint64_t)1 << 46;                                  // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                                  // layout: SWIZZLE_128B
    return d;
}
// Instruction descriptor for kind::tf32, fp32 accumulate, A and B K-major.
__host__ __device__ constexpr uint32_t make_tf32_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// Thread-block clusters: the `csize` CTAs of a cluster own adjacent K-slices and merge their partial R×R tiles through
// distributed shared memory (red.shared::cluster into the leader CTA's smem) before ONE CTA per cluster touches global
// memory — csize× fewer global fp32 reductions on the R×R result.
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void dsmem_red_add(uint32_t local_smem_addr, uint32_t target_cta, float v) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_smem_addr), "r"(target_cta));
    asm volatile("red.relaxed.cluster.shared::cluster.add.f32 [%0], %1;" :: "r"(remote), "f"(v) : "memory");
}

__global__ void __launch_bounds__(kGramThreads, 1)
gram_tf32_kernel(const __grid_constant__ GramMaps maps, const GramGroups grp, int kb0, int kb1, int n_mma, int R,
                 float* __restrict__ out /*[128][128]*/) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
    __shared__ __align__(8) uint64_t accum_bar;
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KB = grp.kb_per_stage, NS = grp.stages;
    const uint32_t group_bytes = (uint32_t)KB * 1024u;               // one group's slab inside a stage
    const uint32_t stage_bytes = (uint32_t)grp.n * group_bytes;
    // split-K: contiguous chunk of k-block *super-steps* (KB k-blocks each) per CTA
    const int total_steps = (kb1 - kb0 + KB - 1) / KB;
    const int per = (total_steps + (int)gridDim.x - 1) / (int)gridDim.x;
    const int s0 = (int)blockIdx.x * per;
    const int s1 = min(total_steps, s0 + per);
    const int num_steps = max(0, s1 - s0);

    if (warp == 0 && lane == 0) {
        for (int g = 0; g < grp.n; ++g)
            asm volatile("prefetch.tensormap [%0];" :: "l"(&maps.m[grp.map[g]]) : "memory");
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
            mbar_init(&accum_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_smem)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    if (num_steps > 0) {
        if (warp == 0) {
            // ===== TMA producer: one 3-D bulk tensor load per 8-row group per stage =====
            if (lane == 0) {
                int stage = 0; uint32_t phase = 0;
                for (int st = s0; st < s1; ++st) {
                    mbar_wait(&empty_bar[stage], phase ^ 1u);
                    mbar_expect_tx(&full_bar[stage], stage_bytes);          // OOB k-blocks / rows are zero-filled and still counted
                    uint8_t* tile = smem + (size_t)stage * stage_bytes;
                    const int kb = kb0 + st * KB;
                    for (int g = 0; g < grp.n; ++g)
                        tma_load_3d(tile + (size_t)g * group_bytes, &maps.m[grp.map[g]], 0, grp.y[g], kb, &full_bar[stage]);
                    if (++stage == NS) { stage = 0; phase ^= 1u; }
                }
            }
        } else if (warp == 1) {
            // ===== MMA issuer (single elected lane): 4·KB tf32 MMAs per stage, A and B are the SAME smem tile =====
            if (lane == 0) {
                const uint32_t idesc = make_tf32_idesc(kGramRows, n_mma);
                int stage = 0; uint32_t phase = 0;
                for (int st = s0; st < s1; ++st) {
                    mbar_wait(&full_bar[stage], phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t tile_addr = smem_u32(smem + (size_t)stage * stage_bytes);
                    const int kb_here = min(KB, kb1 - (kb0 + st * KB));      // tail stage: skip zero-filled k-blocks
                    for (int j = 0; j < kb_here; ++j) {
#pragma unroll
                        for (int k = 0; k < kGramKB / 8; ++k) {              // UMMA_K = 8 (tf32) → 32-byte steps inside the swizzle row
                            const uint64_t desc = make_sw128_kmajor_desc(tile_addr + j * 1024 + k * 32, group_bytes);
                            umma_tf32(tmem_base, desc, desc, idesc, (st > s0 || j > 0 || k > 0) ? 1u : 0u);
                        }
                    }
                    umma_commit(&empty_bar[stage]);                          // frees the smem slot once these MMAs retire
                    if (++stage == NS) { stage = 0; phase ^= 1u; }
                }
                umma_commit(&accum_bar);                                     // accumulator complete
            }
        }
    }
    // ===== epilogue: TMEM → registers → (cluster DSMEM reduce) → fp32 reductions into the global R×R tile =====
    // The pipeline stages are dead once every CTA of the cluster finished its main loop, so the leader's first stage
    // is reused as the R×R (<= 64 KiB) reduction buffer.
    const uint32_t csize = cluster_nctarank(), crank = cluster_ctarank();
    float* red = reinterpret_cast<float*>(smem);
    if (num_steps > 0) {                                                 // this CTA's MMAs have retired → its smem stages are dead
        mbar_wait_backoff(&accum_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (csize > 1) {
        cluster_sync_all();                                              // every CTA of the cluster is past its main loop
        if (crank == 0)
            for (int i = threadIdx.x; i < R * R; i += blockDim.x) red[i] = 0.f;
        cluster_sync_all();
    }
    if (warp >= 2) {
        const int q = warp & 3;                                          // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        for (int c0 = 0; c0 < n_mma && num_steps > 0; c0 += 16) {
            uint32_t r[16];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                           "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row < R) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (c0 + j >= R) break;
                    const float v = __uint_as_float(r[j]);
                    if (csize > 1) dsmem_red_add(smem_u32(red + row * R + c0 + j), 0u, v);
                    else atomicAdd(out + row * kGramRows + c0 + j, v);
                }
            }
        }
    }
    if (csize > 1) {
        cluster_sync_all();                                              // every CTA's partial tile has landed in the leader
        if (crank == 0)
            for (int i = threadIdx.x; i < R * R; i += blockDim.x)
                atomicAdd(out + (i / R) * kGramRows + (i % R), red[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(128));
    }
}

}  // namespace mb

using torch::Tensor;

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        TORCH_CHECK(e == cudaSuccess && q == cudaDriverEntryPointSuccess && p, "cuTensorMapEncodeTiled unavailable");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
}  // namespace

// One 3-D fp32 tensor map per source buffer [rows][row_len] (row_stride elements between rows), viewed as
// {32 floats, rows, row_len/32}: box = {32, 8 rows, kb_per_stage k-blocks}, 128-byte swizzle.
// Returns a CPU byte tensor [nbuf][128].
Tensor gram_make_maps(std::vector<int64_t> base_ptrs, int64_t rows, int64_t row_stride, int64_t row_len, int64_t kb_per_stage) {
    TORCH_CHECK((int)base_ptrs.size() <= mb::kMaxMaps, "at most ", mb::kMaxMaps, " source buffers");
    TORCH_CHECK(kb_per_stage >= 1 && kb_per_stage <= 16);
    TORCH_CHECK(row_stride % 4 == 0 && row_len % 32 == 0, "row stride must be 16-byte aligned and row_len a multiple of 32");
    Tensor t = torch::zeros({(int64_t)base_ptrs.size(), (int64_t)sizeof(CUtensorMap)}, torch::kUInt8);
    auto enc = get_encode();
    for (size_t i = 0; i < base_ptrs.size(); ++i) {
        cuuint64_t gdim[3] = {(cuuint64_t)mb::kGramKB, (cuuint64_t)rows, (cuuint64_t)(row_len / mb::kGramKB)};
        cuuint64_t gstride[2] = {(cuuint64_t)row_stride * 4, (cuuint64_t)mb::kGramKB * 4};
        cuuint32_t box[3] = {(cuuint32_t)mb::kGramKB, 8u, (cuuint32_t)kb_per_stage};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(reinterpret_cast<CUtensorMap*>(t.data_ptr<uint8_t>() + i * sizeof(CUtensorMap)),
                         CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, reinterpret_cast<void*>(base_ptrs[i]), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        TORCH_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code ", (int)r);
    }
    return t;
}

// Stage geometry shared by the host wrappers: k-blocks per TMA instruction and pipeline depth for `ngroups` groups.
// UMMA M is fixed at 128 rows = 16 groups, so the A operand of the LAST stage reads up to 16 group slabs past its
// start: the allocation carries (16 - ngroups) slabs of padding after the last stage (rows >= R are never read back).
static int64_t gram_smem_bytes(int ngroups, int kb, int stages) {
    return ((int64_t)stages * ngroups + (mb::kMaxGroups - ngroups)) * kb * 1024 + 1024;
}
static void gram_geometry(int ngroups, int* kb_per_stage, int* stages) {
    int best_kb = 1, best_st = 2; int64_t best_inflight = 0;
    for (int kb = 8; kb >= 1; kb >>= 1) {
        const int64_t avail = mb::kSmemBudget - (int64_t)(mb::kMaxGroups - ngroups) * kb * 1024;
        int st = (int)std::min<int64_t>(mb::kMaxStages, avail / ((int64_t)ngroups * kb * 1024));
        if (st < 2) continue;
        if (st >= 4) { best_kb = kb; best_st = st; break; }             // deep enough: take the largest TMA box
        const int64_t inflight = (int64_t)st * ngroups * kb * 1024;
        if (inflight > best_inflight) { best_inflight = inflight; best_kb = kb; best_st = st; }
    }
    *kb_per_stage = best_kb;
    *stages = best_st;
}

int64_t gram_kb_per_stage(int64_t ngroups) {
    int kb, st;
    gram_geometry((int)ngroups, &kb, &st);
    return kb;
}

// out[128*128] (+)= Gram of the tile rows (8-row groups `group_map`/`group_y`) over k-blocks [kb0, kb1) (32 floats each).
void gram_tf32(Tensor maps_cpu, std::vector<int64_t> group_map, std::vector<int64_t> group_y, int64_t kb0, int64_t kb1,
               int64_t R, Tensor out, bool zero_out, int64_t max_ctas) {
    c10::cuda::CUDAGuard guard(out.device());
    TORCH_CHECK(out.numel() == mb::kGramRows * mb::kGramRows && out.dtype() == torch::kFloat32 && out.is_contiguous());
    TORCH_CHECK(group_map.size() == group_y.size() && (int)group_map.size() <= mb::kMaxGroups && !group_map.empty());
    const int ngroups = (int)group_map.size();
    const int rows_total = ngroups * 8;
    TORCH_CHECK(R <= rows_total);
    mb::GramMaps maps;
    memset(&maps, 0, sizeof(maps));
    memcpy(&maps, maps_cpu.data_ptr<uint8_t>(), (size_t)maps_cpu.numel());
    mb::GramGroups grp;
    grp.n = ngroups;
    gram_geometry(ngroups, &grp.kb_per_stage, &grp.stages);
    for (int g = 0; g < ngroups; ++g) { grp.map[g] = (int)group_map[g]; grp.y[g] = (int)group_y[g]; }
    const int n_mma = std::max(16, (rows_total + 15) / 16 * 16);
    auto stream = at::cuda::getCurrentCUDAStream();
    if (zero_out) cudaMemsetAsync(out.data_ptr<float>(), 0, out.numel() * sizeof(float), stream);
    if (kb1 <= kb0) return;
    const int smem = (int)gram_smem_bytes(ngroups, grp.kb_per_stage, grp.stages);
    static int attr_smem = 0;
    if (smem > attr_smem) {
        cudaFuncSetAttribute(mb::gram_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_smem = smem;
    }
    int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
    if (max_ctas > 0) sms = std::min<int>(sms, (int)max_ctas);
    const int64_t steps = (kb1 - kb0 + grp.kb_per_stage - 1) / grp.kb_per_stage;
    int grid = (int)std::max<int64_t>(1, std::min<int64_t>(sms, (steps + 3) / 4));      // >= 4 super-steps per CTA
    // Cluster of 2 CTAs (a TPC pair: 148 = 2·74 packs every SM; clusters of 4 strand SMs on 16/18/20-SM GPCs and cost a
    // second wave at one CTA per SM).  The grid is clamped to the number of clusters that can be co-resident.
    int csize = (grid >= 2 && R * R * 4 <= smem - 1024) ? 2 : 1;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(mb::kGramThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (csize > 1) {
        cfg.gridDim = dim3(grid / csize * csize);
        int max_clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&max_clusters, mb::gram_tf32_kernel, &cfg) == cudaSuccess && max_clusters > 0)
            grid = std::min(grid, max_clusters * csize);
        else
            (void)cudaGetLastError();
    }
    grid = std::max(csize, grid / csize * csize);
    cfg.gridDim = dim3(grid);
    float* outp = out.data_ptr<float>();
    int kb0i = (int)kb0, kb1i = (int)kb1, Ri = (int)R;
    cudaError_t err = cudaLaunchKernelEx(&cfg, mb::gram_tf32_kernel, maps, grp, kb0i, kb1i, n_mma, Ri, outp);
    TORCH_CHECK(err == cudaSuccess, "gram_tf32 launch failed: ", cudaGetErrorString(err));
}
