// murmura_b200 — programmatic dependent launch (PDL) for the fused training / scoring tapes (sm_100a).
//
// A federated SGD step at one node per GPU is ~110 dependent launches of 5–8 µs each; the launch latency and the prologue of
// kernel k+1 (barrier init, TMEM allocation, tensor-map prefetch, smem zero fill) are pure serial overhead.  Every kernel of the
// tape therefore (1) releases its dependents at entry (`griddepcontrol.launch_dependents`) and (2) blocks in
// `griddepcontrol.wait` AFTER its own prologue and BEFORE its first global-memory access: the wait returns once the predecessor
// grid has completed and flushed, so the data-flow is exactly the stream order, only the prologues overlap the previous tail.
// Under stream capture the attribute becomes a programmatic edge of the CUDA graph.
#pragma once
#include <cuda_runtime.h>

namespace mb {
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
}  // namespace mb

namespace mbhost {
inline bool& pdl_flag() { static bool on = true; return on; }

// One launch path for every kernel of the tapes: optional cluster shape + the programmatic-serialisation attribute.
template <typename... P, typename... A>
inline cudaError_t launch(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, dim3 cluster, A&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int n = 0;
    if (cluster.x * cluster.y * cluster.z > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = cluster.x; attr[n].val.clusterDim.y = cluster.y; attr[n].val.clusterDim.z = cluster.z;
        ++n;
    }
    if (pdl_flag()) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = attr; cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}
}  // namespace mbhost
