// Programmatic dependent launch (PDL) for the chains of small kernels that make up a step on a small or multi-level hierarchy
// (BASELINE configs 1, 2, 4: 126 to ~500 dependent launches per step, each a few microseconds of work, so the step is the sum
// of launch-to-launch latencies).  A kernel launched through vrt_launch() may be scheduled while its predecessor in the stream
// (or graph branch) is still running: its CTAs become resident and stop at griddepcontrol.wait until the predecessor has
// completed and its writes are visible, so the dependency itself is unchanged — only the launch latency moves off the
// critical path.  Every kernel launched this way must call vrt_pdl_sync() before its first global-memory access.
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>
#include <utility>

__device__ __forceinline__ void vrt_pdl_sync() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the successor may be scheduled from here on
    asm volatile("griddepcontrol.wait;" ::: "memory");                // the predecessor is complete, its memory visible
}

inline bool vrt_pdl_enabled() {            // VRT_PDL=0: plain stream-ordered launches (for A/B and debugging)
    static const bool on = !(getenv("VRT_PDL") && atoi(getenv("VRT_PDL")) == 0);
    return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t vrt_launch_smem(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = vrt_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}
// the same for a kernel whose CTAs form one thread-block cluster of `cluster_x` CTAs (hardware cluster barrier between its phases)
template <typename... KArgs, typename... Args>
inline cudaError_t vrt_launch_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, unsigned cluster_x, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_x; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = vrt_pdl_enabled() ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t vrt_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t stream, Args&&... args) {
    return vrt_launch_smem(kernel, grid, block, 0, stream, std::forward<Args>(args)...);
}
