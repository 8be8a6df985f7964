// Launch helper for programmatic dependent launch (see sm100_ptx.cuh: pdl_launch_dependents / pdl_wait).
// A kernel launched through launch_pdl may start while its predecessor in the stream is still draining; it MUST execute
// griddepcontrol.wait before its first global-memory access (the kernels below do so at their top, or right after their
// shared-memory / TMEM prologue).  DTB200_NO_PDL=1 launches the same kernels without the attribute (A/B switch).
#pragma once
#include <cstdlib>
#include <utility>
#include <cuda_runtime.h>

namespace dtb {
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool pdl = getenv("DTB200_NO_PDL") == nullptr;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}
}  // namespace dtb
