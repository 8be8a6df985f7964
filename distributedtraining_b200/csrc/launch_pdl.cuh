// Launch helper for programmatic dependent launch (see sm100_ptx.cuh: pdl_launch_dependents / pdl_wait).
// A kernel launched through launch_pdl may start while its predecessor in the stream is still draining; it MUST execute
// griddepcontrol.wait before its first global-memory access (the kernels below do so at their top, or right after their
// shared-memory / TMEM prologue).  The attribute is only set with DTB200_PDL_ALL=1 (the GEMMs always use PDL unless DTB200_NO_PDL=1).
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
  // measured (profiles/README.md): PDL on the GEMMs alone -0.9 % of the step, on GEMMs + these small kernels -0.1 % -> opt-in here
  static const bool pdl = getenv("DTB200_NO_PDL") == nullptr && getenv("DTB200_PDL_ALL") != nullptr;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}
}  // namespace dtb
