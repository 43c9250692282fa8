// Kernel launch helper with programmatic dependent launch (PDL).
//
// A UNet forward is a chain of ~440 dependent kernels; at small batches (the reference's one-image-per-call signature) most of them run
// for 5-20 us, so the launch-to-launch bubble and every kernel's own prologue (barrier init, tensor-memory allocation, descriptor
// prefetch, CTA scheduling) are a large share of the step.  Kernels launched through launch_k carry
// cudaLaunchAttributeProgrammaticStreamSerialization: the grid may become resident as soon as every CTA of the previous kernel has
// executed `griddepcontrol.launch_dependents`, runs its prologue, and blocks in `griddepcontrol.wait` (ptx.cuh: pdl_wait) until the
// previous grid has completed and its memory is visible.  Every kernel launched this way executes pdl_wait before its first global
// memory access and pdl_launch right after it, so at most one dependent grid is resident early and the chain stays transitively ordered.
// Stream capture keeps these as programmatic graph edges.
//
// MEASURED (round 2, B200, loop replayed from CUDA graphs): 2.50 images/s with PDL against 2.57 without at batch 8 (-2.6 %: the early
// resident CTAs of the next kernel compete with the running kernel's last wave), 1.475 against 1.460 images/s at batch 1 (+1 %).  Graph
// replay already removes most of the launch bubble, so PDL is OFF by default; HEDIT_PDL=1 enables it (results are bit-identical).
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

namespace hedit {

inline bool pdl_enabled() {
  static const bool v = getenv("HEDIT_PDL") && atoi(getenv("HEDIT_PDL")) != 0;
  return v;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  if (pdl_enabled()) { cfg.attrs = at; cfg.numAttrs = 1; }
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace hedit
