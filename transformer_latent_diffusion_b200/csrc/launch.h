// Kernel launch helper: cudaLaunchKernelEx with the programmatic-stream-serialization (PDL) attribute, so a kernel's
// prologue (barrier init, TMEM alloc, descriptor prefetch, constant loads) overlaps the tail of its predecessor inside
// the per-step CUDA graph.  Every kernel launched through this helper executes griddepcontrol.wait before it touches
// data produced by earlier kernels.
#pragma once
#include "common.h"

namespace tld {

bool pdl_enabled();

template <typename... KArgs, typename... Args>
int launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  TLD_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
  return 0;
}

}  // namespace tld
