// NVTX ranges around the host-side phases (nsys / ncu --nvtx): one edit, every timestep, every UNet launch, the reward networks.
// Header-only NVTX3: a no-op (one pointer test) unless a profiler injects its library.
#pragma once
#include <nvtx3/nvToolsExt.h>

#include <cstdio>

namespace hedit {
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  NvtxRange(const char* fmt, int a, int b) { char buf[96]; snprintf(buf, sizeof buf, fmt, a, b); nvtxRangePushA(buf); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
}  // namespace hedit
