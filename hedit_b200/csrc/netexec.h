// Shared executor pieces of the convolutional engines that run outside the UNet's static plan (VAE decoder, face-swapping DDPM UNet):
// a bump arena with a sizing ("dry") pass, GEMM / implicit-GEMM conv launch helpers and the GroupNorm forward with statistics taken
// from the producing GEMM's epilogue.
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <array>
#include <functional>
#include <string>
#include <vector>

#include "engine.h"

namespace hedit {

class NetExec {
 public:
  std::string err_;
  long launches() const { return launches_; }
  double flops() const { return flops_; }
  size_t arena_bytes() const { return arena_bytes_; }

 protected:
  ~NetExec();
  template <typename T> T* A(size_t n) {
    const size_t bytes = (n * sizeof(T) + 1023) & ~size_t(1023);
    const size_t off = top_;
    top_ += bytes;
    peak_ = std::max(peak_, top_);
    return reinterpret_cast<T*>(arena_ + off);      // arena_ is null in the sizing pass: offsets only, nothing is launched
  }
  // (re)allocate the arena so that `need` bytes fit
  int reserve(size_t need, const char* what);
  int gemm(const op_t* Ain, int lda, int mode, const ConvGeom* cg, const op_t* Wt, int M, int N, int K, const GemmEpilogue& ep, int ldw = 0);
  // 3x3 conv, stride 1 pad 1, or stride 2 (H, W = OUTPUT dims) with padding 1 / (0,1,0,1)
  int conv3(const op_t* x, const op_t* w, int S, int H, int W, int cin, int cout, GemmEpilogue ep, int stride = 1, int pad01 = 0);
  float2* colstats_for(int M, int N, int HW);
  // GroupNorm(+SiLU) of the channel concatenation [x1 | x2] (x2 may be null) -> 16-bit operand (+ raw 16-bit copy); statistics from
  // the producers' colstats when both are present, else a separate statistics pass.  *stats_out = (mean, rstd) per (sample, group).
  int gn_fwd(const float* x1, const float2* cs1, int C1, const float* x2, const float2* cs2, int C2, int S, int HW, const float* g,
             const float* b, float eps, int silu, op_t* out, op_t* raw, float2** stats_out);

  // GroupNorm -> fused q|k|v projection -> single-head attention over N tokens (scores and P.V on the GEMM, probabilities
  // materialised per sample) -> output projection + residual.  Optional outputs feed a backward pass.
  int attn1h_fwd(const float* x, const float2* cs_x, int S, int N, int C, const float* gng, const float* gnb, float eps, const op_t* w_qkv,
                 const float* b_qkv, const op_t* w_o, const float* b_o, float** out, float2** cs_out, float2** gn_stats_out = nullptr,
                 const op_t** qkv_out = nullptr, const op_t** P_out = nullptr);

  // nearest-2x upsample -> 3x3 conv (bias) as four 2x2 phase convs on the coarse grid (weights from cvt_upconv_phases_kernel):
  // x fp32 [S][H][W][C] -> *out fp32 [S][2H][2W][C] with column statistics in *cs_out
  int upconv_fused(const float* x, const op_t* w_phases, const float* bias, int S, int H, int W, int C, float** out, float2** cs_out);

  // CUDA-graph replay of a whole pass.  `body` must be a fixed sequence of launches on st_ determined by `key` (buffer addresses,
  // batch) and by device-resident data only.  A key's first call launches directly; its second call is captured (on a private
  // stream: the caller's may be the legacy default stream) and instantiated; later calls are one cudaGraphLaunch on `st`.
  // HEDIT_NET_GRAPH=0 always launches directly.  Graphs are dropped when the arena or any buffer they address is reallocated.
  using GraphKey = std::array<uintptr_t, 4>;
  bool replay(const GraphKey& key, cudaStream_t st);                 // true: a stored graph was launched
  int run_or_capture(const GraphKey& key, cudaStream_t st, const std::function<int()>& body);
  void drop_graphs();

  int groups_ = 32;
  uint8_t* arena_ = nullptr;
  size_t arena_bytes_ = 0, top_ = 0, peak_ = 0;
  bool dry_ = false;
  cudaStream_t st_ = 0;
  long launches_ = 0;
  double flops_ = 0;

 private:
  struct GraphEntry { GraphKey key; cudaGraphExec_t exec; long launches; double flops; bool bad; };
  std::vector<GraphEntry> graphs_;
  cudaStream_t cap_stream_ = nullptr;
};

}  // namespace hedit
