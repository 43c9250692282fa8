// CLIP-Gram style reward engine (forward loss + gradient with respect to the input image); see clip.cu.
#pragma once
#include <cuda_runtime.h>
#include <map>
#include <string>
#include <vector>

#include "engine.h"

namespace hedit {

struct ClipCfg {
  int resolution = 224, patch = 16, width = 768, heads = 12, layers = 3;    // layers = transformer blocks evaluated (features[2])
};

class ClipGram {
 public:
  explicit ClipGram(const ClipCfg& cfg);
  ~ClipGram();
  bool ok() const { return err_.empty(); }
  const std::string& error() const { return err_; }
  int load_tensor(const char* name, const float* src, const int64_t* dims, int ndim, cudaStream_t st);
  int finalize(std::string* missing);
  int tensor_count() const { return int(slots_.size()); }
  bool tensor_info(int i, std::string& name, std::vector<int64_t>& shape) const;
  // reference style image, already CLIP-normalised, [1][3][R][R] fp32 (device): stores its Gram matrix (base_clip.py:43-52,60-65)
  int set_reference(const float* ref, cudaStream_t st);
  // img [B][3][H][W] fp32 in [-1,1] (device) -> loss[B] = ||Gram(features) - Gram_ref||_F (device).  Keeps the activations for backward().
  int forward(const float* img, int B, int H, int W, float* loss, cudaStream_t st);
  // dLoss[b]/dimg -> dimg [B][3][H][W] (device) for the last forward()
  int backward(float* dimg, cudaStream_t st);
  long launches() const { return launches_; }
  std::string err_;

 private:
  struct Slot {
    enum Kind { F32, ROWS, ROWS_T };
    struct Dst { Kind kind; void* dst; int ld; int off; };
    std::vector<int64_t> shape;
    std::vector<Dst> dsts;
    bool loaded = false;
  };
  struct Lin { int O = 0, I = 0; op_t* w = nullptr; op_t* wt = nullptr; float* b = nullptr; };
  struct Block { float *ln1g = 0, *ln1b = 0, *ln2g = 0, *ln2b = 0; Lin in_proj, out_proj, fc, proj; };
  struct BlockSave { float* x_in = 0; float* x_mid = 0; float2 *st1 = 0, *st2 = 0; op_t* qkv = 0; float* P = 0; float* h = 0; };
  struct Taps { int* rowptr = nullptr; int* idx = nullptr; float* w = nullptr; };

  template <typename T> T* walloc(size_t n);
  template <typename T> T* A(size_t n);
  void reg(const std::string& name, std::vector<int64_t> shape, std::vector<Slot::Dst> dsts);
  void reg_lin(const std::string& wname, const std::string& bname, int O, int I, Lin& l);
  int gemm(const op_t* Ain, int lda, const op_t* Wt, int M, int N, int K, const GemmEpilogue& ep);
  int build_taps(int n_in, int n_out, Taps& fwd, Taps& bwd);
  int run_features(const float* img224, int B, float** feats_out);       // img224: [B][3][R][R] normalised
  int ensure_arena(int B, int H, int W);

  ClipCfg cfg_;
  int T_ = 0;                      // tokens incl. the class token
  std::map<std::string, Slot> slots_;
  std::vector<void*> owned_;
  Lin conv1_;
  float *cls_ = 0, *pos_ = 0, *lnpre_g_ = 0, *lnpre_b_ = 0, *stage_ = 0, *gref_ = 0, *mean_ = 0, *inv_std_ = 0;
  std::vector<Block> blocks_;
  Taps tx_f_, tx_b_, ty_f_, ty_b_; int taps_H_ = 0, taps_W_ = 0;
  // run state
  uint8_t* arena_ = nullptr; size_t arena_bytes_ = 0, top_ = 0;
  cudaStream_t st_ = 0;
  long launches_ = 0;
  bool have_ref_ = false, have_tape_ = false;
  int B_ = 0, H_ = 0, W_ = 0;
  float *x_tok_ = 0, *F_ = 0, *G_ = 0, *loss_ = 0; float2* st_pre_ = 0;
  std::vector<BlockSave> saves_;
  size_t fwd_top_ = 0;
};

}  // namespace hedit
