// VAE encoder engine (forward only); see vae_enc.cu.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "netexec.h"
#include "vae.h"

namespace hedit {

class VaeEncoder : public NetExec {
 public:
  explicit VaeEncoder(const VaeCfg& cfg);
  ~VaeEncoder();
  bool ok() const { return err_.empty(); }
  const std::string& error() const { return err_; }
  int load_tensor(const char* name, const float* src, const int64_t* dims, int ndim, cudaStream_t st);
  int finalize(std::string* missing);
  // img [B][3][H][W] fp32 NCHW (device) -> moments [B][2*latent][H/8][W/8] (mean | logvar), device fp32
  int encode(const float* img, float* moments, int B, int H, int W, cudaStream_t st);

 private:
  struct Slot { std::vector<int64_t> shape; int kind = 0; void* dst = nullptr; int ld = 0, off = 0; bool loaded = false; };   // kind: 0 f32, 1 conv3, 2 rows, 3 conv_in
  struct Conv3W { op_t* w = nullptr; float* b = nullptr; };
  struct ResW { int cin = 0, cout = 0; float *n1g = 0, *n1b = 0, *n2g = 0, *n2b = 0, *bsc = 0; Conv3W c1, c2; op_t* wsc = 0; };
  template <typename T> T* walloc(size_t n);
  void reg(const std::string& name, std::vector<int64_t> shape, int kind, void* dst, int ld, int off);
  void reg_conv3(const std::string& name, int O, int I, Conv3W& w);
  void reg_res(const std::string& name, int cin, int cout, ResW& r);
  int res_fwd(const ResW& w, const float* x, const float2* cs_x, int S, int H, int W, float** out, float2** cs_out);
  int run(const float* img, float* moments, int B, int H, int W);

  VaeCfg cfg_;
  std::map<std::string, Slot> slots_;
  std::vector<void*> owned_;
  float *cin_w_ = 0, *cin_b_ = 0, *no_g_ = 0, *no_b_ = 0, *qw_ = 0, *qb_ = 0, *stage_ = 0, *a_gng_ = 0, *a_gnb_ = 0, *a_bqkv_ = 0, *a_bo_ = 0;
  op_t *a_wqkv_ = 0, *a_wo_ = 0;
  std::vector<ResW> down_[4];
  Conv3W down_conv_[3], conv_out_;
  ResW mid_[2];
};

}  // namespace hedit
