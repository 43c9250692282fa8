// Pixel-space DDPM UNet of the face-swapping path (forward only); see face.cu.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "netexec.h"

namespace hedit {

struct FaceCfg {
  int ch = 128, nlevels = 6, mult[8] = {1, 1, 2, 2, 4, 4, 0, 0};
  int nres = 2, attn_res = 16, resolution = 256, in_ch = 3, out_ch = 3;
};

class FaceUNet : public NetExec {
 public:
  explicit FaceUNet(const FaceCfg& cfg);
  ~FaceUNet();
  bool ok() const { return err_.empty(); }
  const std::string& error() const { return err_; }
  int load_tensor(const char* name, const float* src, const int64_t* dims, int ndim, cudaStream_t st);
  int finalize(std::string* missing);
  int tensor_count() const { return int(slots_.size()); }
  bool tensor_info(int i, std::string& name, std::vector<int64_t>& shape) const;
  // x [S][in_ch][R][R] fp32 NCHW (device), t [S] (host) -> eps [S][out_ch][R][R] (device)
  int forward(const float* x, const float* t_host, float* eps, int S, cudaStream_t st);
  const FaceCfg& cfg() const { return cfg_; }

 private:
  struct Slot {
    enum Kind { F32, CONV_FWD, ROWS, CONV_UP_PHASES };
    struct Dst { Kind kind; void* dst; int ld; int off; };
    std::vector<int64_t> shape;
    std::vector<Dst> dsts;
    bool loaded = false;
  };
  struct Conv3W { op_t* w = nullptr; float* b = nullptr; int O = 0, I = 0; };
  struct ResW { int cin = 0, cout = 0, temb_off = 0; float *n1g = 0, *n1b = 0, *n2g = 0, *n2b = 0, *bsc = 0; Conv3W c1, c2; op_t* wsc = 0; };
  struct AttnW { int C = 0; float *gng = 0, *gnb = 0, *b_qkv = 0, *b_o = 0; op_t *w_qkv = 0, *w_o = 0; };
  struct Level { std::vector<ResW> res; std::vector<AttnW> attn; Conv3W resample; op_t* up_phases = nullptr; bool has_resample = false; };
  struct Act { float* x; float2* cs; int C; };

  template <typename T> T* walloc(size_t n);
  void reg(const std::string& name, std::vector<int64_t> shape, std::vector<Slot::Dst> dsts);
  void reg_conv3(const std::string& name, int O, int I, Conv3W& w);
  void reg_res(const std::string& name, int cin, int cout, ResW& r);
  void reg_attn(const std::string& name, int C, AttnW& a);
  int res_fwd(const ResW& w, const Act& a, const Act* skip, int S, int H, int W, Act* out);
  int attn_fwd(const AttnW& w, const Act& a, int S, int N, Act* out);
  int run(const float* x, float* eps, int S);

  FaceCfg cfg_;
  std::map<std::string, Slot> slots_;
  std::vector<void*> owned_;
  float *t_w1_ = 0, *t_b1_ = 0, *t_w2_ = 0, *t_b2_ = 0, *tproj_w_ = 0, *tproj_b_ = 0, *cin_w_ = 0, *cin_b_ = 0, *no_g_ = 0, *no_b_ = 0, *stage_ = 0;
  Conv3W conv_out_;
  int tproj_total_ = 0;
  std::vector<Level> down_, up_;
  ResW mid1_, mid2_; AttnW mid_attn_;
  // per-call time-embedding buffers (sized for max_S_)
  int max_S_ = 0;
  float *ts_dev_ = 0, *temb_a_ = 0, *temb_b_ = 0, *temb_rows_ = 0;
};

}  // namespace hedit
