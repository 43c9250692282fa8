// VAE decoder engine (forward + input-gradient backward); see vae.cu.
#pragma once
#include <cuda_runtime.h>
#include <map>
#include <string>
#include <vector>

#include "engine.h"
#include "netexec.h"

namespace hedit {

struct VaeCfg {
  int latent_ch = 4, out_ch = 3;
  int boc[4] = {128, 256, 512, 512};
  int layers = 2, groups = 32;
};

class VaeDecoder : public NetExec {
 public:
  explicit VaeDecoder(const VaeCfg& cfg);
  ~VaeDecoder();
  bool ok() const { return err_.empty(); }
  const std::string& error() const { return err_; }
  int load_tensor(const char* name, const float* src, const int64_t* dims, int ndim, cudaStream_t st);
  int finalize(std::string* missing);
  int tensor_count() const { return int(slots_.size()); }
  bool tensor_info(int i, std::string& name, std::vector<int64_t>& shape) const;
  // z [B][latent_ch][h][w] fp32 NCHW (device; already divided by the scaling factor) -> img [B][out_ch][8h][8w] fp32 NCHW (device).
  // Keeps what backward() needs until the next decode().
  int decode(const float* z, float* img, int B, int h, int w, cudaStream_t st);
  // dimg [B][out_ch][8h][8w] -> dz [B][latent_ch][h][w] for the last decode()
  int backward(const float* dimg, float* dz, cudaStream_t st);

 private:
  struct Slot {
    enum Kind { F32, CONV_FWD, CONV_DGRAD, ROWS, ROWS_T, CONVOUT_DGRAD, CONV_UP_PHASES };
    struct Dst { Kind kind; void* dst; int ld; int off; };
    std::vector<int64_t> shape;
    std::vector<Dst> dsts;
    bool loaded = false;
  };
  struct Conv3W { int O = 0, I = 0; op_t* fwd = nullptr; op_t* dgrad = nullptr; float* bias = nullptr; };
  struct ResW {
    int cin = 0, cout = 0;
    float *n1g = 0, *n1b = 0, *n2g = 0, *n2b = 0, *bsc = 0;
    Conv3W c1, c2;
    op_t *wsc = 0, *wsc_t = 0;
  };
  struct AttnW { int C = 0; float *gng = 0, *gnb = 0, *b_qkv = 0, *b_o = 0; op_t *w_qkv = 0, *w_qkv_t = 0, *w_o = 0, *w_o_t = 0; };
  struct GNSave { const float* x; float2* stats; int S, HW, C; const float* gamma; const float* beta; int silu; };
  struct ResSave { GNSave n1, n2; int S = 0, H = 0, W = 0; };
  struct AttnSave { GNSave gn; const op_t* qkv = nullptr; const op_t* P = nullptr; int S = 0, N = 0; };

  template <typename T> T* walloc(size_t n);
  void reg(const std::string& name, std::vector<int64_t> shape, std::vector<Slot::Dst> dsts);
  void reg_conv3(const std::string& name, int O, int I, Conv3W& w);
  void reg_res(const std::string& name, int cin, int cout, ResW& r);
  int gn_fwd1(const float* x, const float2* cs, int S, int HW, int C, const float* g, const float* b, int silu, op_t* out, op_t* raw, float2** stats_out) {
    return gn_fwd(x, cs, C, nullptr, nullptr, 0, S, HW, g, b, 1e-6f, silu, out, raw, stats_out);
  }
  int gn_bwd(const float* g, const GNSave& sv, const float* add, float* dx, op_t* dx16);
  int res_fwd(const ResW& w, ResSave& sv, const float* x, const float2* cs_x, int S, int H, int W, float** out, float2** cs_out);
  int res_bwd(const ResW& w, const ResSave& sv, const float* dout, const op_t* dout16, float** dx, op_t** dx16);
  int attn_fwd(const float* x, const float2* cs_x, int S, int N, float** out, float2** cs_out);
  int attn_bwd(const float* dout, const op_t* dout16, float** dx, op_t** dx16);
  int run_forward(const float* z, float* img, int B, int h, int w);
  int run_backward(const float* dimg, float* dz);
  int ensure_arena(int B, int h, int w);

  VaeCfg cfg_;
  std::map<std::string, Slot> slots_;
  std::vector<void*> owned_;
  float *pq_w_ = 0, *pq_b_ = 0, *cin_w_ = 0, *cin_b_ = 0, *no_g_ = 0, *no_b_ = 0, *cout_b_ = 0, *cout_dgrad_ = 0, *stage_ = 0;
  op_t *cin_dgrad_ = 0, *cout_w_ = 0;
  ResW mid_[2];
  AttnW attn_;
  std::vector<ResW> up_res_[4];
  Conv3W up_conv_[3];
  op_t* up_phases_[3] = {nullptr, nullptr, nullptr};     // fused upsample conv: [4][C][2][2][C]
  // run state
  uint8_t* arena_saved_ = nullptr; size_t fwd_top_ = 0;
  bool have_tape_ = false;
  ResSave mid_sv_[2]; AttnSave attn_sv_; std::vector<ResSave> up_sv_[4]; GNSave out_sv_{};
  int outH_ = 0, outW_ = 0, outC_ = 0, tapeB_ = 0, lat_h_ = 0, lat_w_ = 0;
};

}  // namespace hedit
