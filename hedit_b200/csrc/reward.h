// Reward networks of the face-swapping path, forward + input gradient, native (see reward.cu):
//   ArcFaceNet : IDLoss.get_cosine_loss  (face-swapping/arcface/arcface_model.py:41-70; IR-SE50 backbone arcface/facial_recognition/
//                model_irse.py:9-56 + helpers.py:47-125)
//   LpipsNet   : LPIPS_Loss.get_lpips_loss (arcface_model.py:72-95; the `lpips` package, net = 'vgg', version 0.1: VGG16 features at
//                relu1_2 / 2_2 / 3_3 / 4_3 / 5_3, unit-normalised, squared difference, 1x1 `lin` heads, spatial mean, summed)
#pragma once
#include <map>
#include <string>
#include <vector>

#include "netexec.h"

namespace hedit {

// raw fp32 tensors by name (the reference modules' state_dict keys), kept on the device until finalize() builds the kernel layouts
class RewardWeights {
 public:
  ~RewardWeights();
  int put(const char* name, const float* src, const int64_t* dims, int ndim, cudaStream_t st, std::string& err);
  // device pointer of a tensor with exactly `numel` elements, or null (err set)
  const float* get(const std::string& name, size_t numel, std::string& err) const;
  bool has(const std::string& name) const { return t_.count(name) != 0; }
  void clear();

 private:
  struct T { float* p; size_t n; };
  std::map<std::string, T> t_;
};

class ArcFaceNet : public NetExec {
 public:
  ArcFaceNet();
  ~ArcFaceNet();
  const std::string& error() const { return err_; }
  int load_tensor(const char* name, const float* src, const int64_t* dims, int ndim, cudaStream_t st) { return raw_.put(name, src, dims, ndim, st, err_); }
  int finalize(cudaStream_t st);
  // img [B][3][256][256] fp32 NCHW (device) -> unit-norm embedding [B][512] (device)
  int features(const float* img, int B, float* feat_unit, cudaStream_t st);
  // the face whose identity is transferred (IDLoss.ref): img [1][3][256][256]
  int set_reference(const float* img, cudaStream_t st);
  // loss[b] = 1 - cos(ref, f(img[b])) (device [B] or null) and grad = d loss[b] / d img[b] (device [B][3][256][256])
  int loss_grad(const float* img, int B, float* loss, float* grad, cudaStream_t st);

 private:
  struct Unit {
    int cin, depth, stride;
    float *bn1_s, *bn1_b, *prelu, *b2, *se_w1, *se_w2, *bsc;
    op_t *w1, *w1_d, *w2, *w2_d, *wsc, *wsc_t;
  };
  struct Tape { const float *c1, *r, *gate, *h; int Pin, Vin, Pout, Vout; };
  template <typename T> T* walloc(size_t n);
  int run_forward(const float* img, int B);
  int run_backward(float* grad, int B);
  int ensure_arena(int B, bool backward);

  RewardWeights raw_;
  std::vector<void*> owned_;
  std::vector<Unit> units_;
  std::vector<Tape> tape_;
  float *w0_ = nullptr, *b0_ = nullptr, *prelu0_ = nullptr, *bh_ = nullptr, *ref_hat_ = nullptr;
  op_t* wh_ = nullptr;
  bool ready_ = false, have_ref_ = false;
  // run state
  const float *z0_ = nullptr, *xlast_ = nullptr;
  float *feat_ = nullptr, *df_ = nullptr;
};

class LpipsNet : public NetExec {
 public:
  LpipsNet();
  ~LpipsNet();
  const std::string& error() const { return err_; }
  int load_tensor(const char* name, const float* src, const int64_t* dims, int ndim, cudaStream_t st) { return raw_.put(name, src, dims, ndim, st, err_); }
  int finalize(cudaStream_t st);
  // anchor image(s) (LPIPS_Loss.src): img [n][3][R][R], n = 1 (shared by the whole batch) or the batch size
  int set_source(const float* img, int n, int R, cudaStream_t st);
  // loss[b] = LPIPS(img[b], src[b or 0]) (device [B] or null), grad = d loss[b] / d img[b]
  int loss_grad(const float* img, int B, float* loss, float* grad, cudaStream_t st);

 private:
  struct Conv { int cin, cout; op_t *w, *w_d; float* b; };
  template <typename T> T* walloc(size_t n);
  int run(const float* img, int B, int mode, float* loss, float* grad);

  RewardWeights raw_;
  std::vector<void*> owned_;
  Conv conv_[13];
  float *w0_ = nullptr, *lin_[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  float shift_[3] = {0, 0, 0}, scale_[3] = {1, 1, 1};
  float* nref_[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  float cda_[13] = {0}, tstat_[13] = {0};     // static gradient-normalisation factors (finalize)
  int nsrc_ = 0, R_ = 0;
  bool ready_ = false;
};

}  // namespace hedit
