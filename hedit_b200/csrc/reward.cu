// Native reward networks of the face-swapping sampler h_Edit_R (face-swapping/inversion/h_edit_R.py:104-133): the reference evaluates
// `idloss.get_cosine_loss(x0_hat)` and `lpipsloss.get_lpips_loss(x0_hat)` and differentiates each back to the image 300 times per edit
// (100 steps x K = 3).  Both networks are frozen, so only the INPUT gradient is needed: every conv's backward is the same implicit-GEMM
// conv over the output gradient with transposed, tap-flipped weights prepared at load time (as in vae.cu); nothing touches autograd.
//
//   ArcFaceNet: crop [35:223, 32:220] -> adaptive average pool 112 x 112 -> IR-SE50 (3x3 stem, 24 bottleneck_IR_SE units at 64 / 128 /
//     256 / 512 channels, BatchNorm + Linear(25088, 512) + BatchNorm embedding) -> 1 - cosine similarity with the reference face.
//     BatchNorms that FOLLOW a conv / linear are folded into its weights at load; the BatchNorm that PRECEDES each unit's first conv is
//     applied by the pointwise kernel that writes the conv's 16-bit operand (folding it would change the zero padding).
//   LpipsNet: ScalingLayer -> VGG16 conv stack (13 convs, 4 max-pools) -> 5 feature taps -> LPIPS distance to the source image.
#include "reward.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "elementwise.cuh"
#include "nvtx.h"
#include "reward.cuh"
#include "tmap.h"
#include "vae.cuh"

namespace hedit {

#define RCK(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      char buf_[512];                                                                              \
      snprintf(buf_, sizeof buf_, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      err_ = buf_;                                                                                 \
      return -1;                                                                                   \
    }                                                                                              \
  } while (0)

// ------------------------------------------------------------------------------------------------ raw weights
RewardWeights::~RewardWeights() { clear(); }
void RewardWeights::clear() {
  for (auto& kv : t_) cudaFree(kv.second.p);
  t_.clear();
}
int RewardWeights::put(const char* name, const float* src, const int64_t* dims, int ndim, cudaStream_t st, std::string& err) {
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) n *= size_t(dims[i]);
  if (!name || !src || n == 0) { err = "bad tensor"; return -1; }
  auto it = t_.find(name);
  if (it != t_.end()) { cudaFree(it->second.p); t_.erase(it); }
  float* p = nullptr;
  if (cudaMalloc(&p, n * sizeof(float)) != cudaSuccess) { err = "cudaMalloc failed"; return -1; }
  if (cudaMemcpyAsync(p, src, n * sizeof(float), cudaMemcpyDefault, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
    cudaFree(p); err = std::string("copy failed for ") + name; return -1;
  }
  t_[name] = T{p, n};
  return 0;
}
const float* RewardWeights::get(const std::string& name, size_t numel, std::string& err) const {
  auto it = t_.find(name);
  if (it == t_.end()) { err = "missing weight " + name; return nullptr; }
  if (it->second.n != numel) { err = "shape mismatch for " + name; return nullptr; }
  return it->second.p;
}

static inline dim3 pw_grid(int P, int C, int B) { return dim3((P * P * (C / 4) + 255) / 256, B); }

// ------------------------------------------------------------------------------------------------ ArcFace IR-SE50
ArcFaceNet::ArcFaceNet() {
  const int blocks[4][3] = {{64, 64, 3}, {64, 128, 4}, {128, 256, 14}, {256, 512, 3}};      // helpers.py get_blocks(50)
  for (auto& b : blocks)
    for (int k = 0; k < b[2]; ++k) {
      Unit u{};
      u.cin = k == 0 ? b[0] : b[1]; u.depth = b[1]; u.stride = k == 0 ? 2 : 1;
      units_.push_back(u);
    }
}
ArcFaceNet::~ArcFaceNet() { for (void* p : owned_) cudaFree(p); }

template <typename T>
T* ArcFaceNet::walloc(size_t n) {
  void* p = nullptr;
  if (cudaMalloc(&p, std::max<size_t>(n, 4) * sizeof(T)) != cudaSuccess) { err_ = "cudaMalloc failed"; return nullptr; }
  owned_.push_back(p);
  return reinterpret_cast<T*>(p);
}

namespace {
// scratch copy of a raw weight, optionally scaled per output row
struct Scratch {
  float* p = nullptr;
  ~Scratch() { if (p) cudaFree(p); }
  bool make(const float* src, size_t n, const float* row_scale, size_t n_per_o, cudaStream_t st) {
    if (cudaMalloc(&p, n * sizeof(float)) != cudaSuccess) return false;
    cudaMemcpyAsync(p, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (row_scale) rw_scale_rows_kernel<<<int(std::min<size_t>((n + 255) / 256, 4096)), 256, 0, st>>>(p, row_scale, n_per_o, n);
    return true;
  }
};
}  // namespace

int ArcFaceNet::finalize(cudaStream_t st) {
  err_.clear();
  auto bn = [&](const std::string& P, int C, float** s, float** t) -> bool {
    const float* g = raw_.has(P + ".weight") ? raw_.get(P + ".weight", C, err_) : nullptr;
    const float* b = raw_.has(P + ".bias") ? raw_.get(P + ".bias", C, err_) : nullptr;
    const float* m = raw_.get(P + ".running_mean", C, err_);
    const float* v = raw_.get(P + ".running_var", C, err_);
    if (!m || !v || !err_.empty()) return false;
    *s = walloc<float>(C); *t = walloc<float>(C);
    if (!*s || !*t) return false;
    rw_bn_affine_kernel<<<(C + 127) / 128, 128, 0, st>>>(g, b, m, v, 1e-5f, *s, *t, C);
    return true;
  };
  auto conv3w = [&](const float* w, int O, int I, const float* row_scale, op_t** fwd, op_t** dgr) -> bool {
    const size_t n = size_t(O) * I * 9;
    Scratch sc;
    if (!sc.make(w, n, row_scale, size_t(I) * 9, st)) { err_ = "cudaMalloc failed"; return false; }
    *fwd = walloc<op_t>(n); *dgr = walloc<op_t>(n);
    if (!*fwd || !*dgr) return false;
    const int blocks = int(std::min<size_t>((n + 255) / 256, 4096));
    vae_cvt_weight_kernel<<<blocks, 256, 0, st>>>(sc.p, *fwd, O, I, 0, 0, 0);
    vae_cvt_weight_kernel<<<blocks, 256, 0, st>>>(sc.p, *dgr, O, I, 1, 0, 0);
    return cudaStreamSynchronize(st) == cudaSuccess;
  };
  auto copyf = [&](const float* src, size_t n) -> float* {
    float* d = walloc<float>(n);
    if (d && src) cudaMemcpyAsync(d, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
    return d;
  };
  // stem: Conv2d(3, 64) + BatchNorm2d folded, PReLU
  {
    float *s, *t;
    if (!bn("input_layer.1", 64, &s, &t)) return -1;
    const float* w = raw_.get("input_layer.0.weight", 64 * 27, err_);
    const float* pr = raw_.get("input_layer.2.weight", 64, err_);
    if (!w || !pr) return -1;
    Scratch sc;
    if (!sc.make(w, 64 * 27, s, 27, st)) { err_ = "cudaMalloc failed"; return -1; }
    w0_ = walloc<float>(27 * 64);
    rw_cvt_c3_kernel<<<(64 * 27 + 255) / 256, 256, 0, st>>>(sc.p, w0_, 64);
    b0_ = t; prelu0_ = copyf(pr, 64);
    RCK(cudaStreamSynchronize(st));
  }
  for (size_t i = 0; i < units_.size(); ++i) {
    Unit& u = units_[i];
    const std::string P = "body." + std::to_string(i) + ".";
    const int ci = u.cin, d = u.depth;
    if (!bn(P + "res_layer.0", ci, &u.bn1_s, &u.bn1_b)) return -1;
    const float* w1 = raw_.get(P + "res_layer.1.weight", size_t(d) * ci * 9, err_);
    const float* pr = raw_.get(P + "res_layer.2.weight", d, err_);
    const float* w2 = raw_.get(P + "res_layer.3.weight", size_t(d) * d * 9, err_);
    const float* f1 = raw_.get(P + "res_layer.5.fc1.weight", size_t(d / 16) * d, err_);
    const float* f2 = raw_.get(P + "res_layer.5.fc2.weight", size_t(d) * (d / 16), err_);
    if (!w1 || !pr || !w2 || !f1 || !f2) return -1;
    float* s2;
    if (!bn(P + "res_layer.4", d, &s2, &u.b2)) return -1;
    if (!conv3w(w1, d, ci, nullptr, &u.w1, &u.w1_d) || !conv3w(w2, d, d, s2, &u.w2, &u.w2_d)) return -1;
    u.prelu = copyf(pr, d); u.se_w1 = copyf(f1, size_t(d / 16) * d); u.se_w2 = copyf(f2, size_t(d) * (d / 16));
    if (ci != d) {
      if (u.stride != 2) { err_ = "unsupported IR-SE geometry (projection shortcut with stride 1)"; return -1; }
      const float* ws = raw_.get(P + "shortcut_layer.0.weight", size_t(d) * ci, err_);
      float* ss;
      if (!ws || !bn(P + "shortcut_layer.1", d, &ss, &u.bsc)) return -1;
      Scratch sc;
      if (!sc.make(ws, size_t(d) * ci, ss, ci, st)) { err_ = "cudaMalloc failed"; return -1; }
      u.wsc = walloc<op_t>(size_t(d) * ci); u.wsc_t = walloc<op_t>(size_t(d) * ci);
      const int blocks = int((size_t(d) * ci + 255) / 256);
      vae_cvt_weight_kernel<<<blocks, 256, 0, st>>>(sc.p, u.wsc, d, ci, 2, ci, 0);
      vae_cvt_weight_kernel<<<blocks, 256, 0, st>>>(sc.p, u.wsc_t, d, ci, 3, d, 0);
      RCK(cudaStreamSynchronize(st));
    }
  }
  {
    float *s2, *t2, *s1, *t1;
    if (!bn("output_layer.0", 512, &s2, &t2) || !bn("output_layer.4", 512, &s1, &t1)) return -1;
    const float* wl = raw_.get("output_layer.3.weight", size_t(512) * 512 * 49, err_);
    const float* bl = raw_.get("output_layer.3.bias", 512, err_);
    if (!wl || !bl) return -1;
    wh_ = walloc<op_t>(size_t(512) * 49 * 512); bh_ = walloc<float>(512);
    rw_cvt_head_kernel<<<512, 256, 0, st>>>(wl, bl, s2, t2, s1, t1, wh_, bh_);
  }
  ref_hat_ = walloc<float>(512);
  RCK(cudaGetLastError());
  RCK(cudaStreamSynchronize(st));
  if (!err_.empty()) return -1;
  raw_.clear();
  ready_ = true;
  return 0;
}

int ArcFaceNet::run_forward(const float* img, int B) {
  constexpr int R = 256, N0 = 112, P0 = 128;
  float* pool = A<float>(size_t(B) * N0 * N0 * 3);
  float* z0 = A<float>(size_t(B) * P0 * P0 * 64);
  float* x = A<float>(size_t(B) * P0 * P0 * 64);
  op_t* a16 = A<op_t>(size_t(B) * P0 * P0 * 64);
  if (!dry_) {
    rw_crop_pool_fwd_kernel<<<dim3((N0 * N0 * 3 + 255) / 256, B), 256, 0, st_>>>(img, pool, R, 35, 32, 188, N0);
    rw_conv_c3_fwd_kernel<<<dim3(P0 * P0 / 16, B), 256, 0, st_>>>(pool, w0_, b0_, z0, N0, P0);
    RwActParams ap{z0, prelu0_, units_[0].bn1_s, units_[0].bn1_b, x, a16, P0, N0, 64};
    rw_act_kernel<<<pw_grid(P0, 64, B), 256, 0, st_>>>(ap);
    launches_ += 3;
  }
  z0_ = z0;
  tape_.assign(units_.size(), Tape{});
  int P = P0, V = N0;
  for (size_t i = 0; i < units_.size(); ++i) {
    const Unit& u = units_[i];
    const int ci = u.cin, d = u.depth, Po = P / u.stride, Vo = V / u.stride;
    GemmEpilogue e; memset(&e, 0, sizeof e);
    float* c1 = A<float>(size_t(B) * P * P * d);
    e.out_f32 = c1; e.ldo = d;
    if (conv3(a16, u.w1, B, P, P, ci, d, e)) return -1;
    op_t* p16 = A<op_t>(size_t(B) * P * P * d);
    if (!dry_) {
      RwActParams ap{c1, u.prelu, nullptr, nullptr, nullptr, p16, P, V, d};
      rw_act_kernel<<<pw_grid(P, d, B), 256, 0, st_>>>(ap);
      ++launches_;
    }
    float* r = A<float>(size_t(B) * Po * Po * d);
    memset(&e, 0, sizeof e); e.bias = u.b2; e.out_f32 = r; e.ldo = d;
    if (conv3(p16, u.w2, B, Po, Po, d, d, e, u.stride)) return -1;
    const int nch = (Vo + RW_POOL_ROWS - 1) / RW_POOL_ROWS;
    float* partial = A<float>(size_t(B) * nch * d);
    float* h = A<float>(size_t(B) * 32);
    float* gate = A<float>(size_t(B) * d);
    if (!dry_) {
      rw_pool_partial_kernel<<<dim3(nch, B), 256, 512 * sizeof(float4), st_>>>(r, nullptr, partial, nullptr, Po, Vo, d);
      rw_se_fc_kernel<<<B, 512, 0, st_>>>(partial, nch, u.se_w1, u.se_w2, h, gate, d, 1.f / float(Vo * Vo));
      launches_ += 2;
    }
    const float* sc = x; int sc_stride = u.stride;
    if (u.wsc) {
      op_t* xs16 = A<op_t>(size_t(B) * Po * Po * ci);
      if (!dry_) { rw_gather2_kernel<<<pw_grid(Po, ci, B), 256, 0, st_>>>(x, xs16, Po, ci); ++launches_; }
      float* scb = A<float>(size_t(B) * Po * Po * d);
      memset(&e, 0, sizeof e); e.bias = u.bsc; e.out_f32 = scb; e.ldo = d;
      if (gemm(xs16, ci, A_LINEAR, nullptr, u.wsc, B * Po * Po, d, ci, e)) return -1;
      sc = scb; sc_stride = 1;
    }
    float* out = A<float>(size_t(B) * Po * Po * d);
    const bool last = i + 1 == units_.size();
    op_t* a16n = last ? nullptr : A<op_t>(size_t(B) * Po * Po * d);
    if (!dry_) {
      RwCombineParams cp{r, gate, sc, sc_stride, last ? nullptr : units_[i + 1].bn1_s, last ? nullptr : units_[i + 1].bn1_b, out, a16n, Po, Vo, d};
      rw_combine_kernel<<<pw_grid(Po, d, B), 256, 0, st_>>>(cp);
      ++launches_;
    }
    tape_[i] = Tape{c1, r, gate, h, P, V, Po, Vo};
    x = out; a16 = a16n; P = Po; V = Vo;
  }
  xlast_ = x;
  feat_ = A<float>(size_t(B) * 512);
  if (!dry_) {
    for (int b0 = 0; b0 < B; b0 += RW_HEAD_MAXB) {
      rw_head_fwd_kernel<<<256, 256, 0, st_>>>(x + size_t(b0) * 64 * 512, wh_, bh_, feat_ + size_t(b0) * 512, std::min(RW_HEAD_MAXB, B - b0));
      ++launches_;
    }
    RCK(cudaGetLastError());
  }
  return 0;
}

int ArcFaceNet::run_backward(float* grad, int B) {
  // df_ holds d loss / d feature
  float* dout = A<float>(size_t(B) * 64 * 512);
  if (!dry_) {
    rw_head_bwd_kernel<<<dim3(64, (B + RW_HEAD_MAXB - 1) / RW_HEAD_MAXB), 256, 0, st_>>>(df_, wh_, dout, B);
    ++launches_;
  }
  for (int i = int(units_.size()) - 1; i >= 0; --i) {
    const Unit& u = units_[i];
    const Tape& tp = tape_[i];
    const int ci = u.cin, d = u.depth, P = tp.Pin, V = tp.Vin, Po = tp.Pout, Vo = tp.Vout;
    const int nch = (Vo + RW_POOL_ROWS - 1) / RW_POOL_ROWS;
    float* partial = A<float>(size_t(B) * nch * d);
    float* sqpart = A<float>(size_t(B) * nch * d);
    float* dmean = A<float>(size_t(B) * d);
    float* gscale = A<float>(B);                          // 1 / rms(dout[b]): normalisation of the 16-bit gradient operands of this unit
    const int up = u.stride == 2 ? 1 : 0;
    op_t* dr16 = A<op_t>(size_t(B) * P * P * d);          // (P = 2 Po when up)
    if (!dry_) {
      rw_pool_partial_kernel<<<dim3(nch, B), 256, 512 * sizeof(float4), st_>>>(dout, tp.r, partial, sqpart, Po, Vo, d);
      rw_se_bwd_kernel<<<B, 512, 0, st_>>>(partial, sqpart, nch, u.se_w1, u.se_w2, tp.h, tp.gate, dmean, gscale, d, 1.f / float(Vo * Vo));
      rw_dr_kernel<<<pw_grid(P, d, B), 256, 0, st_>>>(dout, tp.gate, dmean, gscale, dr16, Po, Vo, d, up);
      launches_ += 3;
    }
    GemmEpilogue e; memset(&e, 0, sizeof e);
    float* dp = A<float>(size_t(B) * P * P * d);
    e.out_f32 = dp; e.ldo = d;
    if (conv3(dr16, u.w2_d, B, P, P, d, d, e)) return -1;
    op_t* dc16 = A<op_t>(size_t(B) * P * P * d);
    if (!dry_) { rw_prelu_bwd_kernel<<<pw_grid(P, d, B), 256, 0, st_>>>(dp, tp.c1, u.prelu, dc16, nullptr, P, V, d); ++launches_; }
    float* da = A<float>(size_t(B) * P * P * ci);
    memset(&e, 0, sizeof e); e.out_f32 = da; e.ldo = ci;
    if (conv3(dc16, u.w1_d, B, P, P, d, ci, e)) return -1;
    const float* sg = dout; int sg_stride = u.stride;
    if (u.wsc) {
      op_t* do16 = A<op_t>(size_t(B) * Po * Po * d);
      if (!dry_) { rw_cast_masked_kernel<<<pw_grid(Po, d, B), 256, 0, st_>>>(dout, gscale, do16, Po, Vo, d); ++launches_; }
      float* sgb = A<float>(size_t(B) * Po * Po * ci);
      memset(&e, 0, sizeof e); e.out_f32 = sgb; e.ldo = ci;
      if (gemm(do16, d, A_LINEAR, nullptr, u.wsc_t, B * Po * Po, ci, d, e)) return -1;
      sg = sgb; sg_stride = 2;
    }
    float* dx = A<float>(size_t(B) * P * P * ci);
    if (!dry_) {
      RwUnitInBwdParams q{da, u.bn1_s, gscale, u.wsc ? 1 : 0, sg, sg_stride, dx, nullptr, P, V, ci};
      rw_unit_in_bwd_kernel<<<pw_grid(P, ci, B), 256, 0, st_>>>(q);
      ++launches_;
    }
    dout = dx;
  }
  constexpr int R = 256, N0 = 112, P0 = 128;
  float* dz0 = A<float>(size_t(B) * P0 * P0 * 64);
  float* dpool = A<float>(size_t(B) * N0 * N0 * 3);
  if (!dry_) {
    rw_prelu_bwd_kernel<<<pw_grid(P0, 64, B), 256, 0, st_>>>(dout, z0_, prelu0_, nullptr, dz0, P0, N0, 64);
    rw_conv_c3_bwd_kernel<<<dim3((N0 * N0 + 7) / 8, B), 256, 0, st_>>>(dz0, w0_, dpool, N0, P0);
    rw_crop_pool_bwd_kernel<<<dim3((3 * R * R + 255) / 256, B), 256, 0, st_>>>(dpool, grad, R, 35, 32, 188, N0);
    launches_ += 3;
    RCK(cudaGetLastError());
  }
  return 0;
}

int ArcFaceNet::ensure_arena(int B, bool backward) {
  uint8_t* saved = arena_;
  dry_ = true; top_ = 0; peak_ = 0; arena_ = nullptr;
  int r = run_forward(nullptr, B);
  if (!r) { df_ = A<float>(size_t(B) * 512); if (backward) r = run_backward(nullptr, B); }
  dry_ = false; arena_ = saved;
  if (r) return -1;
  return reserve(peak_ + (size_t(1) << 20), "ArcFace");
}

int ArcFaceNet::features(const float* img, int B, float* feat_unit, cudaStream_t st) {
  if (!ready_) { err_ = "ArcFace weights not finalized"; return -1; }
  if (B < 1) { err_ = "empty batch"; return -1; }
  if (ensure_arena(B, false)) return -1;
  st_ = st; top_ = 0; launches_ = 0; flops_ = 0;
  if (run_forward(img, B)) return -1;
  rw_cos_loss_kernel<<<B, 128, 0, st_>>>(feat_, nullptr, nullptr, nullptr, feat_unit, 1);
  ++launches_;
  RCK(cudaGetLastError());
  return 0;
}

int ArcFaceNet::set_reference(const float* img, cudaStream_t st) {
  if (features(img, 1, ref_hat_, st)) return -1;
  RCK(cudaStreamSynchronize(st));
  have_ref_ = true;
  drop_graphs();
  return 0;
}

int ArcFaceNet::loss_grad(const float* img, int B, float* loss, float* grad, cudaStream_t st) {
  NvtxRange nvtx_("hedit.arcface.loss_grad B=%d%.0d", B, 0);
  if (!ready_ || !have_ref_) { err_ = "ArcFace: finalize() and set_reference() first"; return -1; }
  if (B < 1 || !img || !grad) { err_ = "bad arguments"; return -1; }
  if (ensure_arena(B, true)) return -1;
  const GraphKey key{reinterpret_cast<uintptr_t>(img), reinterpret_cast<uintptr_t>(grad), reinterpret_cast<uintptr_t>(loss), uintptr_t(B)};
  return run_or_capture(key, st, [&]() -> int {
    top_ = 0; launches_ = 0; flops_ = 0;
    if (run_forward(img, B)) return -1;
    df_ = A<float>(size_t(B) * 512);
    rw_cos_loss_kernel<<<B, 128, 0, st_>>>(feat_, ref_hat_, loss, df_, nullptr, 0);
    ++launches_;
    if (run_backward(grad, B)) return -1;
    RCK(cudaGetLastError());
    return 0;
  });
}

// ------------------------------------------------------------------------------------------------ LPIPS-VGG16
static const int kVggCout[13] = {64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512, 512};
static const int kVggTap[5] = {1, 3, 6, 9, 12};              // relu1_2, relu2_2, relu3_3, relu4_3, relu5_3
static inline int vgg_pool_after(int i) { return (i == 1 || i == 3 || i == 6 || i == 9) ? 2 : 1; }
static inline int vgg_tap_of(int i) { for (int k = 0; k < 5; ++k) if (kVggTap[k] == i) return k; return -1; }

LpipsNet::LpipsNet() {
  int cin = 3;
  for (int i = 0; i < 13; ++i) { conv_[i] = Conv{cin, kVggCout[i], nullptr, nullptr, nullptr}; cin = kVggCout[i]; }
}
LpipsNet::~LpipsNet() {
  for (void* p : owned_) cudaFree(p);
  for (float* p : nref_) if (p) cudaFree(p);
}
template <typename T>
T* LpipsNet::walloc(size_t n) {
  void* p = nullptr;
  if (cudaMalloc(&p, std::max<size_t>(n, 4) * sizeof(T)) != cudaSuccess) { err_ = "cudaMalloc failed"; return nullptr; }
  owned_.push_back(p);
  return reinterpret_cast<T*>(p);
}

int LpipsNet::finalize(cudaStream_t st) {
  err_.clear();
  for (int i = 0; i < 13; ++i) {
    Conv& c = conv_[i];
    const std::string P = "conv" + std::to_string(i);
    const size_t n = size_t(c.cout) * c.cin * 9;
    const float* w = raw_.get(P + ".weight", n, err_);
    const float* b = raw_.get(P + ".bias", c.cout, err_);
    if (!w || !b) return -1;
    c.b = walloc<float>(c.cout);
    RCK(cudaMemcpyAsync(c.b, b, c.cout * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (i == 0) {
      w0_ = walloc<float>(27 * 64);
      rw_cvt_c3_kernel<<<(64 * 27 + 255) / 256, 256, 0, st>>>(w, w0_, 64);
    } else {
      c.w = walloc<op_t>(n); c.w_d = walloc<op_t>(n);
      const int blocks = int(std::min<size_t>((n + 255) / 256, 4096));
      vae_cvt_weight_kernel<<<blocks, 256, 0, st>>>(w, c.w, c.cout, c.cin, 0, 0, 0);
      vae_cvt_weight_kernel<<<blocks, 256, 0, st>>>(w, c.w_d, c.cout, c.cin, 1, 0, 0);
    }
  }
  // static estimate of each dgrad conv's rms gain (sum w^2 / cin over random-sign inputs), times ~0.7 for the ReLU mask and 0.5 for a
  // max-pool's routing: the backward multiplies by the inverse so its 16-bit gradient operands stay near unit rms (reward.cuh)
  {
    float* d_ss = walloc<float>(16);
    float h_ss[13] = {0};
    for (int i = 1; i < 13; ++i) {
      const Conv& c = conv_[i];
      rw_sumsq_kernel<<<1, 1024, 0, st>>>(raw_.get("conv" + std::to_string(i) + ".weight", size_t(c.cout) * c.cin * 9, err_), size_t(c.cout) * c.cin * 9, d_ss + i);
    }
    RCK(cudaMemcpyAsync(h_ss, d_ss, sizeof h_ss, cudaMemcpyDeviceToHost, st));
    RCK(cudaStreamSynchronize(st));
    for (int i = 1; i < 13; ++i) {
      const float gain = std::sqrt(std::max(h_ss[i], 1e-30f) / float(conv_[i].cin)) * 0.7f * (vgg_pool_after(i - 1) == 2 ? 0.5f : 1.f);
      cda_[i] = 1.f / std::max(gain, 1e-6f);
    }
    tstat_[12] = 1.f;
    for (int i = 12; i >= 1; --i) tstat_[i - 1] = tstat_[i] * cda_[i];
  }
  for (int k = 0; k < 5; ++k) {
    const int C = kVggCout[kVggTap[k]];
    const float* l = raw_.get("lin" + std::to_string(k) + ".weight", C, err_);
    if (!l) return -1;
    lin_[k] = walloc<float>(C);
    RCK(cudaMemcpyAsync(lin_[k], l, C * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  const float* sh = raw_.get("shift", 3, err_);
  const float* sc = raw_.get("scale", 3, err_);
  if (!sh || !sc) return -1;
  RCK(cudaMemcpyAsync(shift_, sh, 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
  RCK(cudaMemcpyAsync(scale_, sc, 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
  RCK(cudaGetLastError());
  RCK(cudaStreamSynchronize(st));
  raw_.clear();
  ready_ = true;
  return 0;
}

template <int C>
static void launch_tap(const float* z, float* nref, const float* ref, int ref_bstride, const float* lin, float* gtap, float* partial, float* gsq,
                       int HW, int B, int mode, cudaStream_t st) {
  rw_lpips_tap_kernel<C><<<dim3((HW + 7) / 8, B), 256, 0, st>>>(z, nref, ref, ref_bstride, lin, gtap, partial, gsq, HW, mode);
}

// mode 1: forward only, unit-normalised tap features written to nref_ (source set-up).  mode 0: loss + input gradient.
int LpipsNet::run(const float* img, int B, int mode, float* loss, float* grad) {
  const int R = R_;
  const float3 shift = make_float3(shift_[0], shift_[1], shift_[2]);
  const float3 inv_scale = make_float3(1.f / scale_[0], 1.f / scale_[1], 1.f / scale_[2]);
  float* xs = A<float>(size_t(B) * R * R * 3);
  float* z[13]; int Hs[13];
  float* gtap[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  int nblk_total = 0, blk_off[5];
  for (int k = 0, H = R; k < 5; ++k) { blk_off[k] = nblk_total; nblk_total += (H * H + 7) / 8; H /= 2; }
  float* partial = A<float>(size_t(B) * nblk_total);
  float* gsq = A<float>(size_t(B) * nblk_total);
  float* gs = A<float>(B);
  z[0] = A<float>(size_t(B) * R * R * 64); Hs[0] = R;
  if (!dry_) {
    rw_vgg_prep_kernel<<<dim3((R * R + 255) / 256, B), 256, 0, st_>>>(img, xs, R, shift, inv_scale);
    rw_conv_c3_fwd_kernel<<<dim3(R * R / 16, B), 256, 0, st_>>>(xs, w0_, conv_[0].b, z[0], R, R);
    launches_ += 2;
  }
  int H = R;
  for (int i = 1; i < 13; ++i) {
    const Conv& c = conv_[i];
    const int pool = vgg_pool_after(i - 1), Ho = H / pool;
    op_t* a16 = A<op_t>(size_t(B) * Ho * Ho * c.cin);
    if (!dry_) { rw_vgg_act_kernel<<<pw_grid(Ho, c.cin, B), 256, 0, st_>>>(z[i - 1], a16, H, H, c.cin, pool); ++launches_; }
    H = Ho;
    z[i] = A<float>(size_t(B) * H * H * c.cout); Hs[i] = H;
    GemmEpilogue e; memset(&e, 0, sizeof e);
    e.bias = c.b; e.out_f32 = z[i]; e.ldo = c.cout;
    if (conv3(a16, c.w, B, H, H, c.cin, c.cout, e)) return -1;
  }
  for (int k = 0; k < 5; ++k) {
    const int i = kVggTap[k], C = kVggCout[i], HW = Hs[i] * Hs[i];
    if (mode == 0) gtap[k] = A<float>(size_t(B) * HW * C);
    if (dry_) continue;
    const int rbs = nsrc_ == 1 ? 0 : HW * C;
    // loss partials: tap k owns [B][blocks_k] at offset blk_off[k] * B
    float* pk = partial + size_t(blk_off[k]) * B;
    float* gk = gsq + size_t(blk_off[k]) * B;
    switch (C) {
      case 64: launch_tap<64>(z[i], nref_[k], nref_[k], rbs, lin_[k], gtap[k], pk, gk, HW, B, mode, st_); break;
      case 128: launch_tap<128>(z[i], nref_[k], nref_[k], rbs, lin_[k], gtap[k], pk, gk, HW, B, mode, st_); break;
      case 256: launch_tap<256>(z[i], nref_[k], nref_[k], rbs, lin_[k], gtap[k], pk, gk, HW, B, mode, st_); break;
      default: launch_tap<512>(z[i], nref_[k], nref_[k], rbs, lin_[k], gtap[k], pk, gk, HW, B, mode, st_); break;
    }
    ++launches_;
  }
  if (mode == 1) { if (!dry_) RCK(cudaGetLastError()); return 0; }
  if (!dry_) {
    RwLpipsScaleParams sp;
    for (int k = 0; k < 5; ++k) {
      const int i = kVggTap[k], HW = Hs[i] * Hs[i];
      sp.nblk[k] = (HW + 7) / 8; sp.boff[k] = blk_off[k]; sp.tstat[k] = tstat_[i]; sp.inv_count[k] = 1.f / (float(HW) * float(kVggCout[i]));
    }
    rw_lpips_scale_kernel<<<B, 256, 0, st_>>>(partial, gsq, sp, B, loss, gs);
    ++launches_;
  }
  // backward
  const float* da = nullptr;
  for (int i = 12; i >= 1; --i) {
    const Conv& c = conv_[i];
    const int Hi = Hs[i], pool = i == 12 ? 1 : vgg_pool_after(i), k = vgg_tap_of(i);
    op_t* dz16 = A<op_t>(size_t(B) * Hi * Hi * c.cout);
    if (!dry_) {
      const float cda = i == 12 ? 0.f : cda_[i + 1];
      if (pool == 2) rw_vgg_bwd_kernel<2><<<pw_grid(Hi / 2, c.cout, B), 256, 0, st_>>>(da, z[i], k >= 0 ? gtap[k] : nullptr, dz16, nullptr, Hi, Hi, c.cout, cda, gs, tstat_[i]);
      else rw_vgg_bwd_kernel<1><<<pw_grid(Hi, c.cout, B), 256, 0, st_>>>(da, z[i], k >= 0 ? gtap[k] : nullptr, dz16, nullptr, Hi, Hi, c.cout, cda, gs, tstat_[i]);
      ++launches_;
    }
    float* dan = A<float>(size_t(B) * Hi * Hi * c.cin);
    GemmEpilogue e; memset(&e, 0, sizeof e);
    e.out_f32 = dan; e.ldo = c.cin;
    if (conv3(dz16, c.w_d, B, Hi, Hi, c.cout, c.cin, e)) return -1;
    da = dan;
  }
  float* dz0 = A<float>(size_t(B) * R * R * 64);
  float* dxs = A<float>(size_t(B) * R * R * 3);
  if (!dry_) {
    rw_vgg_bwd_kernel<1><<<pw_grid(R, 64, B), 256, 0, st_>>>(da, z[0], nullptr, nullptr, dz0, R, R, 64, cda_[1], gs, 0.f);
    rw_conv_c3_bwd_kernel<<<dim3((R * R + 7) / 8, B), 256, 0, st_>>>(dz0, w0_, dxs, R, R);
    rw_vgg_unprep_kernel<<<dim3((R * R + 255) / 256, B), 256, 0, st_>>>(dxs, grad, R, inv_scale, gs, tstat_[0]);
    launches_ += 3;
    RCK(cudaGetLastError());
  }
  return 0;
}

int LpipsNet::set_source(const float* img, int n, int R, cudaStream_t st) {
  if (!ready_) { err_ = "LPIPS weights not finalized"; return -1; }
  if (n < 1 || (R != 128 && R != 256 && R != 512)) { err_ = "LPIPS: image side must be 128, 256 or 512"; return -1; }
  drop_graphs();
  for (int k = 0, H = R; k < 5; ++k, H /= 2) {
    if (nref_[k]) { cudaFree(nref_[k]); nref_[k] = nullptr; }
    RCK(cudaMalloc(&nref_[k], size_t(n) * H * H * kVggCout[kVggTap[k]] * sizeof(float)));
  }
  nsrc_ = n; R_ = R;
  uint8_t* saved = arena_;
  dry_ = true; top_ = 0; peak_ = 0; arena_ = nullptr;
  int r = run(nullptr, n, 1, nullptr, nullptr);
  dry_ = false; arena_ = saved;
  if (r || reserve(peak_ + (size_t(1) << 20), "LPIPS")) return -1;
  st_ = st; top_ = 0; launches_ = 0; flops_ = 0;
  if (run(img, n, 1, nullptr, nullptr)) return -1;
  RCK(cudaStreamSynchronize(st));
  return 0;
}

int LpipsNet::loss_grad(const float* img, int B, float* loss, float* grad, cudaStream_t st) {
  NvtxRange nvtx_("hedit.lpips.loss_grad B=%d%.0d", B, 0);
  if (!ready_ || !nsrc_) { err_ = "LPIPS: finalize() and set_source() first"; return -1; }
  if (B < 1 || !img || !grad || (nsrc_ != 1 && nsrc_ != B)) { err_ = "LPIPS: batch must match the number of source images (or use one source)"; return -1; }
  uint8_t* saved = arena_;
  dry_ = true; top_ = 0; peak_ = 0; arena_ = nullptr;
  int r = run(nullptr, B, 0, loss, nullptr);
  dry_ = false; arena_ = saved;
  if (r || reserve(peak_ + (size_t(1) << 20), "LPIPS")) return -1;
  const GraphKey key{reinterpret_cast<uintptr_t>(img), reinterpret_cast<uintptr_t>(grad), reinterpret_cast<uintptr_t>(loss), uintptr_t(B)};
  return run_or_capture(key, st, [&]() -> int {
    top_ = 0; launches_ = 0; flops_ = 0;
    return run(img, B, 0, loss, grad);
  });
}

}  // namespace hedit
