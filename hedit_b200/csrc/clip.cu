// CLIP-Gram style reward: loss(img) = || F^T F - F_ref^T F_ref ||_F with F = the patch-token features after the 3rd ViT block of the
// CLIP image tower evaluated on the bicubically resized, normalised image, and dLoss/dimg -- the image branch of the reference's style
// reward (text-guided-n-style/clip_guidance/base_clip.py:55-66 on clip/model.py:202-221,167-188,339-359), which the reference
// differentiates with torch.autograd inside its Langevin loop (text-guided-n-style/inversion/h_edit.py:161-164).
// Patch embedding and every linear layer (forward and dgrad, with transposed weights prepared at load) run on the tcgen05 GEMM; the
// 197-token attention, LayerNorm, QuickGELU, resize and Gram pieces are the fp32 kernels of clip.cuh.
#include "clip.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "clip.cuh"
#include "tmap.h"
#include "vae.cuh"

namespace hedit {

#define CCK(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      char buf_[512];                                                                              \
      snprintf(buf_, sizeof buf_, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      err_ = buf_;                                                                                 \
      return -1;                                                                                   \
    }                                                                                              \
  } while (0)

static constexpr float kGradScale = 1.0f / 64.0f;      // keeps 16-bit gradient operands well inside the fp16 range; undone in fp32

template <typename T>
T* ClipGram::walloc(size_t n) {
  void* p = nullptr;
  if (cudaMalloc(&p, std::max<size_t>(n, 4) * sizeof(T)) != cudaSuccess) { err_ = "cudaMalloc failed"; return nullptr; }
  cudaMemset(p, 0, std::max<size_t>(n, 4) * sizeof(T));
  owned_.push_back(p);
  return reinterpret_cast<T*>(p);
}
template <typename T>
T* ClipGram::A(size_t n) {
  const size_t bytes = (n * sizeof(T) + 1023) & ~size_t(1023);
  const size_t off = top_;
  top_ += bytes;
  return reinterpret_cast<T*>(arena_ + off);
}

void ClipGram::reg(const std::string& name, std::vector<int64_t> shape, std::vector<Slot::Dst> dsts) {
  Slot s; s.shape = std::move(shape); s.dsts = std::move(dsts);
  slots_[name] = s;
}
void ClipGram::reg_lin(const std::string& wname, const std::string& bname, int O, int I, Lin& l) {
  l.O = O; l.I = I;
  l.w = walloc<op_t>(size_t(O) * I); l.wt = walloc<op_t>(size_t(I) * O);
  reg(wname, {O, I}, {{Slot::ROWS, l.w, I, 0}, {Slot::ROWS_T, l.wt, O, 0}});
  if (!bname.empty()) { l.b = walloc<float>(O); reg(bname, {O}, {{Slot::F32, l.b, 0, 0}}); }
}

ClipGram::ClipGram(const ClipCfg& cfg) : cfg_(cfg) {
  const int W = cfg.width, P = cfg.patch, G = cfg.resolution / cfg.patch;
  T_ = G * G + 1;
  conv1_.O = W; conv1_.I = 3 * P * P;
  conv1_.w = walloc<op_t>(size_t(W) * 3 * P * P); conv1_.wt = walloc<op_t>(size_t(3) * P * P * W);
  reg("conv1.weight", {W, 3, P, P}, {{Slot::ROWS, conv1_.w, 3 * P * P, 0}, {Slot::ROWS_T, conv1_.wt, W, 0}});
  cls_ = walloc<float>(W); pos_ = walloc<float>(size_t(T_) * W); lnpre_g_ = walloc<float>(W); lnpre_b_ = walloc<float>(W);
  reg("class_embedding", {W}, {{Slot::F32, cls_, 0, 0}});
  reg("positional_embedding", {T_, W}, {{Slot::F32, pos_, 0, 0}});
  reg("ln_pre.weight", {W}, {{Slot::F32, lnpre_g_, 0, 0}}); reg("ln_pre.bias", {W}, {{Slot::F32, lnpre_b_, 0, 0}});
  blocks_.resize(cfg.layers);
  for (int i = 0; i < cfg.layers; ++i) {
    Block& b = blocks_[i];
    const std::string p = "transformer.resblocks." + std::to_string(i);
    b.ln1g = walloc<float>(W); b.ln1b = walloc<float>(W); b.ln2g = walloc<float>(W); b.ln2b = walloc<float>(W);
    reg(p + ".ln_1.weight", {W}, {{Slot::F32, b.ln1g, 0, 0}}); reg(p + ".ln_1.bias", {W}, {{Slot::F32, b.ln1b, 0, 0}});
    reg(p + ".ln_2.weight", {W}, {{Slot::F32, b.ln2g, 0, 0}}); reg(p + ".ln_2.bias", {W}, {{Slot::F32, b.ln2b, 0, 0}});
    reg_lin(p + ".attn.in_proj_weight", p + ".attn.in_proj_bias", 3 * W, W, b.in_proj);
    reg_lin(p + ".attn.out_proj.weight", p + ".attn.out_proj.bias", W, W, b.out_proj);
    reg_lin(p + ".mlp.c_fc.weight", p + ".mlp.c_fc.bias", 4 * W, W, b.fc);
    reg_lin(p + ".mlp.c_proj.weight", p + ".mlp.c_proj.bias", W, 4 * W, b.proj);
  }
  size_t mx = 0;
  for (auto& kv : slots_) { size_t n = 1; for (auto d : kv.second.shape) n *= size_t(d); mx = std::max(mx, n); }
  stage_ = walloc<float>(mx);
  gref_ = walloc<float>(size_t(W) * W);
  // Normalize((2m-1), 2s) of an image in [-1,1] (base_clip.py:38-41)
  const float m[3] = {0.48145466f * 2 - 1, 0.4578275f * 2 - 1, 0.40821073f * 2 - 1}, s[3] = {0.26862954f * 2, 0.26130258f * 2, 0.27577711f * 2};
  const float is[3] = {1.f / s[0], 1.f / s[1], 1.f / s[2]};
  mean_ = walloc<float>(3); inv_std_ = walloc<float>(3);
  cudaMemcpy(mean_, m, sizeof m, cudaMemcpyHostToDevice); cudaMemcpy(inv_std_, is, sizeof is, cudaMemcpyHostToDevice);
  saves_.resize(cfg.layers);
}

ClipGram::~ClipGram() {
  for (void* p : owned_) cudaFree(p);
  if (arena_) cudaFree(arena_);
}

int ClipGram::load_tensor(const char* name, const float* src, const int64_t* dims, int ndim, cudaStream_t st) {
  auto it = slots_.find(name);
  if (it == slots_.end()) return 1;                 // tensors of unused blocks / ln_post / proj are skipped by the caller
  Slot& s = it->second;
  size_t n = 1, want = 1;
  for (int i = 0; i < ndim; ++i) n *= size_t(dims[i]);
  for (auto d : s.shape) want *= size_t(d);
  if (n != want) { err_ = std::string("shape mismatch for ") + name; return -3; }
  CCK(cudaMemcpyAsync(stage_, src, n * sizeof(float), cudaMemcpyDefault, st));
  const int O = int(s.shape[0]), I = int(n / O);
  const int blocks = int(std::min<size_t>((n + 255) / 256, 4096));
  for (auto& d : s.dsts) {
    if (d.kind == Slot::F32) CCK(cudaMemcpyAsync(d.dst, stage_, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    else vae_cvt_weight_kernel<<<blocks, 256, 0, st>>>(stage_, reinterpret_cast<op_t*>(d.dst), O, I, d.kind == Slot::ROWS ? 2 : 3, d.ld, d.off);
  }
  CCK(cudaGetLastError());
  CCK(cudaStreamSynchronize(st));
  s.loaded = true;
  return 0;
}

int ClipGram::finalize(std::string* missing) {
  int n = 0;
  for (auto& kv : slots_)
    if (!kv.second.loaded) { if (missing && n < 8) *missing += kv.first + " "; ++n; }
  if (n) { err_ = "missing weights: " + (missing ? *missing : std::string("?")); return -n; }
  return 0;
}

bool ClipGram::tensor_info(int i, std::string& name, std::vector<int64_t>& shape) const {
  if (i < 0 || i >= int(slots_.size())) return false;
  auto it = slots_.begin();
  std::advance(it, i);
  name = it->first; shape = it->second.shape;
  return true;
}

int ClipGram::gemm(const op_t* Ain, int lda, const op_t* Wt, int M, int N, int K, const GemmEpilogue& ep) {
  GemmParams g; int bn;
  GemmEpilogue e = ep;
  if (e.rows_per_group == 0) e.rows_per_group = 1;
  if (!make_gemm(g, bn, Ain, lda, A_LINEAR, nullptr, Wt, M, N, K, e, err_)) return -1;
  CCK(launch_gemm(g, bn, st_));
  ++launches_;
  return 0;
}

// torch upsample_bicubic2d taps (A = -0.75, align_corners = False, float arithmetic, border clamp) and their transpose
int ClipGram::build_taps(int n_in, int n_out, Taps& fwd, Taps& bwd) {
  std::vector<int> rp(n_out + 1), ix(size_t(n_out) * 4);
  std::vector<float> w(size_t(n_out) * 4);
  const float scale = float(n_in) / float(n_out), Acoef = -0.75f;
  auto c1 = [&](float x) { return ((Acoef + 2) * x - (Acoef + 3)) * x * x + 1; };
  auto c2 = [&](float x) { return ((Acoef * x - 5 * Acoef) * x + 8 * Acoef) * x - 4 * Acoef; };
  for (int o = 0; o < n_out; ++o) {
    const float src = scale * (o + 0.5f) - 0.5f;
    const float fl = std::floor(src);
    const int i0 = int(fl);
    const float t = src - fl;
    const float cw[4] = {c2(t + 1.f), c1(t), c1(1.f - t), c2(2.f - t)};
    rp[o] = 4 * o;
    for (int k = 0; k < 4; ++k) { ix[4 * o + k] = std::min(std::max(i0 - 1 + k, 0), n_in - 1); w[4 * o + k] = cw[k]; }
  }
  rp[n_out] = 4 * n_out;
  std::vector<int> rpt(n_in + 1, 0), ixt(ix.size());
  std::vector<float> wt(w.size());
  for (size_t k = 0; k < ix.size(); ++k) ++rpt[ix[k] + 1];
  for (int i = 0; i < n_in; ++i) rpt[i + 1] += rpt[i];
  std::vector<int> fill(rpt.begin(), rpt.end() - 1);
  for (int o = 0; o < n_out; ++o)
    for (int k = 0; k < 4; ++k) { const int i = ix[4 * o + k]; ixt[fill[i]] = o; wt[fill[i]] = w[4 * o + k]; ++fill[i]; }
  auto up = [&](Taps& t, const std::vector<int>& r, const std::vector<int>& x, const std::vector<float>& ww) -> int {
    t.rowptr = walloc<int>(r.size()); t.idx = walloc<int>(x.size()); t.w = walloc<float>(ww.size());
    if (!t.rowptr || !t.idx || !t.w) return -1;
    cudaMemcpy(t.rowptr, r.data(), r.size() * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(t.idx, x.data(), x.size() * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(t.w, ww.data(), ww.size() * sizeof(float), cudaMemcpyHostToDevice);
    return 0;
  };
  if (up(fwd, rp, ix, w) || up(bwd, rpt, ixt, wt)) return -1;
  return 0;
}

int ClipGram::ensure_arena(int B, int H, int W) {
  const size_t Wd = cfg_.width, R = cfg_.resolution, M = size_t(B) * T_, K0 = size_t(3) * cfg_.patch * cfg_.patch;
  const size_t PP = size_t(B) * cfg_.heads * T_ * T_;
  size_t f32 = size_t(B) * 3 * H * R * 2 + size_t(B) * 3 * R * R * 2 + M * Wd * (4 + size_t(cfg_.layers) * 2 + 6) + size_t(cfg_.layers) * (PP + M * 4 * Wd) +
               PP + M * 4 * Wd * 2 + size_t(B) * Wd * Wd + M * K0 + size_t(B) * 3 * H * W + M * (8 + 4 * size_t(cfg_.layers));
  size_t h16 = M * K0 + M * Wd * 6 + size_t(cfg_.layers) * M * 3 * Wd + M * 4 * Wd * 2 + M * 3 * Wd;
  const size_t need = f32 * 4 + h16 * 2 + (size_t(8) << 20);
  if (need > arena_bytes_) {
    if (arena_) cudaFree(arena_);
    arena_ = nullptr; arena_bytes_ = 0;
    if (cudaMalloc(&arena_, need) != cudaSuccess) { err_ = "CLIP arena cudaMalloc failed"; return -1; }
    arena_bytes_ = need;
  }
  if (taps_H_ != H || taps_W_ != W) {
    if (build_taps(W, cfg_.resolution, tx_f_, tx_b_) || build_taps(H, cfg_.resolution, ty_f_, ty_b_)) { err_ = "resize taps allocation failed"; return -1; }
    taps_H_ = H; taps_W_ = W;
  }
  return 0;
}

// img224 [B][3][R][R] (normalised) -> F [B][T-1][W] = patch-token features after the last evaluated block; fills the tape
int ClipGram::run_features(const float* img224, int B, float** feats_out) {
  const int Wd = cfg_.width, R = cfg_.resolution, P = cfg_.patch, G = R / P, K0 = 3 * P * P, M = B * T_, Hh = cfg_.heads;
  static bool attr = false;
  if (!attr) {
    const int sm_f = 2 * kAttMaxN * kAttLd * 2 + 8 * kAttMaxN * 4 + 8 * 64 * 4;
    cudaFuncSetAttribute(att_small_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_f);
    cudaFuncSetAttribute(att_small_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_f);
    cudaFuncSetAttribute(att_small_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kAttMaxN * kAttLd * 2);
    attr = true;
  }
  const int sm_f = 2 * kAttMaxN * kAttLd * 2 + 8 * kAttMaxN * 4 + 8 * 64 * 4;
  op_t* patches = A<op_t>(size_t(B) * G * G * K0);
  patchify_kernel<<<1024, 256, 0, st_>>>(img224, patches, B, R, P);
  float* xt = A<float>(size_t(M) * Wd);
  class_token_kernel<<<B, 256, 0, st_>>>(xt, cls_, pos_, T_, Wd);
  launches_ += 2;
  for (int b = 0; b < B; ++b) {       // patch embedding + positional embedding (residual operand), written below the class row
    GemmEpilogue e; memset(&e, 0, sizeof e);
    e.residual = pos_ + Wd; e.ldr = Wd; e.out_f32 = xt + (size_t(b) * T_ + 1) * Wd; e.ldo = Wd;
    if (gemm(patches + size_t(b) * G * G * K0, K0, conv1_.w, G * G, Wd, K0, e)) return -1;
  }
  x_tok_ = xt;
  st_pre_ = A<float2>(M);
  float* x = A<float>(size_t(M) * Wd);
  const int lnb = (M + 7) / 8;
  ln_fwd_kernel<<<lnb, 256, 0, st_>>>(xt, lnpre_g_, lnpre_b_, x, nullptr, st_pre_, M, Wd, 1e-5f);
  ++launches_;
  op_t* y16 = A<op_t>(size_t(M) * Wd);
  op_t* att16 = A<op_t>(size_t(M) * Wd);
  op_t* a16 = A<op_t>(size_t(M) * 4 * Wd);
  const float scale = 1.0f / std::sqrt(float(Wd / Hh));
  for (int i = 0; i < cfg_.layers; ++i) {
    const Block& w = blocks_[i];
    BlockSave& sv = saves_[i];
    sv.x_in = x; sv.st1 = A<float2>(M); sv.st2 = A<float2>(M);
    ln_fwd_kernel<<<lnb, 256, 0, st_>>>(x, w.ln1g, w.ln1b, nullptr, y16, sv.st1, M, Wd, 1e-5f);
    sv.qkv = A<op_t>(size_t(M) * 3 * Wd);
    GemmEpilogue e; memset(&e, 0, sizeof e);
    e.bias = w.in_proj.b; e.out_bf16 = sv.qkv; e.ldob = 3 * Wd;
    if (gemm(y16, Wd, w.in_proj.w, M, 3 * Wd, Wd, e)) return -1;
    sv.P = A<float>(size_t(B) * Hh * T_ * T_);
    att_small_fwd_kernel<<<dim3((T_ + 31) / 32, Hh, B), 256, sm_f, st_>>>(sv.qkv, sv.P, att16, T_, Hh, scale);
    sv.x_mid = A<float>(size_t(M) * Wd);
    memset(&e, 0, sizeof e); e.bias = w.out_proj.b; e.residual = x; e.ldr = Wd; e.out_f32 = sv.x_mid; e.ldo = Wd;
    if (gemm(att16, Wd, w.out_proj.w, M, Wd, Wd, e)) return -1;
    ln_fwd_kernel<<<lnb, 256, 0, st_>>>(sv.x_mid, w.ln2g, w.ln2b, nullptr, y16, sv.st2, M, Wd, 1e-5f);
    sv.h = A<float>(size_t(M) * 4 * Wd);
    memset(&e, 0, sizeof e); e.bias = w.fc.b; e.out_f32 = sv.h; e.ldo = 4 * Wd;
    if (gemm(y16, Wd, w.fc.w, M, 4 * Wd, Wd, e)) return -1;
    quickgelu_fwd_kernel<<<1024, 256, 0, st_>>>(sv.h, a16, size_t(M) * 4 * Wd);
    float* xo = A<float>(size_t(M) * Wd);
    memset(&e, 0, sizeof e); e.bias = w.proj.b; e.residual = sv.x_mid; e.ldr = Wd; e.out_f32 = xo; e.ldo = Wd;
    if (gemm(a16, 4 * Wd, w.proj.w, M, Wd, 4 * Wd, e)) return -1;
    x = xo;
    launches_ += 4;
  }
  float* F = A<float>(size_t(B) * (T_ - 1) * Wd);
  drop_class_rows_kernel<<<512, 256, 0, st_>>>(x, F, nullptr, B, T_, Wd);
  ++launches_;
  *feats_out = F;
  CCK(cudaGetLastError());
  return 0;
}

int ClipGram::set_reference(const float* ref, cudaStream_t st) {
  if (ensure_arena(1, cfg_.resolution, cfg_.resolution)) return -1;
  st_ = st; top_ = 0; launches_ = 0;
  float* F;
  if (run_features(ref, 1, &F)) return -1;
  const int Wd = cfg_.width;
  gram_residual_kernel<<<dim3(Wd / 16, Wd / 16, 1), dim3(16, 16), 0, st_>>>(F, nullptr, gref_, T_ - 1, Wd);
  CCK(cudaGetLastError());
  CCK(cudaStreamSynchronize(st_));
  have_ref_ = true; have_tape_ = false;
  return 0;
}

int ClipGram::forward(const float* img, int B, int H, int W, float* loss, cudaStream_t st) {
  if (!have_ref_) { err_ = "set_reference() first"; return -1; }
  if (B < 1 || H < 8 || W < 8) { err_ = "bad image shape"; return -1; }
  if (ensure_arena(B, H, W)) return -1;
  st_ = st; top_ = 0; launches_ = 0; have_tape_ = false;
  const int R = cfg_.resolution, Wd = cfg_.width;
  float* tmpx = A<float>(size_t(B) * 3 * H * R);
  float* img224 = A<float>(size_t(B) * 3 * R * R);
  sparse_resize_kernel<<<dim3(512, B), 256, 0, st_>>>(img, tmpx, tx_f_.rowptr, tx_f_.idx, tx_f_.w, 3, H, W, R, 1, nullptr, nullptr, 0);
  sparse_resize_kernel<<<dim3(512, B), 256, 0, st_>>>(tmpx, img224, ty_f_.rowptr, ty_f_.idx, ty_f_.w, 3, R, H, R, 0, mean_, inv_std_, 0);
  launches_ += 2;
  float* F;
  if (run_features(img224, B, &F)) return -1;
  G_ = A<float>(size_t(B) * Wd * Wd);
  loss_ = A<float>(B);
  gram_residual_kernel<<<dim3(Wd / 16, Wd / 16, B), dim3(16, 16), 0, st_>>>(F, gref_, G_, T_ - 1, Wd);
  frob_norm_kernel<<<B, 256, 0, st_>>>(G_, loss_, size_t(Wd) * Wd);
  launches_ += 2;
  if (loss) CCK(cudaMemcpyAsync(loss, loss_, B * sizeof(float), cudaMemcpyDeviceToDevice, st_));
  CCK(cudaGetLastError());
  F_ = F; B_ = B; H_ = H; W_ = W; fwd_top_ = top_; have_tape_ = true;
  return 0;
}

int ClipGram::backward(float* dimg, cudaStream_t st) {
  if (!have_tape_) { err_ = "backward() needs a preceding forward()"; return -1; }
  st_ = st; top_ = fwd_top_;
  const int B = B_, Wd = cfg_.width, R = cfg_.resolution, P = cfg_.patch, G = R / P, K0 = 3 * P * P, M = B * T_, Hh = cfg_.heads;
  const int sm_f = 2 * kAttMaxN * kAttLd * 2 + 8 * kAttMaxN * 4 + 8 * 64 * 4;
  const float scale = 1.0f / std::sqrt(float(Wd / Hh));
  const int lnb = (M + 7) / 8;
  float* g = A<float>(size_t(M) * Wd);           // dLoss/dx_out of the current block (scaled by kGradScale)
  gram_grad_kernel<<<dim3(Wd / 16, (T_ - 1 + 15) / 16, B), dim3(16, 16), 0, st_>>>(F_, G_, loss_, g, T_ - 1, Wd, kGradScale);
  ++launches_;
  op_t* g16 = A<op_t>(size_t(M) * Wd);
  float* da = A<float>(size_t(M) * 4 * Wd);
  op_t* dh16 = A<op_t>(size_t(M) * 4 * Wd);
  float* dy = A<float>(size_t(M) * Wd);
  op_t* dO16 = A<op_t>(size_t(M) * Wd);
  float* dS = A<float>(size_t(B) * Hh * T_ * T_);
  op_t* dqkv = A<op_t>(size_t(M) * 3 * Wd);
  float* gm = A<float>(size_t(M) * Wd);
  float* g2 = A<float>(size_t(M) * Wd);
  for (int i = cfg_.layers - 1; i >= 0; --i) {
    const Block& w = blocks_[i];
    const BlockSave& sv = saves_[i];
    GemmEpilogue e;
    // MLP branch: x_out = x_mid + proj(quickgelu(fc(ln2(x_mid))))
    cast_rows_kernel<<<1024, 256, 0, st_>>>(g, g16, size_t(M) * Wd);
    memset(&e, 0, sizeof e); e.out_f32 = da; e.ldo = 4 * Wd;
    if (gemm(g16, Wd, w.proj.wt, M, 4 * Wd, Wd, e)) return -1;
    quickgelu_bwd_kernel<<<1024, 256, 0, st_>>>(da, sv.h, dh16, size_t(M) * 4 * Wd);
    memset(&e, 0, sizeof e); e.out_f32 = dy; e.ldo = Wd;
    if (gemm(dh16, 4 * Wd, w.fc.wt, M, Wd, 4 * Wd, e)) return -1;
    ln_bwd_kernel<<<lnb, 256, 0, st_>>>(dy, sv.x_mid, sv.st2, w.ln2g, g, gm, g16, M, Wd);
    // attention branch: x_mid = x_in + out_proj(attn(in_proj(ln1(x_in))))
    memset(&e, 0, sizeof e); e.out_bf16 = dO16; e.ldob = Wd;
    if (gemm(g16, Wd, w.out_proj.wt, M, Wd, Wd, e)) return -1;
    att_small_bwd_dq_kernel<<<dim3((T_ + 31) / 32, Hh, B), 256, sm_f, st_>>>(sv.qkv, sv.P, dO16, dS, dqkv, T_, Hh, scale);
    att_small_bwd_dkv_kernel<<<dim3((T_ + 31) / 32, Hh, B), 256, 2 * kAttMaxN * kAttLd * 2, st_>>>(sv.qkv, sv.P, dS, dO16, dqkv, T_, Hh);
    memset(&e, 0, sizeof e); e.out_f32 = dy; e.ldo = Wd;
    if (gemm(dqkv, 3 * Wd, w.in_proj.wt, M, Wd, 3 * Wd, e)) return -1;
    ln_bwd_kernel<<<lnb, 256, 0, st_>>>(dy, sv.x_in, sv.st1, w.ln1g, gm, g2, nullptr, M, Wd);
    std::swap(g, g2);
    launches_ += 6;
  }
  // ln_pre, class-token drop, patch-embedding dgrad, un-patchify (back to fp32 scale), normalisation, bicubic resize backward
  ln_bwd_kernel<<<lnb, 256, 0, st_>>>(g, x_tok_, st_pre_, lnpre_g_, nullptr, gm, nullptr, M, Wd);
  op_t* gp16 = A<op_t>(size_t(B) * (T_ - 1) * Wd);
  drop_class_rows_kernel<<<512, 256, 0, st_>>>(gm, nullptr, gp16, B, T_, Wd);
  float* gpatch = A<float>(size_t(B) * G * G * K0);
  GemmEpilogue e; memset(&e, 0, sizeof e); e.out_f32 = gpatch; e.ldo = K0;
  if (gemm(gp16, Wd, conv1_.wt, B * G * G, K0, Wd, e)) return -1;
  float* g224 = A<float>(size_t(B) * 3 * R * R);
  unpatchify_kernel<<<1024, 256, 0, st_>>>(gpatch, g224, B, R, P, 1.0f / kGradScale);
  float* tmpy = A<float>(size_t(B) * 3 * H_ * R);
  sparse_resize_kernel<<<dim3(512, B), 256, 0, st_>>>(g224, tmpy, ty_b_.rowptr, ty_b_.idx, ty_b_.w, 3, R, R, H_, 0, nullptr, inv_std_, 1);
  sparse_resize_kernel<<<dim3(512, B), 256, 0, st_>>>(tmpy, dimg, tx_b_.rowptr, tx_b_.idx, tx_b_.w, 3, H_, R, W_, 1, nullptr, nullptr, 0);
  launches_ += 5;
  CCK(cudaGetLastError());
  return 0;
}

}  // namespace hedit
