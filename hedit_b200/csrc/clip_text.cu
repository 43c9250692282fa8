// CLIP text tower forward (`model.text_encoder(ids)[0]`, transformers CLIPTextModel as called by encode_text,
// text-guided/inversion/inversion_utils.py:13-36): token + position embedding, `layers` pre-LN transformer blocks with causal
// attention and QuickGELU, final LayerNorm.  Linear layers on the tcgen05 GEMM; the 77-token causal attention, LayerNorm and QuickGELU are
// the fp32 kernels shared with the CLIP image branch (clip.cuh).
#include "clip_text.h"

#include <cmath>
#include <cstdio>
#include <cstring>

#include "clip.cuh"
#include "tmap.h"
#include "vae.cuh"

namespace hedit {

#define TCK(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      char buf_[512];                                                                              \
      snprintf(buf_, sizeof buf_, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      err_ = buf_;                                                                                 \
      return -1;                                                                                   \
    }                                                                                              \
  } while (0)

template <typename T>
T* ClipText::walloc(size_t n) {
  void* p = nullptr;
  if (cudaMalloc(&p, std::max<size_t>(n, 4) * sizeof(T)) != cudaSuccess) { err_ = "cudaMalloc failed"; return nullptr; }
  cudaMemset(p, 0, std::max<size_t>(n, 4) * sizeof(T));
  owned_.push_back(p);
  return reinterpret_cast<T*>(p);
}
void ClipText::reg(const std::string& name, std::vector<int64_t> shape, int kind, void* dst, int ld, int off) {
  Slot s; s.shape = std::move(shape); s.kind = kind; s.dst = dst; s.ld = ld; s.off = off;
  slots_[name] = s;
}

ClipText::ClipText(const TextCfg& cfg) : cfg_(cfg) {
  const int W = cfg.width, F = cfg.ffn;
  tok_ = walloc<float>(size_t(cfg.vocab) * W); pos_ = walloc<float>(size_t(cfg.tokens) * W);
  reg("text_model.embeddings.token_embedding.weight", {cfg.vocab, W}, 0, tok_, 0, 0);
  reg("text_model.embeddings.position_embedding.weight", {cfg.tokens, W}, 0, pos_, 0, 0);
  layers_.resize(cfg.layers);
  for (int i = 0; i < cfg.layers; ++i) {
    Layer& l = layers_[i];
    const std::string p = "text_model.encoder.layers." + std::to_string(i);
    auto f = [&](int n) { return walloc<float>(n); };
    l.ln1g = f(W); l.ln1b = f(W); l.ln2g = f(W); l.ln2b = f(W); l.b_qkv = f(3 * W); l.b_o = f(W); l.b_fc1 = f(F); l.b_fc2 = f(W);
    l.w_qkv = walloc<op_t>(size_t(3) * W * W); l.w_o = walloc<op_t>(size_t(W) * W); l.w_fc1 = walloc<op_t>(size_t(F) * W); l.w_fc2 = walloc<op_t>(size_t(W) * F);
    reg(p + ".layer_norm1.weight", {W}, 0, l.ln1g, 0, 0); reg(p + ".layer_norm1.bias", {W}, 0, l.ln1b, 0, 0);
    reg(p + ".layer_norm2.weight", {W}, 0, l.ln2g, 0, 0); reg(p + ".layer_norm2.bias", {W}, 0, l.ln2b, 0, 0);
    const char* nm[3] = {"q_proj", "k_proj", "v_proj"};
    for (int j = 0; j < 3; ++j) {
      reg(p + ".self_attn." + nm[j] + ".weight", {W, W}, 2, l.w_qkv + size_t(j) * W * W, W, 0);
      reg(p + ".self_attn." + nm[j] + ".bias", {W}, 0, l.b_qkv + j * W, 0, 0);
    }
    reg(p + ".self_attn.out_proj.weight", {W, W}, 2, l.w_o, W, 0); reg(p + ".self_attn.out_proj.bias", {W}, 0, l.b_o, 0, 0);
    reg(p + ".mlp.fc1.weight", {F, W}, 2, l.w_fc1, W, 0); reg(p + ".mlp.fc1.bias", {F}, 0, l.b_fc1, 0, 0);
    reg(p + ".mlp.fc2.weight", {W, F}, 2, l.w_fc2, F, 0); reg(p + ".mlp.fc2.bias", {W}, 0, l.b_fc2, 0, 0);
  }
  lnf_g_ = walloc<float>(W); lnf_b_ = walloc<float>(W);
  reg("text_model.final_layer_norm.weight", {W}, 0, lnf_g_, 0, 0); reg("text_model.final_layer_norm.bias", {W}, 0, lnf_b_, 0, 0);
  size_t mx = 0;
  for (auto& kv : slots_) { size_t m = 1; for (auto d : kv.second.shape) m *= size_t(d); mx = std::max(mx, m); }
  stage_ = walloc<float>(mx);
}

ClipText::~ClipText() {
  for (void* p : owned_) cudaFree(p);
  if (ids_dev_) cudaFree(ids_dev_);
}

int ClipText::load_tensor(const char* name, const float* src, const int64_t* dims, int ndim, cudaStream_t st) {
  auto it = slots_.find(name);
  if (it == slots_.end()) return 1;
  Slot& s = it->second;
  size_t n = 1, want = 1;
  for (int i = 0; i < ndim; ++i) n *= size_t(dims[i]);
  for (auto d : s.shape) want *= size_t(d);
  if (n != want) { err_ = std::string("shape mismatch for ") + name; return -3; }
  TCK(cudaMemcpyAsync(stage_, src, n * sizeof(float), cudaMemcpyDefault, st));
  const int O = int(s.shape[0]), I = s.shape.size() > 1 ? int(s.shape[1]) : 1;
  if (s.kind == 0) TCK(cudaMemcpyAsync(s.dst, stage_, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  else vae_cvt_weight_kernel<<<int(std::min<size_t>((n + 255) / 256, 4096)), 256, 0, st>>>(stage_, reinterpret_cast<op_t*>(s.dst), O, I, 2, s.ld, s.off);
  TCK(cudaGetLastError());
  TCK(cudaStreamSynchronize(st));
  s.loaded = true;
  return 0;
}

int ClipText::finalize(std::string* missing) {
  int n = 0;
  for (auto& kv : slots_)
    if (!kv.second.loaded) { if (missing && n < 8) *missing += kv.first + " "; ++n; }
  if (n) { err_ = "missing weights: " + (missing ? *missing : std::string("?")); return -n; }
  return 0;
}

int ClipText::run(const int* ids_dev, int B, float* out) {
  const int W = cfg_.width, F = cfg_.ffn, T = cfg_.tokens, H = cfg_.heads, M = B * T;
  const int sm_f = 2 * kAttMaxN * kAttLd * 2 + 8 * kAttMaxN * 4 + 8 * 64 * 4;
  float* x = A<float>(size_t(M) * W);
  op_t* y16 = A<op_t>(size_t(M) * W);
  op_t* qkv = A<op_t>(size_t(M) * 3 * W);
  op_t* att16 = A<op_t>(size_t(M) * W);
  float* h = A<float>(size_t(M) * F);
  op_t* a16 = A<op_t>(size_t(M) * F);
  float* x2 = A<float>(size_t(M) * W);
  if (dry_) { flops_ = 0; return 0; }
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(att_small_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_f); attr = true; }
  embed_tokens_kernel<<<M, 256, 0, st_>>>(ids_dev, tok_, pos_, x, T, W, cfg_.vocab);
  ++launches_;
  const int lnb = (M + 7) / 8;
  const float scale = 1.0f / std::sqrt(float(W / H));
  for (const Layer& l : layers_) {
    ln_fwd_kernel<<<lnb, 256, 0, st_>>>(x, l.ln1g, l.ln1b, nullptr, y16, nullptr, M, W, 1e-5f);
    GemmEpilogue e; memset(&e, 0, sizeof e);
    e.bias = l.b_qkv; e.out_bf16 = qkv; e.ldob = 3 * W;
    if (gemm(y16, W, A_LINEAR, nullptr, l.w_qkv, M, 3 * W, W, e)) return -1;
    att_small_fwd_kernel<<<dim3((T + 31) / 32, H, B), 256, sm_f, st_>>>(qkv, nullptr, att16, T, H, scale, 1);
    memset(&e, 0, sizeof e); e.bias = l.b_o; e.residual = x; e.ldr = W; e.out_f32 = x2; e.ldo = W;
    if (gemm(att16, W, A_LINEAR, nullptr, l.w_o, M, W, W, e)) return -1;
    ln_fwd_kernel<<<lnb, 256, 0, st_>>>(x2, l.ln2g, l.ln2b, nullptr, y16, nullptr, M, W, 1e-5f);
    memset(&e, 0, sizeof e); e.bias = l.b_fc1; e.out_f32 = h; e.ldo = F;
    if (gemm(y16, W, A_LINEAR, nullptr, l.w_fc1, M, F, W, e)) return -1;
    quickgelu_fwd_kernel<<<1024, 256, 0, st_>>>(h, a16, size_t(M) * F);
    memset(&e, 0, sizeof e); e.bias = l.b_fc2; e.residual = x2; e.ldr = W; e.out_f32 = x; e.ldo = W;
    if (gemm(a16, F, A_LINEAR, nullptr, l.w_fc2, M, W, F, e)) return -1;
    launches_ += 4;
  }
  ln_fwd_kernel<<<lnb, 256, 0, st_>>>(x, lnf_g_, lnf_b_, out, nullptr, nullptr, M, W, 1e-5f);
  ++launches_;
  TCK(cudaGetLastError());
  return 0;
}

int ClipText::forward(const int32_t* ids_host, int B, float* out, cudaStream_t st) {
  if (B < 1) { err_ = "bad batch"; return -1; }
  const int n = B * cfg_.tokens;
  if (n > ids_cap_) {
    if (ids_dev_) cudaFree(ids_dev_);
    if (cudaMalloc(&ids_dev_, n * sizeof(int)) != cudaSuccess) { ids_dev_ = nullptr; ids_cap_ = 0; err_ = "cudaMalloc failed"; return -1; }
    ids_cap_ = n;
  }
  uint8_t* saved = arena_;
  dry_ = true; top_ = 0; peak_ = 0; arena_ = nullptr;
  run(nullptr, B, nullptr);
  dry_ = false; arena_ = saved;
  if (reserve(peak_ + (size_t(1) << 20), "CLIP text")) return -1;
  st_ = st; top_ = 0; launches_ = 0; flops_ = 0;
  TCK(cudaMemcpyAsync(ids_dev_, ids_host, n * sizeof(int), cudaMemcpyHostToDevice, st));
  return run(ids_dev_, B, out);
}

}  // namespace hedit
