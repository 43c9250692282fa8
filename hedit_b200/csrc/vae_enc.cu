// VAE encoder (diffusers AutoencoderKL encode direction, SD-1.x topology; restated in oracle/vae.py `Encoder`): image -> latent
// moments, the `model.vae.encode(x).latent_dist` of the drivers (text-guided/main_p2p.py:154-159).  Forward only; same kernels as the
// decoder (netexec.cu): wide-image implicit-GEMM convs, stride-2 downsamplers padded (0,1,0,1), fused GroupNorm statistics, the
// materialised single-head mid-block attention.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "elementwise.cuh"
#include "tmap.h"
#include "vae.cuh"
#include "vae_enc.h"

namespace hedit {

#define ECK(call)                                                                                  \
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
T* VaeEncoder::walloc(size_t n) {
  void* p = nullptr;
  if (cudaMalloc(&p, std::max<size_t>(n, 4) * sizeof(T)) != cudaSuccess) { err_ = "cudaMalloc failed"; return nullptr; }
  cudaMemset(p, 0, std::max<size_t>(n, 4) * sizeof(T));
  owned_.push_back(p);
  return reinterpret_cast<T*>(p);
}
void VaeEncoder::reg(const std::string& name, std::vector<int64_t> shape, int kind, void* dst, int ld, int off) {
  Slot s; s.shape = std::move(shape); s.kind = kind; s.dst = dst; s.ld = ld; s.off = off;
  slots_[name] = s;
}
void VaeEncoder::reg_conv3(const std::string& name, int O, int I, Conv3W& w) {
  w.w = walloc<op_t>(size_t(O) * 9 * I); w.b = walloc<float>(O);
  reg(name + ".weight", {O, I, 3, 3}, 1, w.w, 0, 0);
  reg(name + ".bias", {O}, 0, w.b, 0, 0);
}
void VaeEncoder::reg_res(const std::string& name, int cin, int cout, ResW& r) {
  r.cin = cin; r.cout = cout;
  r.n1g = walloc<float>(cin); r.n1b = walloc<float>(cin); r.n2g = walloc<float>(cout); r.n2b = walloc<float>(cout);
  reg(name + ".norm1.weight", {cin}, 0, r.n1g, 0, 0); reg(name + ".norm1.bias", {cin}, 0, r.n1b, 0, 0);
  reg(name + ".norm2.weight", {cout}, 0, r.n2g, 0, 0); reg(name + ".norm2.bias", {cout}, 0, r.n2b, 0, 0);
  reg_conv3(name + ".conv1", cout, cin, r.c1);
  reg_conv3(name + ".conv2", cout, cout, r.c2);
  if (cin != cout) {
    r.wsc = walloc<op_t>(size_t(cout) * cin); r.bsc = walloc<float>(cout);
    reg(name + ".conv_shortcut.weight", {cout, cin, 1, 1}, 2, r.wsc, cin, 0);
    reg(name + ".conv_shortcut.bias", {cout}, 0, r.bsc, 0, 0);
  }
}

VaeEncoder::VaeEncoder(const VaeCfg& cfg) : cfg_(cfg) {
  groups_ = cfg.groups;
  const int L2 = 2 * cfg.latent_ch, C0 = cfg.boc[0], C3 = cfg.boc[3];
  cin_w_ = walloc<float>(size_t(C0) * 36); cin_b_ = walloc<float>(C0);
  reg("encoder.conv_in.weight", {C0, cfg.out_ch, 3, 3}, 3, cin_w_, 0, 0);
  reg("encoder.conv_in.bias", {C0}, 0, cin_b_, 0, 0);
  int prev = C0;
  for (int i = 0; i < 4; ++i) {
    const int c = cfg.boc[i];
    for (int l = 0; l < cfg.layers; ++l) {
      down_[i].emplace_back();
      reg_res("encoder.down_blocks." + std::to_string(i) + ".resnets." + std::to_string(l), l == 0 ? prev : c, c, down_[i].back());
    }
    if (i < 3) reg_conv3("encoder.down_blocks." + std::to_string(i) + ".downsamplers.0.conv", c, c, down_conv_[i]);
    prev = c;
  }
  reg_res("encoder.mid_block.resnets.0", C3, C3, mid_[0]);
  reg_res("encoder.mid_block.resnets.1", C3, C3, mid_[1]);
  {
    const std::string P = "encoder.mid_block.attentions.0";
    a_gng_ = walloc<float>(C3); a_gnb_ = walloc<float>(C3);
    a_wqkv_ = walloc<op_t>(size_t(3) * C3 * C3); a_bqkv_ = walloc<float>(3 * C3); a_wo_ = walloc<op_t>(size_t(C3) * C3); a_bo_ = walloc<float>(C3);
    reg(P + ".group_norm.weight", {C3}, 0, a_gng_, 0, 0); reg(P + ".group_norm.bias", {C3}, 0, a_gnb_, 0, 0);
    const char* nm[3] = {"to_q", "to_k", "to_v"};
    for (int j = 0; j < 3; ++j) {
      reg(P + "." + nm[j] + ".weight", {C3, C3}, 2, a_wqkv_ + size_t(j) * C3 * C3, C3, 0);
      reg(P + "." + nm[j] + ".bias", {C3}, 0, a_bqkv_ + j * C3, 0, 0);
    }
    reg(P + ".to_out.0.weight", {C3, C3}, 2, a_wo_, C3, 0);
    reg(P + ".to_out.0.bias", {C3}, 0, a_bo_, 0, 0);
  }
  no_g_ = walloc<float>(C3); no_b_ = walloc<float>(C3);
  reg("encoder.conv_norm_out.weight", {C3}, 0, no_g_, 0, 0); reg("encoder.conv_norm_out.bias", {C3}, 0, no_b_, 0, 0);
  reg_conv3("encoder.conv_out", L2, C3, conv_out_);
  qw_ = walloc<float>(L2 * L2); qb_ = walloc<float>(L2);
  reg("quant_conv.weight", {L2, L2, 1, 1}, 0, qw_, 0, 0);
  reg("quant_conv.bias", {L2}, 0, qb_, 0, 0);
  size_t mx = 0;
  for (auto& kv : slots_) { size_t m = 1; for (auto d : kv.second.shape) m *= size_t(d); mx = std::max(mx, m); }
  stage_ = walloc<float>(mx);
}

VaeEncoder::~VaeEncoder() {
  for (void* p : owned_) cudaFree(p);
}

__global__ void vae_pad_conv_in_w_kernel(const float* __restrict__ src, float* __restrict__ dst, int O, int I) {      // [O][I][9] -> [O][4][9]
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < O * 36; i += gridDim.x * blockDim.x) {
    const int tap = i % 9, c = (i / 9) % 4, o = i / 36;
    dst[i] = (c < I) ? src[(size_t(o) * I + c) * 9 + tap] : 0.f;
  }
}

int VaeEncoder::load_tensor(const char* name, const float* src, const int64_t* dims, int ndim, cudaStream_t st) {
  auto it = slots_.find(name);
  if (it == slots_.end()) { err_ = std::string("unknown tensor ") + name; return -2; }
  Slot& s = it->second;
  size_t n = 1, want = 1;
  for (int i = 0; i < ndim; ++i) n *= size_t(dims[i]);
  for (auto d : s.shape) want *= size_t(d);
  if (n != want) { err_ = std::string("shape mismatch for ") + name; return -3; }
  ECK(cudaMemcpyAsync(stage_, src, n * sizeof(float), cudaMemcpyDefault, st));
  const int O = int(s.shape[0]), I = s.shape.size() > 1 ? int(s.shape[1]) : 1;
  const int blocks = int(std::min<size_t>((n + 255) / 256, 4096));
  switch (s.kind) {
    case 0: ECK(cudaMemcpyAsync(s.dst, stage_, n * sizeof(float), cudaMemcpyDeviceToDevice, st)); break;
    case 1: vae_cvt_weight_kernel<<<blocks, 256, 0, st>>>(stage_, reinterpret_cast<op_t*>(s.dst), O, I, 0, 0, 0); break;
    case 2: vae_cvt_weight_kernel<<<blocks, 256, 0, st>>>(stage_, reinterpret_cast<op_t*>(s.dst), O, I, 2, s.ld, s.off); break;
    case 3: vae_pad_conv_in_w_kernel<<<64, 256, 0, st>>>(stage_, reinterpret_cast<float*>(s.dst), O, I); break;
  }
  ECK(cudaGetLastError());
  ECK(cudaStreamSynchronize(st));
  s.loaded = true;
  return 0;
}

int VaeEncoder::finalize(std::string* missing) {
  int n = 0;
  for (auto& kv : slots_)
    if (!kv.second.loaded) { if (missing && n < 8) *missing += kv.first + " "; ++n; }
  if (n) { err_ = "missing weights: " + (missing ? *missing : std::string("?")); return -n; }
  return 0;
}

int VaeEncoder::res_fwd(const ResW& w, const float* x, const float2* cs_x, int S, int H, int W, float** out, float2** cs_out) {
  const int HW = H * W, M = S * HW;
  op_t* a1 = A<op_t>(size_t(M) * w.cin);
  op_t* raw = w.wsc ? A<op_t>(size_t(M) * w.cin) : nullptr;
  if (gn_fwd(x, cs_x, w.cin, nullptr, nullptr, 0, S, HW, w.n1g, w.n1b, 1e-6f, 1, a1, raw, nullptr)) return -1;
  float* h1 = A<float>(size_t(M) * w.cout);
  GemmEpilogue e1; memset(&e1, 0, sizeof e1);
  e1.bias = w.c1.b; e1.out_f32 = h1; e1.ldo = w.cout; e1.colstats = colstats_for(M, w.cout, HW);
  if (conv3(a1, w.c1.w, S, H, W, w.cin, w.cout, e1)) return -1;
  op_t* a2 = A<op_t>(size_t(M) * w.cout);
  if (gn_fwd(h1, e1.colstats, w.cout, nullptr, nullptr, 0, S, HW, w.n2g, w.n2b, 1e-6f, 1, a2, nullptr, nullptr)) return -1;
  const float* resid = x;
  if (w.wsc) {
    float* sc = A<float>(size_t(M) * w.cout);
    GemmEpilogue es; memset(&es, 0, sizeof es);
    es.bias = w.bsc; es.out_f32 = sc; es.ldo = w.cout;
    if (gemm(raw, w.cin, A_LINEAR, nullptr, w.wsc, M, w.cout, w.cin, es)) return -1;
    resid = sc;
  }
  float* o = A<float>(size_t(M) * w.cout);
  GemmEpilogue e2; memset(&e2, 0, sizeof e2);
  e2.bias = w.c2.b; e2.residual = resid; e2.ldr = w.cout; e2.out_f32 = o; e2.ldo = w.cout; e2.colstats = colstats_for(M, w.cout, HW);
  if (conv3(a2, w.c2.w, S, H, W, w.cout, w.cout, e2)) return -1;
  *out = o; *cs_out = e2.colstats;
  return 0;
}

int VaeEncoder::run(const float* img, float* moments, int B, int Hin, int Win) {
  const VaeCfg& c = cfg_;
  const int C0 = c.boc[0], L2 = 2 * c.latent_ch;
  float* x4 = A<float>(size_t(B) * 4 * Hin * Win);
  float* x = A<float>(size_t(B) * Hin * Win * C0);
  if (!dry_) {
    vae_pad4_kernel<<<dim3(std::max(1, 4 * Hin * Win / 1024), B), 256, 0, st_>>>(img, x4, c.out_ch, Hin * Win);
    const size_t sm = (36 * size_t(C0) + 4 * (kConvInRows + 2) * (Win + 2)) * sizeof(float);
    cudaFuncSetAttribute(conv_in_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    conv_in_kernel<<<dim3((Hin + kConvInRows - 1) / kConvInRows, B), 256, sm, st_>>>(x4, cin_w_, cin_b_, x, Hin, Win, C0);
    launches_ += 2;
    ECK(cudaGetLastError());
  }
  float2* cs = nullptr;
  int H = Hin, W = Win, C = C0;
  float* y; float2* csy;
  for (int i = 0; i < 4; ++i) {
    for (auto& r : down_[i]) {
      if (res_fwd(r, x, cs, B, H, W, &y, &csy)) return -1;
      x = y; cs = csy;
    }
    C = c.boc[i];
    if (i < 3) {
      op_t* xb = A<op_t>(size_t(B) * H * W * C);
      if (!dry_) { vae_cast_kernel<<<4096, 256, 0, st_>>>(x, xb, size_t(B) * H * W * C / 4); ++launches_; }
      H /= 2; W /= 2;
      float* o = A<float>(size_t(B) * H * W * C);
      GemmEpilogue e; memset(&e, 0, sizeof e);
      e.bias = down_conv_[i].b; e.out_f32 = o; e.ldo = C; e.colstats = colstats_for(B * H * W, C, H * W);
      if (conv3(xb, down_conv_[i].w, B, H, W, C, C, e, 2, 1)) return -1;
      x = o; cs = e.colstats;
    }
  }
  if (res_fwd(mid_[0], x, cs, B, H, W, &y, &csy)) return -1;
  x = y; cs = csy;
  if (attn1h_fwd(x, cs, B, H * W, C, a_gng_, a_gnb_, 1e-6f, a_wqkv_, a_bqkv_, a_wo_, a_bo_, &y, &csy)) return -1;
  x = y; cs = csy;
  if (res_fwd(mid_[1], x, cs, B, H, W, &y, &csy)) return -1;
  x = y; cs = csy;
  op_t* fin = A<op_t>(size_t(B) * H * W * C);
  if (gn_fwd(x, cs, C, nullptr, nullptr, 0, B, H * W, no_g_, no_b_, 1e-6f, 1, fin, nullptr, nullptr)) return -1;
  float* pre = A<float>(size_t(B) * L2 * H * W);
  GemmEpilogue e; memset(&e, 0, sizeof e);
  e.bias = conv_out_.b; e.out_f32 = pre; e.ldo = L2; e.nchw_hw = H * W;
  if (conv3(fin, conv_out_.w, B, H, W, C, L2, e)) return -1;
  if (!dry_) {
    vae_pointwise4_kernel<<<dim3(std::max(1, H * W / 256), B), 256, 0, st_>>>(pre, qw_, qb_, moments, L2, H * W, 0);
    ++launches_;
    ECK(cudaGetLastError());
  }
  return 0;
}

int VaeEncoder::encode(const float* img, float* moments, int B, int H, int W, cudaStream_t st) {
  if (B < 1 || H % 8 || W % 8 || ((W / 8) > 128 ? (W / 8) % 128 != 0 : 128 % (W / 8) != 0)) { err_ = "image size must be a multiple of 8 with latent width dividing 128 (or a multiple of it)"; return -1; }
  uint8_t* saved = arena_;
  dry_ = true; top_ = 0; peak_ = 0; arena_ = nullptr;
  int r = run(nullptr, nullptr, B, H, W);
  dry_ = false; arena_ = saved;
  if (r) return -1;
  if (reserve(peak_ + (size_t(1) << 20), "VAE encoder")) return -1;
  st_ = st; top_ = 0; launches_ = 0; flops_ = 0;
  if (run(img, moments, B, H, W)) return -1;
  ECK(cudaGetLastError());
  return 0;
}

}  // namespace hedit
