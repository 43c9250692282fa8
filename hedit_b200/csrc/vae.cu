// VAE decoder (diffusers AutoencoderKL decode direction, SD-1.x topology; restated in oracle/vae.py) forward and input-gradient
// backward on the same hand-written kernels as the UNet engine: tcgen05 implicit-GEMM convs / linears (gemm.cuh) with GroupNorm
// statistics fused into their epilogues, plus the glue kernels of vae.cuh.  This is the decode the reference's style path runs and
// differentiates inside its Langevin loop (text-guided-n-style/inversion/h_edit.py:155-164) and the final latent -> image decode
// (text-guided/main_p2p.py:262-275).
//
// Backward computes dLoss/dz only (weights are frozen).  conv dgrad = the same implicit-GEMM conv over the output gradient with
// transposed, tap-flipped weights prepared at load time; activations needed by the backward (GroupNorm inputs + statistics, the
// attention probabilities) are kept in the arena by decode(keep = true).
#include "vae.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "elementwise.cuh"
#include "tmap.h"
#include "vae.cuh"

namespace hedit {

#define VCK(call)                                                                                  \
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
T* VaeDecoder::walloc(size_t n) {
  void* p = nullptr;
  if (cudaMalloc(&p, std::max<size_t>(n, 4) * sizeof(T)) != cudaSuccess) { err_ = "cudaMalloc failed"; return nullptr; }
  cudaMemset(p, 0, std::max<size_t>(n, 4) * sizeof(T));
  owned_.push_back(p);
  return reinterpret_cast<T*>(p);
}

void VaeDecoder::reg(const std::string& name, std::vector<int64_t> shape, std::vector<Slot::Dst> dsts) {
  Slot s; s.shape = std::move(shape); s.dsts = std::move(dsts);
  slots_[name] = s;
}

void VaeDecoder::reg_conv3(const std::string& name, int O, int I, Conv3W& w) {
  w.O = O; w.I = I;
  w.fwd = walloc<op_t>(size_t(O) * 9 * I); w.dgrad = walloc<op_t>(size_t(I) * 9 * O); w.bias = walloc<float>(O);
  reg(name + ".weight", {O, I, 3, 3}, {{Slot::CONV_FWD, w.fwd, 0, 0}, {Slot::CONV_DGRAD, w.dgrad, 0, 0}});
  reg(name + ".bias", {O}, {{Slot::F32, w.bias, 0, 0}});
}

void VaeDecoder::reg_res(const std::string& name, int cin, int cout, ResW& r) {
  r.cin = cin; r.cout = cout;
  r.n1g = walloc<float>(cin); r.n1b = walloc<float>(cin); r.n2g = walloc<float>(cout); r.n2b = walloc<float>(cout);
  reg(name + ".norm1.weight", {cin}, {{Slot::F32, r.n1g, 0, 0}}); reg(name + ".norm1.bias", {cin}, {{Slot::F32, r.n1b, 0, 0}});
  reg(name + ".norm2.weight", {cout}, {{Slot::F32, r.n2g, 0, 0}}); reg(name + ".norm2.bias", {cout}, {{Slot::F32, r.n2b, 0, 0}});
  reg_conv3(name + ".conv1", cout, cin, r.c1);
  reg_conv3(name + ".conv2", cout, cout, r.c2);
  if (cin != cout) {
    r.wsc = walloc<op_t>(size_t(cout) * cin); r.wsc_t = walloc<op_t>(size_t(cin) * cout); r.bsc = walloc<float>(cout);
    reg(name + ".conv_shortcut.weight", {cout, cin, 1, 1}, {{Slot::ROWS, r.wsc, cin, 0}, {Slot::ROWS_T, r.wsc_t, cout, 0}});
    reg(name + ".conv_shortcut.bias", {cout}, {{Slot::F32, r.bsc, 0, 0}});
  }
}

VaeDecoder::VaeDecoder(const VaeCfg& cfg) : cfg_(cfg) {
  groups_ = cfg.groups;
  const int L = cfg.latent_ch, C3 = cfg.boc[3], C0 = cfg.boc[0];
  pq_w_ = walloc<float>(L * L); pq_b_ = walloc<float>(L);
  reg("post_quant_conv.weight", {L, L, 1, 1}, {{Slot::F32, pq_w_, 0, 0}});
  reg("post_quant_conv.bias", {L}, {{Slot::F32, pq_b_, 0, 0}});
  cin_w_ = walloc<float>(size_t(C3) * L * 9); cin_b_ = walloc<float>(C3); cin_dgrad_ = walloc<op_t>(size_t(L) * 9 * C3);
  reg("decoder.conv_in.weight", {C3, L, 3, 3}, {{Slot::F32, cin_w_, 0, 0}, {Slot::CONV_DGRAD, cin_dgrad_, 0, 0}});
  reg("decoder.conv_in.bias", {C3}, {{Slot::F32, cin_b_, 0, 0}});
  reg_res("decoder.mid_block.resnets.0", C3, C3, mid_[0]);
  reg_res("decoder.mid_block.resnets.1", C3, C3, mid_[1]);
  {
    AttnW& a = attn_; a.C = C3;
    const std::string P = "decoder.mid_block.attentions.0";
    a.gng = walloc<float>(C3); a.gnb = walloc<float>(C3);
    a.w_qkv = walloc<op_t>(size_t(3) * C3 * C3); a.w_qkv_t = walloc<op_t>(size_t(C3) * 3 * C3); a.b_qkv = walloc<float>(3 * C3);
    a.w_o = walloc<op_t>(size_t(C3) * C3); a.w_o_t = walloc<op_t>(size_t(C3) * C3); a.b_o = walloc<float>(C3);
    reg(P + ".group_norm.weight", {C3}, {{Slot::F32, a.gng, 0, 0}}); reg(P + ".group_norm.bias", {C3}, {{Slot::F32, a.gnb, 0, 0}});
    const char* nm[3] = {"to_q", "to_k", "to_v"};
    for (int j = 0; j < 3; ++j) {
      reg(P + "." + nm[j] + ".weight", {C3, C3}, {{Slot::ROWS, a.w_qkv + size_t(j) * C3 * C3, C3, 0}, {Slot::ROWS_T, a.w_qkv_t, 3 * C3, j * C3}});
      reg(P + "." + nm[j] + ".bias", {C3}, {{Slot::F32, a.b_qkv + j * C3, 0, 0}});
    }
    reg(P + ".to_out.0.weight", {C3, C3}, {{Slot::ROWS, a.w_o, C3, 0}, {Slot::ROWS_T, a.w_o_t, C3, 0}});
    reg(P + ".to_out.0.bias", {C3}, {{Slot::F32, a.b_o, 0, 0}});
  }
  int prev = C3;
  for (int i = 0; i < 4; ++i) {
    const int c = cfg.boc[3 - i];
    for (int l = 0; l < cfg.layers + 1; ++l) {
      up_res_[i].emplace_back();
      reg_res("decoder.up_blocks." + std::to_string(i) + ".resnets." + std::to_string(l), l == 0 ? prev : c, c, up_res_[i].back());
    }
    if (i < 3) {
      const std::string nm = "decoder.up_blocks." + std::to_string(i) + ".upsamplers.0.conv";
      reg_conv3(nm, c, c, up_conv_[i]);
      up_phases_[i] = walloc<op_t>(size_t(4) * c * 4 * c);
      slots_[nm + ".weight"].dsts.push_back({Slot::CONV_UP_PHASES, up_phases_[i], 0, 0});
    }
    prev = c;
  }
  no_g_ = walloc<float>(C0); no_b_ = walloc<float>(C0);
  reg("decoder.conv_norm_out.weight", {C0}, {{Slot::F32, no_g_, 0, 0}}); reg("decoder.conv_norm_out.bias", {C0}, {{Slot::F32, no_b_, 0, 0}});
  cout_w_ = walloc<op_t>(size_t(cfg.out_ch) * 9 * C0); cout_b_ = walloc<float>(cfg.out_ch); cout_dgrad_ = walloc<float>(size_t(C0) * 36);
  reg("decoder.conv_out.weight", {cfg.out_ch, C0, 3, 3}, {{Slot::CONV_FWD, cout_w_, 0, 0}, {Slot::CONVOUT_DGRAD, cout_dgrad_, 0, 0}});
  reg("decoder.conv_out.bias", {cfg.out_ch}, {{Slot::F32, cout_b_, 0, 0}});
  size_t mx = 0;
  for (auto& kv : slots_) { size_t n = 1; for (auto d : kv.second.shape) n *= size_t(d); mx = std::max(mx, n); }
  stage_ = walloc<float>(mx);
}

VaeDecoder::~VaeDecoder() {
  for (void* p : owned_) cudaFree(p);
}

int VaeDecoder::load_tensor(const char* name, const float* src, const int64_t* dims, int ndim, cudaStream_t st) {
  auto it = slots_.find(name);
  if (it == slots_.end()) { err_ = std::string("unknown tensor ") + name; return -2; }
  Slot& s = it->second;
  size_t n = 1, want = 1;
  for (int i = 0; i < ndim; ++i) n *= size_t(dims[i]);
  for (auto d : s.shape) want *= size_t(d);
  if (n != want) { err_ = std::string("shape mismatch for ") + name; return -3; }
  VCK(cudaMemcpyAsync(stage_, src, n * sizeof(float), cudaMemcpyDefault, st));
  const int O = int(s.shape[0]), I = s.shape.size() > 1 ? int(s.shape[1]) : 1;
  const int blocks = int(std::min<size_t>((n + 255) / 256, 4096));
  for (auto& d : s.dsts) {
    switch (d.kind) {
      case Slot::F32: VCK(cudaMemcpyAsync(d.dst, stage_, n * sizeof(float), cudaMemcpyDeviceToDevice, st)); break;
      case Slot::CONV_FWD: vae_cvt_weight_kernel<<<blocks, 256, 0, st>>>(stage_, reinterpret_cast<op_t*>(d.dst), O, I, 0, 0, 0); break;
      case Slot::CONV_DGRAD: vae_cvt_weight_kernel<<<blocks, 256, 0, st>>>(stage_, reinterpret_cast<op_t*>(d.dst), O, I, 1, 0, 0); break;
      case Slot::ROWS: vae_cvt_weight_kernel<<<blocks, 256, 0, st>>>(stage_, reinterpret_cast<op_t*>(d.dst), O, I, 2, d.ld, d.off); break;
      case Slot::ROWS_T: vae_cvt_weight_kernel<<<blocks, 256, 0, st>>>(stage_, reinterpret_cast<op_t*>(d.dst), O, I, 3, d.ld, d.off); break;
      case Slot::CONVOUT_DGRAD: vae_cvt_convout_dgrad_kernel<<<blocks, 256, 0, st>>>(stage_, reinterpret_cast<float*>(d.dst), O, I); break;
      case Slot::CONV_UP_PHASES: cvt_upconv_phases_kernel<<<blocks, 256, 0, st>>>(stage_, reinterpret_cast<op_t*>(d.dst), O, I); break;
    }
  }
  VCK(cudaGetLastError());
  VCK(cudaStreamSynchronize(st));
  s.loaded = true;
  return 0;
}

int VaeDecoder::finalize(std::string* missing) {
  int n = 0;
  for (auto& kv : slots_)
    if (!kv.second.loaded) { if (missing && n < 8) *missing += kv.first + " "; ++n; }
  if (n) { err_ = "missing weights: " + (missing ? *missing : std::string("?")); return -n; }
  return 0;
}

bool VaeDecoder::tensor_info(int i, std::string& name, std::vector<int64_t>& shape) const {
  if (i < 0 || i >= int(slots_.size())) return false;
  auto it = slots_.begin();
  std::advance(it, i);
  name = it->first; shape = it->second.shape;
  return true;
}

// GroupNorm(+SiLU) backward: g = dL/d(out) fp32 -> dx (+ add) as fp32 and/or 16-bit
int VaeDecoder::gn_bwd(const float* g, const GNSave& sv, const float* add, float* dx, op_t* dx16) {
  const int chunk = sv.HW >= 65536 ? 128 : (sv.HW >= 4096 ? 64 : 16);
  const int nch = (sv.HW + chunk - 1) / chunk;
  float2* partial = A<float2>(size_t(sv.S) * nch * cfg_.groups);
  float2* red = A<float2>(size_t(sv.S) * cfg_.groups);
  if (dry_) return 0;
  GNBwdParams p{g, sv.x, sv.C, sv.HW, cfg_.groups, chunk, nch, sv.stats, sv.gamma, sv.beta, sv.silu, partial, red, add, dx, dx16};
  const int quads = sv.C / 4;
  gn_bwd_stats_kernel<<<dim3(nch, sv.S), std::min(512, std::max(256, quads * std::max(1, (256 + quads - 1) / quads))), 0, st_>>>(p);
  gn_bwd_reduce_kernel<<<dim3(cfg_.groups, sv.S), 128, 0, st_>>>(partial, red, nch, cfg_.groups, float(1.0 / (double(sv.HW) * (sv.C / cfg_.groups))));
  GNBwdParams q = p; q.chunk = sv.HW >= 4096 ? 32 : 16;
  const int threads = std::max(256, quads * std::max(1, (256 + quads - 1) / quads));
  gn_bwd_apply_kernel<<<dim3((sv.HW + q.chunk - 1) / q.chunk, sv.S), threads, 0, st_>>>(q);
  launches_ += 3;
  VCK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------ ResnetBlock2D (no time embedding)
int VaeDecoder::res_fwd(const ResW& w, ResSave& sv, const float* x, const float2* cs_x, int S, int H, int W, float** out, float2** cs_out) {
  const int HW = H * W, M = S * HW;
  op_t* a1 = A<op_t>(size_t(M) * w.cin);
  op_t* raw = w.wsc ? A<op_t>(size_t(M) * w.cin) : nullptr;
  sv.n1 = GNSave{x, nullptr, S, HW, w.cin, w.n1g, w.n1b, 1};
  if (gn_fwd1(x, cs_x, S, HW, w.cin, w.n1g, w.n1b, 1, a1, raw, &sv.n1.stats)) return -1;
  float* h1 = A<float>(size_t(M) * w.cout);
  GemmEpilogue e1; memset(&e1, 0, sizeof e1);
  e1.bias = w.c1.bias; e1.out_f32 = h1; e1.ldo = w.cout; e1.colstats = colstats_for(M, w.cout, HW);
  if (conv3(a1, w.c1.fwd, S, H, W, w.cin, w.cout, e1)) return -1;
  op_t* a2 = A<op_t>(size_t(M) * w.cout);
  sv.n2 = GNSave{h1, nullptr, S, HW, w.cout, w.n2g, w.n2b, 1};
  if (gn_fwd1(h1, e1.colstats, S, HW, w.cout, w.n2g, w.n2b, 1, a2, nullptr, &sv.n2.stats)) return -1;
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
  e2.bias = w.c2.bias; e2.residual = resid; e2.ldr = w.cout; e2.out_f32 = o; e2.ldo = w.cout; e2.colstats = colstats_for(M, w.cout, HW);
  if (conv3(a2, w.c2.fwd, S, H, W, w.cout, w.cout, e2)) return -1;
  sv.H = H; sv.W = W; sv.S = S;
  *out = o; *cs_out = e2.colstats;
  return 0;
}

// d_out (fp32 + 16-bit) -> d_x (fp32 + 16-bit)
int VaeDecoder::res_bwd(const ResW& w, const ResSave& sv, const float* dout, const op_t* dout16, float** dx, op_t** dx16) {
  const int S = sv.S, H = sv.H, W = sv.W, M = S * H * W;
  GemmEpilogue e; memset(&e, 0, sizeof e);
  float* g2 = A<float>(size_t(M) * w.cout);                 // dL/d(silu(norm2(h1)))
  e.out_f32 = g2; e.ldo = w.cout;
  if (conv3(dout16, w.c2.dgrad, S, H, W, w.cout, w.cout, e)) return -1;
  op_t* dh1 = A<op_t>(size_t(M) * w.cout);
  if (gn_bwd(g2, sv.n2, nullptr, nullptr, dh1)) return -1;
  float* g1 = A<float>(size_t(M) * w.cin);                  // dL/d(silu(norm1(x)))
  memset(&e, 0, sizeof e); e.out_f32 = g1; e.ldo = w.cin;
  if (conv3(dh1, w.c1.dgrad, S, H, W, w.cout, w.cin, e)) return -1;
  const float* skip = dout;                                 // gradient through the shortcut
  if (w.wsc) {
    float* gs = A<float>(size_t(M) * w.cin);
    memset(&e, 0, sizeof e); e.out_f32 = gs; e.ldo = w.cin;
    if (gemm(dout16, w.cout, A_LINEAR, nullptr, w.wsc_t, M, w.cin, w.cout, e)) return -1;
    skip = gs;
  }
  float* o = A<float>(size_t(M) * w.cin);
  op_t* o16 = A<op_t>(size_t(M) * w.cin);
  if (gn_bwd(g1, sv.n1, skip, o, o16)) return -1;
  *dx = o; *dx16 = o16;
  return 0;
}

// ------------------------------------------------------------------------------------------------ mid-block attention (1 head of dim C)
int VaeDecoder::attn_fwd(const float* x, const float2* cs_x, int S, int N, float** out, float2** cs_out) {
  const AttnW& w = attn_;
  AttnSave& sv = attn_sv_;
  sv.gn = GNSave{x, nullptr, S, N, w.C, w.gng, w.gnb, 0};
  if (attn1h_fwd(x, cs_x, S, N, w.C, w.gng, w.gnb, 1e-6f, w.w_qkv, w.b_qkv, w.w_o, w.b_o, out, cs_out, &sv.gn.stats, &sv.qkv, &sv.P)) return -1;
  sv.S = S; sv.N = N;
  return 0;
}

int VaeDecoder::attn_bwd(const float* dout, const op_t* dout16, float** dx, op_t** dx16) {
  const AttnW& w = attn_;
  const AttnSave& sv = attn_sv_;
  const int C = w.C, S = sv.S, N = sv.N, M = S * N;
  const float scale = 1.0f / std::sqrt(float(C));
  GemmEpilogue e; memset(&e, 0, sizeof e);
  op_t* dO = A<op_t>(size_t(M) * C);
  e.out_bf16 = dO; e.ldob = C;
  if (gemm(dout16, C, A_LINEAR, nullptr, w.w_o_t, M, C, C, e)) return -1;
  op_t* dqkv = A<op_t>(size_t(M) * 3 * C);
  op_t* dOt = A<op_t>(size_t(C) * N); op_t* Pt = A<op_t>(size_t(N) * N); op_t* dS = A<op_t>(size_t(N) * N); op_t* dSt = A<op_t>(size_t(N) * N);
  op_t* Kt = A<op_t>(size_t(C) * N); op_t* Qt = A<op_t>(size_t(C) * N);
  float* dP = A<float>(size_t(N) * N);
  const dim3 tb(32, 8);
  for (int s = 0; s < S; ++s) {
    const op_t* qkv = sv.qkv + size_t(s) * N * 3 * C;
    const op_t* P = sv.P + size_t(s) * N * N;
    const op_t* dOs = dO + size_t(s) * N * C;
    op_t* dq = dqkv + size_t(s) * N * 3 * C;
    if (!dry_) {
      transpose_h16_kernel<<<dim3((C + 31) / 32, (N + 31) / 32, 1), tb, 0, st_>>>(dOs, 0, C, dOt, 0, N, N, C);
      transpose_h16_kernel<<<dim3((N + 31) / 32, (N + 31) / 32, 1), tb, 0, st_>>>(P, 0, N, Pt, 0, N, N, N);
      transpose_h16_kernel<<<dim3((C + 31) / 32, (N + 31) / 32, 1), tb, 0, st_>>>(qkv, 0, 3 * C, Qt, 0, N, N, C);
      transpose_h16_kernel<<<dim3((C + 31) / 32, (N + 31) / 32, 1), tb, 0, st_>>>(qkv + C, 0, 3 * C, Kt, 0, N, N, C);
      launches_ += 4;
    }
    // dV = P^T dO
    memset(&e, 0, sizeof e); e.out_bf16 = dq + 2 * C; e.ldob = 3 * C;
    if (gemm(Pt, N, A_LINEAR, nullptr, dOt, N, C, N, e)) return -1;
    // dP = dO V^T
    memset(&e, 0, sizeof e); e.out_f32 = dP; e.ldo = N;
    if (gemm(dOs, C, A_LINEAR, nullptr, qkv + 2 * C, N, N, C, e, 3 * C)) return -1;
    if (!dry_) {
      attn_softmax_bwd_rows_kernel<<<dim3(N, 1), 256, 0, st_>>>(P, dP, dS, N, scale);
      transpose_h16_kernel<<<dim3((N + 31) / 32, (N + 31) / 32, 1), tb, 0, st_>>>(dS, 0, N, dSt, 0, N, N, N);
      launches_ += 2;
    }
    // dQ = dS K ; dK = dS^T Q
    memset(&e, 0, sizeof e); e.out_bf16 = dq; e.ldob = 3 * C;
    if (gemm(dS, N, A_LINEAR, nullptr, Kt, N, C, N, e)) return -1;
    memset(&e, 0, sizeof e); e.out_bf16 = dq + C; e.ldob = 3 * C;
    if (gemm(dSt, N, A_LINEAR, nullptr, Qt, N, C, N, e)) return -1;
  }
  float* gy = A<float>(size_t(M) * C);                       // dL/d(group_norm(x))
  memset(&e, 0, sizeof e); e.out_f32 = gy; e.ldo = C;
  if (gemm(dqkv, 3 * C, A_LINEAR, nullptr, w.w_qkv_t, M, C, 3 * C, e)) return -1;
  float* o = A<float>(size_t(M) * C); op_t* o16 = A<op_t>(size_t(M) * C);
  if (gn_bwd(gy, sv.gn, dout, o, o16)) return -1;            // + residual connection
  *dx = o; *dx16 = o16;
  return 0;
}

// ------------------------------------------------------------------------------------------------ whole decoder
int VaeDecoder::run_forward(const float* z, float* img, int B, int h, int w) {
  const VaeCfg& c = cfg_;
  const int L = c.latent_ch, C3 = c.boc[3];
  float* zq = A<float>(size_t(B) * L * h * w);
  float* x = A<float>(size_t(B) * h * w * C3);
  if (!dry_) {
    vae_pointwise4_kernel<<<dim3(std::max(1, h * w / 256), B), 256, 0, st_>>>(z, pq_w_, pq_b_, zq, L, h * w, 0);
    const size_t sm = (36 * size_t(C3) + 4 * (kConvInRows + 2) * (w + 2)) * sizeof(float);
    cudaFuncSetAttribute(conv_in_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    conv_in_kernel<<<dim3((h + kConvInRows - 1) / kConvInRows, B), 256, sm, st_>>>(zq, cin_w_, cin_b_, x, h, w, C3);
    launches_ += 2;
    VCK(cudaGetLastError());
  }
  float2* cs = nullptr;
  int H = h, W = w;
  float* y; float2* csy;
  if (res_fwd(mid_[0], mid_sv_[0], x, cs, B, H, W, &y, &csy)) return -1;
  x = y; cs = csy;
  if (attn_fwd(x, cs, B, H * W, &y, &csy)) return -1;
  x = y; cs = csy;
  if (res_fwd(mid_[1], mid_sv_[1], x, cs, B, H, W, &y, &csy)) return -1;
  x = y; cs = csy;
  int C = C3;
  for (int i = 0; i < 4; ++i) {
    up_sv_[i].resize(up_res_[i].size());
    for (size_t l = 0; l < up_res_[i].size(); ++l) {
      if (res_fwd(up_res_[i][l], up_sv_[i][l], x, cs, B, H, W, &y, &csy)) return -1;
      x = y; cs = csy;
    }
    C = c.boc[3 - i];
    if (i < 3) {
      float* o; float2* cso;
      if (upconv_fused(x, up_phases_[i], up_conv_[i].bias, B, H, W, C, &o, &cso)) return -1;
      H *= 2; W *= 2;
      x = o; cs = cso;
    }
  }
  op_t* fin = A<op_t>(size_t(B) * H * W * C);
  out_sv_ = GNSave{x, nullptr, B, H * W, C, no_g_, no_b_, 1};
  if (gn_fwd1(x, cs, B, H * W, C, no_g_, no_b_, 1, fin, nullptr, &out_sv_.stats)) return -1;
  GemmEpilogue e; memset(&e, 0, sizeof e);
  e.bias = cout_b_; e.out_f32 = img; e.ldo = c.out_ch; e.nchw_hw = H * W;
  if (conv3(fin, cout_w_, B, H, W, C, c.out_ch, e)) return -1;
  outH_ = H; outW_ = W; outC_ = C; tapeB_ = B; lat_h_ = h; lat_w_ = w;
  return 0;
}

int VaeDecoder::run_backward(const float* dimg, float* dz) {
  const VaeCfg& c = cfg_;
  const int B = tapeB_, H0 = outH_, W0 = outW_, C0 = outC_, L = c.latent_ch;
  // conv_out input gradient: 3 -> C0 channels through conv_in_kernel (CUDA cores, K = 27) on a zero-padded 4-channel NCHW gradient
  float* g4 = A<float>(size_t(B) * 4 * H0 * W0);
  float* gfin = A<float>(size_t(B) * H0 * W0 * C0);
  float* zero_bias = A<float>(C0);
  if (!dry_) {
    VCK(cudaMemsetAsync(zero_bias, 0, C0 * sizeof(float), st_));
    vae_pad4_kernel<<<dim3(std::max(1, 4 * H0 * W0 / 1024), B), 256, 0, st_>>>(dimg, g4, c.out_ch, H0 * W0);
    const size_t sm = (36 * size_t(C0) + 4 * (kConvInRows + 2) * (W0 + 2)) * sizeof(float);
    cudaFuncSetAttribute(conv_in_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    conv_in_kernel<<<dim3((H0 + kConvInRows - 1) / kConvInRows, B), 256, sm, st_>>>(g4, cout_dgrad_, zero_bias, gfin, H0, W0, C0);
    launches_ += 2;
    VCK(cudaGetLastError());
  }
  float* dx = A<float>(size_t(B) * H0 * W0 * C0);
  op_t* dx16 = A<op_t>(size_t(B) * H0 * W0 * C0);
  if (gn_bwd(gfin, out_sv_, nullptr, dx, dx16)) return -1;
  int H = H0, W = W0;
  for (int i = 3; i >= 0; --i) {
    const int C = c.boc[3 - i];
    if (i < 3) {
      // upsampler: conv dgrad at the fine resolution, then sum the 2x2 blocks
      float* gu = A<float>(size_t(B) * H * W * C);
      GemmEpilogue e; memset(&e, 0, sizeof e); e.out_f32 = gu; e.ldo = C;
      if (conv3(dx16, up_conv_[i].dgrad, B, H, W, C, C, e)) return -1;
      H /= 2; W /= 2;
      float* d = A<float>(size_t(B) * H * W * C); op_t* d16 = A<op_t>(size_t(B) * H * W * C);
      if (!dry_) {
        const size_t total = size_t(B) * H * W * (C / 4);
        upsample2x_bwd_kernel<<<int(std::min<size_t>((total + 255) / 256, 16384)), 256, 0, st_>>>(gu, d, d16, B, H, W, C);
        ++launches_;
      }
      dx = d; dx16 = d16;
    }
    for (int l = int(up_res_[i].size()) - 1; l >= 0; --l) {
      float* d; op_t* d16;
      if (res_bwd(up_res_[i][l], up_sv_[i][l], dx, dx16, &d, &d16)) return -1;
      dx = d; dx16 = d16;
    }
  }
  float* d; op_t* d16;
  if (res_bwd(mid_[1], mid_sv_[1], dx, dx16, &d, &d16)) return -1;
  dx = d; dx16 = d16;
  if (attn_bwd(dx, dx16, &d, &d16)) return -1;
  dx = d; dx16 = d16;
  if (res_bwd(mid_[0], mid_sv_[0], dx, dx16, &d, &d16)) return -1;
  dx = d; dx16 = d16;
  // conv_in input gradient (C3 -> 4 channels, NCHW epilogue), then post_quant_conv^T
  float* gz = A<float>(size_t(B) * L * H * W);
  GemmEpilogue e; memset(&e, 0, sizeof e); e.out_f32 = gz; e.ldo = L; e.nchw_hw = H * W;
  if (conv3(dx16, cin_dgrad_, B, H, W, c.boc[3], L, e)) return -1;
  if (!dry_) {
    vae_pointwise4_kernel<<<dim3(std::max(1, H * W / 256), B), 256, 0, st_>>>(gz, pq_w_, nullptr, dz, L, H * W, 1);
    ++launches_;
    VCK(cudaGetLastError());
  }
  return 0;
}

int VaeDecoder::ensure_arena(int B, int h, int w) {
  // sizing pass (forward + backward) with no launches
  dry_ = true; top_ = 0; peak_ = 0; arena_saved_ = arena_; arena_ = nullptr;
  int r = run_forward(nullptr, nullptr, B, h, w);
  if (!r) r = run_backward(nullptr, nullptr);
  dry_ = false; arena_ = arena_saved_;
  if (r) return -1;
  return reserve(peak_ + (size_t(1) << 20), "VAE");
}

int VaeDecoder::decode(const float* z, float* img, int B, int h, int w, cudaStream_t st) {
  if (B < 1 || (w > 128 ? w % 128 != 0 : 128 % w != 0) || h < 1) { err_ = "latent width must divide 128 or be a multiple of it"; return -1; }
  if (ensure_arena(B, h, w)) return -1;
  st_ = st; top_ = 0; launches_ = 0; flops_ = 0;
  have_tape_ = false;
  if (run_forward(z, img, B, h, w)) return -1;
  fwd_top_ = top_;
  have_tape_ = true;
  VCK(cudaGetLastError());
  return 0;
}

int VaeDecoder::backward(const float* dimg, float* dz, cudaStream_t st) {
  if (!have_tape_) { err_ = "backward() needs a preceding decode()"; return -1; }
  st_ = st; top_ = fwd_top_;
  if (run_backward(dimg, dz)) return -1;
  VCK(cudaGetLastError());
  return 0;
}

}  // namespace hedit
