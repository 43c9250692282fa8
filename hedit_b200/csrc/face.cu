// Forward of the pixel-space DDPM UNet the face-swapping path denoises with (face-swapping/diffusion/diffusion.py:193-341 `Model`;
// restated in oracle/face_unet.py): ResnetBlocks with a time-embedding add and 1x1 shortcuts over skip-concatenated inputs, single-head
// attention blocks at one resolution, stride-2 downsamplers padded (0,1,0,1), nearest-2x + conv upsamplers.  Same kernels as the VAE
// decoder (netexec.cu): tcgen05 implicit-GEMM convs incl. images wider than one tile and the asymmetric stride-2 variant, GroupNorm
// statistics from the producing GEMM's epilogue (also across the two halves of a skip concatenation), materialised d = C attention.
#include "face.h"

#include <cmath>
#include <cstdio>
#include <cstring>

#include "elementwise.cuh"
#include "tmap.h"
#include "vae.cuh"

namespace hedit {

#define FCK(call)                                                                                  \
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
T* FaceUNet::walloc(size_t n) {
  void* p = nullptr;
  if (cudaMalloc(&p, std::max<size_t>(n, 4) * sizeof(T)) != cudaSuccess) { err_ = "cudaMalloc failed"; return nullptr; }
  cudaMemset(p, 0, std::max<size_t>(n, 4) * sizeof(T));
  owned_.push_back(p);
  return reinterpret_cast<T*>(p);
}
void FaceUNet::reg(const std::string& name, std::vector<int64_t> shape, std::vector<Slot::Dst> dsts) {
  Slot s; s.shape = std::move(shape); s.dsts = std::move(dsts);
  slots_[name] = s;
}
void FaceUNet::reg_conv3(const std::string& name, int O, int I, Conv3W& w) {
  w.O = O; w.I = I;
  w.w = walloc<op_t>(size_t(O) * 9 * I); w.b = walloc<float>(O);
  reg(name + ".weight", {O, I, 3, 3}, {{Slot::CONV_FWD, w.w, 0, 0}});
  reg(name + ".bias", {O}, {{Slot::F32, w.b, 0, 0}});
}
void FaceUNet::reg_res(const std::string& name, int cin, int cout, ResW& r) {
  r.cin = cin; r.cout = cout; r.temb_off = tproj_total_;
  tproj_total_ += cout;
  r.n1g = walloc<float>(cin); r.n1b = walloc<float>(cin); r.n2g = walloc<float>(cout); r.n2b = walloc<float>(cout);
  reg(name + ".norm1.weight", {cin}, {{Slot::F32, r.n1g, 0, 0}}); reg(name + ".norm1.bias", {cin}, {{Slot::F32, r.n1b, 0, 0}});
  reg(name + ".norm2.weight", {cout}, {{Slot::F32, r.n2g, 0, 0}}); reg(name + ".norm2.bias", {cout}, {{Slot::F32, r.n2b, 0, 0}});
  reg_conv3(name + ".conv1", cout, cin, r.c1);
  reg_conv3(name + ".conv2", cout, cout, r.c2);
  if (cin != cout) {
    r.wsc = walloc<op_t>(size_t(cout) * cin); r.bsc = walloc<float>(cout);
    reg(name + ".nin_shortcut.weight", {cout, cin, 1, 1}, {{Slot::ROWS, r.wsc, cin, 0}});
    reg(name + ".nin_shortcut.bias", {cout}, {{Slot::F32, r.bsc, 0, 0}});
  }
}
void FaceUNet::reg_attn(const std::string& name, int C, AttnW& a) {
  a.C = C;
  a.gng = walloc<float>(C); a.gnb = walloc<float>(C); a.w_qkv = walloc<op_t>(size_t(3) * C * C); a.b_qkv = walloc<float>(3 * C);
  a.w_o = walloc<op_t>(size_t(C) * C); a.b_o = walloc<float>(C);
  reg(name + ".norm.weight", {C}, {{Slot::F32, a.gng, 0, 0}}); reg(name + ".norm.bias", {C}, {{Slot::F32, a.gnb, 0, 0}});
  const char* nm[3] = {"q", "k", "v"};
  for (int j = 0; j < 3; ++j) {
    reg(name + "." + nm[j] + ".weight", {C, C, 1, 1}, {{Slot::ROWS, a.w_qkv + size_t(j) * C * C, C, 0}});
    reg(name + "." + nm[j] + ".bias", {C}, {{Slot::F32, a.b_qkv + j * C, 0, 0}});
  }
  reg(name + ".proj_out.weight", {C, C, 1, 1}, {{Slot::ROWS, a.w_o, C, 0}});
  reg(name + ".proj_out.bias", {C}, {{Slot::F32, a.b_o, 0, 0}});
}

FaceUNet::FaceUNet(const FaceCfg& cfg) : cfg_(cfg) {
  groups_ = 32;
  const int ch = cfg.ch, tch = 4 * ch, n = cfg.nlevels;
  t_w1_ = walloc<float>(size_t(tch) * ch); t_b1_ = walloc<float>(tch); t_w2_ = walloc<float>(size_t(tch) * tch); t_b2_ = walloc<float>(tch);
  reg("temb.dense.0.weight", {tch, ch}, {{Slot::F32, t_w1_, 0, 0}}); reg("temb.dense.0.bias", {tch}, {{Slot::F32, t_b1_, 0, 0}});
  reg("temb.dense.1.weight", {tch, tch}, {{Slot::F32, t_w2_, 0, 0}}); reg("temb.dense.1.bias", {tch}, {{Slot::F32, t_b2_, 0, 0}});
  cin_w_ = walloc<float>(size_t(ch) * 4 * 9); cin_b_ = walloc<float>(ch);       // 4th input channel stays zero
  reg("conv_in.bias", {ch}, {{Slot::F32, cin_b_, 0, 0}});
  reg("conv_in.weight", {ch, cfg.in_ch, 3, 3}, {{Slot::F32, nullptr, 0, 0}});     // scattered into the padded layout by load_tensor
  std::vector<int> in_mult(1, 1);
  for (int i = 0; i < n; ++i) in_mult.push_back(cfg.mult[i]);
  int res = cfg.resolution, cur = ch;
  down_.resize(n); up_.resize(n);
  for (int i = 0; i < n; ++i) {
    cur = ch * in_mult[i];
    for (int j = 0; j < cfg.nres; ++j) {
      down_[i].res.emplace_back();
      reg_res("down." + std::to_string(i) + ".block." + std::to_string(j), cur, ch * cfg.mult[i], down_[i].res.back());
      cur = ch * cfg.mult[i];
      if (res == cfg.attn_res) { down_[i].attn.emplace_back(); reg_attn("down." + std::to_string(i) + ".attn." + std::to_string(j), cur, down_[i].attn.back()); }
    }
    if (i != n - 1) { down_[i].has_resample = true; reg_conv3("down." + std::to_string(i) + ".downsample.conv", cur, cur, down_[i].resample); res /= 2; }
  }
  reg_res("mid.block_1", cur, cur, mid1_);
  reg_attn("mid.attn_1", cur, mid_attn_);
  reg_res("mid.block_2", cur, cur, mid2_);
  for (int i = n - 1; i >= 0; --i) {
    int skip = ch * cfg.mult[i];
    for (int j = 0; j < cfg.nres + 1; ++j) {
      if (j == cfg.nres) skip = ch * in_mult[i];
      up_[i].res.emplace_back();
      reg_res("up." + std::to_string(i) + ".block." + std::to_string(j), cur + skip, ch * cfg.mult[i], up_[i].res.back());
      cur = ch * cfg.mult[i];
      if (res == cfg.attn_res) { up_[i].attn.emplace_back(); reg_attn("up." + std::to_string(i) + ".attn." + std::to_string(j), cur, up_[i].attn.back()); }
    }
    if (i != 0) {
      up_[i].has_resample = true;
      const std::string nm = "up." + std::to_string(i) + ".upsample.conv";
      reg_conv3(nm, cur, cur, up_[i].resample);
      up_[i].up_phases = walloc<op_t>(size_t(4) * cur * 4 * cur);
      slots_[nm + ".weight"].dsts.push_back({Slot::CONV_UP_PHASES, up_[i].up_phases, 0, 0});
      res *= 2;
    }
  }
  no_g_ = walloc<float>(cur); no_b_ = walloc<float>(cur);
  reg("norm_out.weight", {cur}, {{Slot::F32, no_g_, 0, 0}}); reg("norm_out.bias", {cur}, {{Slot::F32, no_b_, 0, 0}});
  reg_conv3("conv_out", cfg.out_ch, cur, conv_out_);
  // time-embedding projections of every ResnetBlock, concatenated (evaluated once per call)
  tproj_w_ = walloc<float>(size_t(tproj_total_) * tch); tproj_b_ = walloc<float>(tproj_total_);
  auto reg_tp = [&](const std::string& name, const ResW& r) {
    reg(name + ".temb_proj.weight", {r.cout, tch}, {{Slot::F32, tproj_w_ + size_t(r.temb_off) * tch, 0, 0}});
    reg(name + ".temb_proj.bias", {r.cout}, {{Slot::F32, tproj_b_ + r.temb_off, 0, 0}});
  };
  for (int i = 0; i < n; ++i) {
    for (size_t j = 0; j < down_[i].res.size(); ++j) reg_tp("down." + std::to_string(i) + ".block." + std::to_string(j), down_[i].res[j]);
    for (size_t j = 0; j < up_[i].res.size(); ++j) reg_tp("up." + std::to_string(i) + ".block." + std::to_string(j), up_[i].res[j]);
  }
  reg_tp("mid.block_1", mid1_); reg_tp("mid.block_2", mid2_);
  size_t mx = 0;
  for (auto& kv : slots_) { size_t m = 1; for (auto d : kv.second.shape) m *= size_t(d); mx = std::max(mx, m); }
  stage_ = walloc<float>(mx);
}

FaceUNet::~FaceUNet() {
  for (void* p : owned_) cudaFree(p);
}

__global__ void face_pad_conv_in_w_kernel(const float* __restrict__ src, float* __restrict__ dst, int O, int I) {      // [O][I][9] -> [O][4][9]
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < O * 36; i += gridDim.x * blockDim.x) {
    const int tap = i % 9, c = (i / 9) % 4, o = i / 36;
    dst[i] = (c < I) ? src[(size_t(o) * I + c) * 9 + tap] : 0.f;
  }
}

int FaceUNet::load_tensor(const char* name, const float* src, const int64_t* dims, int ndim, cudaStream_t st) {
  auto it = slots_.find(name);
  if (it == slots_.end()) { err_ = std::string("unknown tensor ") + name; return -2; }
  Slot& s = it->second;
  size_t n = 1, want = 1;
  for (int i = 0; i < ndim; ++i) n *= size_t(dims[i]);
  for (auto d : s.shape) want *= size_t(d);
  if (n != want) { err_ = std::string("shape mismatch for ") + name; return -3; }
  FCK(cudaMemcpyAsync(stage_, src, n * sizeof(float), cudaMemcpyDefault, st));
  const int O = int(s.shape[0]), I = s.shape.size() > 1 ? int(s.shape[1]) : 1;
  const int blocks = int(std::min<size_t>((n + 255) / 256, 4096));
  if (std::string(name) == "conv_in.weight") {
    face_pad_conv_in_w_kernel<<<64, 256, 0, st>>>(stage_, cin_w_, O, I);
  } else {
    for (auto& d : s.dsts) {
      switch (d.kind) {
        case Slot::F32: FCK(cudaMemcpyAsync(d.dst, stage_, n * sizeof(float), cudaMemcpyDeviceToDevice, st)); break;
        case Slot::CONV_FWD: vae_cvt_weight_kernel<<<blocks, 256, 0, st>>>(stage_, reinterpret_cast<op_t*>(d.dst), O, I, 0, 0, 0); break;
        case Slot::ROWS: vae_cvt_weight_kernel<<<blocks, 256, 0, st>>>(stage_, reinterpret_cast<op_t*>(d.dst), O, I, 2, d.ld, d.off); break;
        case Slot::CONV_UP_PHASES: cvt_upconv_phases_kernel<<<blocks, 256, 0, st>>>(stage_, reinterpret_cast<op_t*>(d.dst), O, I); break;
      }
    }
  }
  FCK(cudaGetLastError());
  FCK(cudaStreamSynchronize(st));
  s.loaded = true;
  return 0;
}

int FaceUNet::finalize(std::string* missing) {
  int n = 0;
  for (auto& kv : slots_)
    if (!kv.second.loaded) { if (missing && n < 8) *missing += kv.first + " "; ++n; }
  if (n) { err_ = "missing weights: " + (missing ? *missing : std::string("?")); return -n; }
  return 0;
}

bool FaceUNet::tensor_info(int i, std::string& name, std::vector<int64_t>& shape) const {
  if (i < 0 || i >= int(slots_.size())) return false;
  auto it = slots_.begin();
  std::advance(it, i);
  name = it->first; shape = it->second.shape;
  return true;
}

// ResnetBlock (diffusion.py:74-137) on the channel concatenation [a | skip]
int FaceUNet::res_fwd(const ResW& w, const Act& a, const Act* skip, int S, int H, int W, Act* out) {
  const int HW = H * W, M = S * HW;
  const int C2 = skip ? skip->C : 0;
  if (a.C + C2 != w.cin) { err_ = "internal: resblock channel mismatch"; return -1; }
  op_t* a1 = A<op_t>(size_t(M) * w.cin);
  op_t* raw = w.wsc ? A<op_t>(size_t(M) * w.cin) : nullptr;
  if (gn_fwd(a.x, a.cs, a.C, skip ? skip->x : nullptr, skip ? skip->cs : nullptr, C2, S, HW, w.n1g, w.n1b, 1e-6f, 1, a1, raw, nullptr)) return -1;
  float* h1 = A<float>(size_t(M) * w.cout);
  GemmEpilogue e1; memset(&e1, 0, sizeof e1);
  e1.bias = w.c1.b; e1.rowvec = temb_rows_ + w.temb_off; e1.ldrv = tproj_total_; e1.rows_per_group = HW; e1.out_f32 = h1; e1.ldo = w.cout;
  e1.colstats = colstats_for(M, w.cout, HW);
  if (conv3(a1, w.c1.w, S, H, W, w.cin, w.cout, e1)) return -1;
  op_t* a2 = A<op_t>(size_t(M) * w.cout);
  if (gn_fwd(h1, e1.colstats, w.cout, nullptr, nullptr, 0, S, HW, w.n2g, w.n2b, 1e-6f, 1, a2, nullptr, nullptr)) return -1;
  const float* resid = a.x;
  if (w.wsc) {
    float* sc = A<float>(size_t(M) * w.cout);
    GemmEpilogue es; memset(&es, 0, sizeof es);
    es.bias = w.bsc; es.out_f32 = sc; es.ldo = w.cout;
    if (gemm(raw, w.cin, A_LINEAR, nullptr, w.wsc, M, w.cout, w.cin, es)) return -1;
    resid = sc;
  } else if (skip) { err_ = "internal: concatenated input needs a shortcut projection"; return -1; }
  float* o = A<float>(size_t(M) * w.cout);
  GemmEpilogue e2; memset(&e2, 0, sizeof e2);
  e2.bias = w.c2.b; e2.residual = resid; e2.ldr = w.cout; e2.out_f32 = o; e2.ldo = w.cout; e2.colstats = colstats_for(M, w.cout, HW);
  if (conv3(a2, w.c2.w, S, H, W, w.cout, w.cout, e2)) return -1;
  *out = Act{o, e2.colstats, w.cout};
  return 0;
}

int FaceUNet::attn_fwd(const AttnW& w, const Act& a, int S, int N, Act* out) {
  float* o; float2* cs;
  if (attn1h_fwd(a.x, a.cs, S, N, w.C, w.gng, w.gnb, 1e-6f, w.w_qkv, w.b_qkv, w.w_o, w.b_o, &o, &cs)) return -1;
  *out = Act{o, cs, w.C};
  return 0;
}

int FaceUNet::run(const float* x, float* eps, int S) {
  const FaceCfg& c = cfg_;
  const int R = c.resolution, n = c.nlevels;
  float* x4 = A<float>(size_t(S) * 4 * R * R);
  float* h0 = A<float>(size_t(S) * R * R * c.ch);
  if (!dry_) {
    vae_pad4_kernel<<<dim3(std::max(1, 4 * R * R / 1024), S), 256, 0, st_>>>(x, x4, c.in_ch, R * R);
    const size_t sm = (36 * size_t(c.ch) + 4 * (kConvInRows + 2) * (R + 2)) * sizeof(float);
    cudaFuncSetAttribute(conv_in_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    conv_in_kernel<<<dim3((R + kConvInRows - 1) / kConvInRows, S), 256, sm, st_>>>(x4, cin_w_, cin_b_, h0, R, R, c.ch);
    launches_ += 2;
    FCK(cudaGetLastError());
  }
  std::vector<Act> hs;
  hs.push_back(Act{h0, nullptr, c.ch});
  int H = R;
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < c.nres; ++j) {
      Act h;
      if (res_fwd(down_[i].res[j], hs.back(), nullptr, S, H, H, &h)) return -1;
      if (!down_[i].attn.empty()) { Act g; if (attn_fwd(down_[i].attn[j], h, S, H * H, &g)) return -1; h = g; }
      hs.push_back(h);
    }
    if (down_[i].has_resample) {
      const Act& a = hs.back();
      op_t* xb = A<op_t>(size_t(S) * H * H * a.C);
      if (!dry_) { vae_cast_kernel<<<4096, 256, 0, st_>>>(a.x, xb, size_t(S) * H * H * a.C / 4); ++launches_; }
      H /= 2;
      float* y = A<float>(size_t(S) * H * H * a.C);
      GemmEpilogue e; memset(&e, 0, sizeof e);
      e.bias = down_[i].resample.b; e.out_f32 = y; e.ldo = a.C; e.colstats = colstats_for(S * H * H, a.C, H * H);
      if (conv3(xb, down_[i].resample.w, S, H, H, a.C, a.C, e, 2, 1)) return -1;
      hs.push_back(Act{y, e.colstats, a.C});
    }
  }
  Act h = hs.back(), g;
  if (res_fwd(mid1_, h, nullptr, S, H, H, &g)) return -1;
  h = g;
  if (attn_fwd(mid_attn_, h, S, H * H, &g)) return -1;
  h = g;
  if (res_fwd(mid2_, h, nullptr, S, H, H, &g)) return -1;
  h = g;
  for (int i = n - 1; i >= 0; --i) {
    for (int j = 0; j < c.nres + 1; ++j) {
      Act sk = hs.back(); hs.pop_back();
      if (res_fwd(up_[i].res[j], h, &sk, S, H, H, &g)) return -1;
      h = g;
      if (!up_[i].attn.empty()) { if (attn_fwd(up_[i].attn[j], h, S, H * H, &g)) return -1; h = g; }
    }
    if (up_[i].has_resample) {
      float* y; float2* csy;
      if (upconv_fused(h.x, up_[i].up_phases, up_[i].resample.b, S, H, H, h.C, &y, &csy)) return -1;
      H *= 2;
      h = Act{y, csy, h.C};
    }
  }
  op_t* fin = A<op_t>(size_t(S) * H * H * h.C);
  if (gn_fwd(h.x, h.cs, h.C, nullptr, nullptr, 0, S, H * H, no_g_, no_b_, 1e-6f, 1, fin, nullptr, nullptr)) return -1;
  GemmEpilogue e; memset(&e, 0, sizeof e);
  e.bias = conv_out_.b; e.out_f32 = eps; e.ldo = c.out_ch; e.nchw_hw = H * H;
  if (conv3(fin, conv_out_.w, S, H, H, h.C, c.out_ch, e)) return -1;
  return 0;
}

int FaceUNet::forward(const float* x, const float* t_host, float* eps, int S, cudaStream_t st) {
  if (S < 1) { err_ = "bad batch"; return -1; }
  const int tch = 4 * cfg_.ch;
  if (S > max_S_) {
    drop_graphs();                       // they address the old time-embedding buffers
    ts_dev_ = walloc<float>(S); temb_a_ = walloc<float>(size_t(S) * tch); temb_b_ = walloc<float>(size_t(S) * tch);
    temb_rows_ = walloc<float>(size_t(S) * tproj_total_);
    if (!ts_dev_ || !temb_a_ || !temb_b_ || !temb_rows_) return -1;
    max_S_ = S;
  }
  // the time steps are the only host input: copied outside the (replayable) launch sequence
  FCK(cudaMemcpyAsync(ts_dev_, t_host, S * sizeof(float), cudaMemcpyHostToDevice, st));
  const GraphKey key{reinterpret_cast<uintptr_t>(x), reinterpret_cast<uintptr_t>(eps), uintptr_t(S), 0};
  if (replay(key, st)) return 0;
  // sizing pass, then the real one
  uint8_t* saved = arena_;
  dry_ = true; top_ = 0; peak_ = 0; arena_ = nullptr;
  int r = run(nullptr, nullptr, S);
  dry_ = false; arena_ = saved;
  if (r) return -1;
  if (reserve(peak_ + (size_t(1) << 20), "face UNet")) return -1;
  return run_or_capture(key, st, [&]() -> int {
    top_ = 0; launches_ = 0; flops_ = 0;
    const int wpb = 8;
    small_linear_kernel<<<(tch + wpb - 1) / wpb, wpb * 32, 0, st_>>>(ts_dev_, 1, t_w1_, t_b1_, temb_a_, tch, S, tch, cfg_.ch, 3, 1);
    small_linear_kernel<<<(tch + wpb - 1) / wpb, wpb * 32, 0, st_>>>(temb_a_, tch, t_w2_, t_b2_, temb_b_, tch, S, tch, tch, 0, 0);
    small_linear_kernel<<<(tproj_total_ + wpb - 1) / wpb, wpb * 32, 0, st_>>>(temb_b_, tch, tproj_w_, tproj_b_, temb_rows_, tproj_total_, S, tproj_total_, tch, 1, 0);
    launches_ += 3;
    if (run(x, eps, S)) return -1;
    FCK(cudaGetLastError());
    return 0;
  });
}

}  // namespace hedit
