// Engine implementation: weight ingestion (reference state-dict names -> engine layouts), plan construction for the
// SD-1.x UNet (diffusers-0.18 topology, see oracle/sd_unet.py for the restated reference), and the forward launcher.
#include "engine.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "compat_attn.cuh"
#include "elementwise.cuh"
#include "launch.h"
#include "nvtx.h"
#include "tmap.h"

namespace hedit {

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      char buf_[512];                                                                              \
      snprintf(buf_, sizeof buf_, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      err_ = buf_;                                                                                 \
      return -1;                                                                                   \
    }                                                                                              \
  } while (0)

// ------------------------------------------------------------------------------------------------ weight conversion
__global__ void cvt_rows_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int rows, int K, int geglu, int half_rows) {
  const size_t total = size_t(rows) * K;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int r = int(i / K), k = int(i % K);
    int sr = r;
    if (geglu) {  // dst rows: 32-row chunks = 16 value rows | their 16 gate rows
      const int ch = r >> 5, j = r & 31;
      sr = (j < 16) ? (ch * 16 + j) : (half_rows + ch * 16 + (j - 16));
    }
    dst[i] = to_op(src[size_t(sr) * K + k]);
  }
}
__global__ void perm_geglu_vec_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int half_rows) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int ch = r >> 5, j = r & 31;
  dst[r] = src[(j < 16) ? (ch * 16 + j) : (half_rows + ch * 16 + (j - 16))];
}
// [O][I][3][3] fp32 -> [O][3][3][I] bf16
__global__ void cvt_conv3_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int O, int I) {
  const size_t total = size_t(O) * I * 9;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int ci = int(i % I);
    const int tap = int((i / I) % 9);
    const int o = int(i / (size_t(I) * 9));
    dst[i] = to_op(src[(size_t(o) * I + ci) * 9 + tap]);
  }
}

template <typename T>
T* Engine::dalloc(size_t n) {
  void* p = nullptr;
  if (cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)) != cudaSuccess) {
    err_ = "cudaMalloc failed";
    return nullptr;
  }
  owned_.push_back(p);
  return reinterpret_cast<T*>(p);
}
void* Engine::scratch_alloc(size_t bytes) { return dalloc<uint8_t>(bytes); }

void Engine::reg(const std::string& name, WeightSlot::Kind k, void* dst, size_t off, std::vector<int64_t> shape) {
  WeightSlot s;
  s.kind = k; s.dst = dst; s.dst_off_elems = off; s.shape = std::move(shape);
  slots_[name] = s;
}

Engine::Engine(const UNetCfg& cfg, int max_samples, int max_ctx) : cfg_(cfg), maxS_(max_samples), maxCtx_(max_ctx) {
  const int temb = cfg.boc[0] * 4;
  const int G = cfg.groups;
  (void)G;
  // ---- module structure in forward order
  struct RSpec { std::string name; int cin, cout; };
  struct TSpec { std::string name; int C; int tokens; };
  std::vector<RSpec> rs; std::vector<TSpec> ts;
  int res = cfg.sample;
  int cout = cfg.boc[0];
  std::vector<int> skip_ch = {cfg.boc[0]};
  for (int i = 0; i < 4; ++i) {
    const int cin = cout; cout = cfg.boc[i];
    for (int l = 0; l < cfg.layers; ++l) {
      rs.push_back({"down_blocks." + std::to_string(i) + ".resnets." + std::to_string(l), l == 0 ? cin : cout, cout});
      if (i < 3) ts.push_back({"down_blocks." + std::to_string(i) + ".attentions." + std::to_string(l), cout, res * res});
      skip_ch.push_back(cout);
    }
    if (i < 3) { skip_ch.push_back(cout); res /= 2; }
  }
  rs.push_back({"mid_block.resnets.0", cfg.boc[3], cfg.boc[3]});
  ts.push_back({"mid_block.attentions.0", cfg.boc[3], res * res});
  rs.push_back({"mid_block.resnets.1", cfg.boc[3], cfg.boc[3]});
  int prev = cfg.boc[3];
  for (int i = 0; i < 4; ++i) {
    const int oc = cfg.boc[3 - i];
    for (int l = 0; l < cfg.layers + 1; ++l) {
      const int sk = skip_ch.back(); skip_ch.pop_back();
      const int rin = (l == 0 ? prev : oc) + sk;
      rs.push_back({"up_blocks." + std::to_string(i) + ".resnets." + std::to_string(l), rin, oc});
      if (i > 0) ts.push_back({"up_blocks." + std::to_string(i) + ".attentions." + std::to_string(l), oc, res * res});
    }
    prev = oc;
    if (i < 3) res *= 2;
  }
  // ---- allocate + register
  conv_in_w_ = dalloc<float>(size_t(cfg.boc[0]) * cfg.in_ch * 9); conv_in_b_ = dalloc<float>(cfg.boc[0]);
  reg("conv_in.weight", WeightSlot::F32_COPY, conv_in_w_, 0, {cfg.boc[0], cfg.in_ch, 3, 3});
  reg("conv_in.bias", WeightSlot::F32_COPY, conv_in_b_, 0, {cfg.boc[0]});
  t_w1_ = dalloc<float>(size_t(temb) * cfg.boc[0]); t_b1_ = dalloc<float>(temb);
  t_w2_ = dalloc<float>(size_t(temb) * temb); t_b2_ = dalloc<float>(temb);
  reg("time_embedding.linear_1.weight", WeightSlot::F32_COPY, t_w1_, 0, {temb, cfg.boc[0]});
  reg("time_embedding.linear_1.bias", WeightSlot::F32_COPY, t_b1_, 0, {temb});
  reg("time_embedding.linear_2.weight", WeightSlot::F32_COPY, t_w2_, 0, {temb, temb});
  reg("time_embedding.linear_2.bias", WeightSlot::F32_COPY, t_b2_, 0, {temb});
  tproj_total_ = 0;
  for (auto& r : rs) tproj_total_ += r.cout;
  tproj_w_ = dalloc<float>(size_t(tproj_total_) * temb); tproj_b_ = dalloc<float>(tproj_total_);
  int toff = 0;
  for (auto& r : rs) {
    ResW w; w.cin = r.cin; w.cout = r.cout; w.temb_off = toff;
    w.n1g = dalloc<float>(r.cin); w.n1b = dalloc<float>(r.cin); w.n2g = dalloc<float>(r.cout); w.n2b = dalloc<float>(r.cout);
    w.b1 = dalloc<float>(r.cout); w.b2 = dalloc<float>(r.cout);
    w.w1 = dalloc<bf16>(size_t(r.cout) * 9 * r.cin); w.w2 = dalloc<bf16>(size_t(r.cout) * 9 * r.cout);
    reg(r.name + ".norm1.weight", WeightSlot::F32_COPY, w.n1g, 0, {r.cin});
    reg(r.name + ".norm1.bias", WeightSlot::F32_COPY, w.n1b, 0, {r.cin});
    reg(r.name + ".conv1.weight", WeightSlot::BF16_CONV3, w.w1, 0, {r.cout, r.cin, 3, 3});
    reg(r.name + ".conv1.bias", WeightSlot::F32_COPY, w.b1, 0, {r.cout});
    reg(r.name + ".time_emb_proj.weight", WeightSlot::F32_COPY, tproj_w_, size_t(toff) * temb, {r.cout, temb});
    reg(r.name + ".time_emb_proj.bias", WeightSlot::F32_COPY, tproj_b_, toff, {r.cout});
    reg(r.name + ".norm2.weight", WeightSlot::F32_COPY, w.n2g, 0, {r.cout});
    reg(r.name + ".norm2.bias", WeightSlot::F32_COPY, w.n2b, 0, {r.cout});
    reg(r.name + ".conv2.weight", WeightSlot::BF16_CONV3, w.w2, 0, {r.cout, r.cout, 3, 3});
    reg(r.name + ".conv2.bias", WeightSlot::F32_COPY, w.b2, 0, {r.cout});
    if (r.cin != r.cout) {
      w.wsc = dalloc<bf16>(size_t(r.cout) * r.cin); w.bsc = dalloc<float>(r.cout);
      reg(r.name + ".conv_shortcut.weight", WeightSlot::BF16_ROWS, w.wsc, 0, {r.cout, r.cin, 1, 1});
      reg(r.name + ".conv_shortcut.bias", WeightSlot::F32_COPY, w.bsc, 0, {r.cout});
    }
    toff += r.cout;
    res_.push_back(w);
  }
  int ci = 0;
  for (auto& t : ts) {
    TfW w; w.C = t.C; w.cross_index = ci++;
    const int C = t.C, D = cfg.ctx_dim;
    auto f = [&](int n) { return dalloc<float>(n); };
    w.gng = f(C); w.gnb = f(C); w.b_in = f(C); w.b_out = f(C);
    w.ln1g = f(C); w.ln1b = f(C); w.ln2g = f(C); w.ln2b = f(C); w.ln3g = f(C); w.ln3b = f(C);
    w.b_o1 = f(C); w.b_o2 = f(C); w.b_ff1 = f(8 * C); w.b_ff2 = f(C);
    w.w_in = dalloc<bf16>(size_t(C) * C); w.w_out = dalloc<bf16>(size_t(C) * C);
    w.w_qkv = dalloc<bf16>(size_t(3) * C * C); w.w_o1 = dalloc<bf16>(size_t(C) * C);
    w.w_q2 = dalloc<bf16>(size_t(C) * C); w.w_kv2 = dalloc<bf16>(size_t(2) * C * D); w.w_o2 = dalloc<bf16>(size_t(C) * C);
    w.w_ff1 = dalloc<bf16>(size_t(8) * C * C); w.w_ff2 = dalloc<bf16>(size_t(C) * 4 * C);
    w.kv_cache = dalloc<bf16>(size_t(maxCtx_) * cfg.ctx_len * 2 * C);
    const std::string P = t.name, B = t.name + ".transformer_blocks.0";
    reg(P + ".norm.weight", WeightSlot::F32_COPY, w.gng, 0, {C});
    reg(P + ".norm.bias", WeightSlot::F32_COPY, w.gnb, 0, {C});
    reg(P + ".proj_in.weight", WeightSlot::BF16_ROWS, w.w_in, 0, {C, C, 1, 1});
    reg(P + ".proj_in.bias", WeightSlot::F32_COPY, w.b_in, 0, {C});
    reg(P + ".proj_out.weight", WeightSlot::BF16_ROWS, w.w_out, 0, {C, C, 1, 1});
    reg(P + ".proj_out.bias", WeightSlot::F32_COPY, w.b_out, 0, {C});
    reg(B + ".norm1.weight", WeightSlot::F32_COPY, w.ln1g, 0, {C}); reg(B + ".norm1.bias", WeightSlot::F32_COPY, w.ln1b, 0, {C});
    reg(B + ".norm2.weight", WeightSlot::F32_COPY, w.ln2g, 0, {C}); reg(B + ".norm2.bias", WeightSlot::F32_COPY, w.ln2b, 0, {C});
    reg(B + ".norm3.weight", WeightSlot::F32_COPY, w.ln3g, 0, {C}); reg(B + ".norm3.bias", WeightSlot::F32_COPY, w.ln3b, 0, {C});
    reg(B + ".attn1.to_q.weight", WeightSlot::BF16_ROWS, w.w_qkv, 0, {C, C});
    reg(B + ".attn1.to_k.weight", WeightSlot::BF16_ROWS, w.w_qkv, size_t(C) * C, {C, C});
    reg(B + ".attn1.to_v.weight", WeightSlot::BF16_ROWS, w.w_qkv, size_t(2) * C * C, {C, C});
    reg(B + ".attn1.to_out.0.weight", WeightSlot::BF16_ROWS, w.w_o1, 0, {C, C});
    reg(B + ".attn1.to_out.0.bias", WeightSlot::F32_COPY, w.b_o1, 0, {C});
    reg(B + ".attn2.to_q.weight", WeightSlot::BF16_ROWS, w.w_q2, 0, {C, C});
    reg(B + ".attn2.to_k.weight", WeightSlot::BF16_ROWS, w.w_kv2, 0, {C, D});
    reg(B + ".attn2.to_v.weight", WeightSlot::BF16_ROWS, w.w_kv2, size_t(C) * D, {C, D});
    reg(B + ".attn2.to_out.0.weight", WeightSlot::BF16_ROWS, w.w_o2, 0, {C, C});
    reg(B + ".attn2.to_out.0.bias", WeightSlot::F32_COPY, w.b_o2, 0, {C});
    reg(B + ".ff.net.0.proj.weight", WeightSlot::BF16_GEGLU_ROWS, w.w_ff1, 0, {8 * C, C});
    reg(B + ".ff.net.0.proj.bias", WeightSlot::F32_GEGLU_VEC, w.b_ff1, 0, {8 * C});
    reg(B + ".ff.net.2.weight", WeightSlot::BF16_ROWS, w.w_ff2, 0, {C, 4 * C});
    reg(B + ".ff.net.2.bias", WeightSlot::F32_COPY, w.b_ff2, 0, {C});
    tfs_.push_back(w);
    tf_tokens_.push_back(t.tokens);
    tf_place_.push_back(t.name.rfind("down_blocks", 0) == 0 ? 0 : t.name.rfind("mid_block", 0) == 0 ? 1 : 2);
  }
  for (int i = 0; i < 3; ++i) {
    const int c = cfg.boc[i];
    bf16* w = dalloc<bf16>(size_t(c) * 9 * c); float* b = dalloc<float>(c);
    down_w_.push_back(w); down_b_.push_back(b);
    reg("down_blocks." + std::to_string(i) + ".downsamplers.0.conv.weight", WeightSlot::BF16_CONV3, w, 0, {c, c, 3, 3});
    reg("down_blocks." + std::to_string(i) + ".downsamplers.0.conv.bias", WeightSlot::F32_COPY, b, 0, {c});
    const int cu = cfg.boc[3 - i];
    bf16* wu = dalloc<bf16>(size_t(cu) * 9 * cu); float* bu = dalloc<float>(cu);
    bf16* wp = dalloc<bf16>(size_t(4) * cu * 4 * cu);
    up_w_.push_back(wu); up_b_.push_back(bu); up_wp_.push_back(wp);
    reg("up_blocks." + std::to_string(i) + ".upsamplers.0.conv.weight", WeightSlot::BF16_CONV3, wu, 0, {cu, cu, 3, 3});
    slots_["up_blocks." + std::to_string(i) + ".upsamplers.0.conv.weight"].dst_up = wp;
    reg("up_blocks." + std::to_string(i) + ".upsamplers.0.conv.bias", WeightSlot::F32_COPY, bu, 0, {cu});
  }
  norm_out_g_ = dalloc<float>(cfg.boc[0]); norm_out_b_ = dalloc<float>(cfg.boc[0]);
  conv_out_w_ = dalloc<bf16>(size_t(cfg.out_ch) * cfg.boc[0] * 9); conv_out_b_ = dalloc<float>(cfg.out_ch);
  reg("conv_norm_out.weight", WeightSlot::F32_COPY, norm_out_g_, 0, {cfg.boc[0]});
  reg("conv_norm_out.bias", WeightSlot::F32_COPY, norm_out_b_, 0, {cfg.boc[0]});
  reg("conv_out.weight", WeightSlot::BF16_CONV3, conv_out_w_, 0, {cfg.out_ch, cfg.boc[0], 3, 3});
  reg("conv_out.bias", WeightSlot::F32_COPY, conv_out_b_, 0, {cfg.out_ch});

  temb_act_ = dalloc<float>(size_t(maxT_) * temb * 2);
  temb_table_ = dalloc<float>(size_t(maxT_) * tproj_total_);
  temb_rows_ = dalloc<float>(size_t(maxS_) * tproj_total_);
  ts_dev_ = dalloc<float>(maxT_);
  ctx_bf16_ = dalloc<bf16>(size_t(maxCtx_) * cfg.ctx_len * cfg.ctx_dim);
  size_t mx = 0;
  for (auto& kv : slots_) { size_t n = 1; for (auto d : kv.second.shape) n *= size_t(d); mx = std::max(mx, n); }
  stage_elems_ = mx;
  stage_ = dalloc<float>(mx);
  // number of LocalBlend layers: cross-attention layers of the down/up path with 16x16 tokens
  n_blend_layers_ = 0;
  for (size_t i = 0; i < ts.size(); ++i)
    if (ts[i].tokens == 256 && ts[i].name.rfind("mid_block", 0) != 0) ++n_blend_layers_;
}

Engine::~Engine() {
  drop_graphs();
  if (cap_stream_) cudaStreamDestroy(cap_stream_);
  for (auto& s : loop_pool_) if (s.first) cudaFree(s.first);
  for (void* p : owned_) cudaFree(p);
  if (arena_) cudaFree(arena_);
  if (compat_probs_) cudaFree(compat_probs_);
  if (compat_aux_) cudaFree(compat_aux_);
}

void Engine::drop_graphs() {
  for (auto& g : fwd_graphs_) if (g.exec) cudaGraphExecDestroy(g.exec);
  fwd_graphs_.clear();
}

void* Engine::loop_slot(size_t index, size_t bytes) {
  if (index >= loop_pool_.size()) loop_pool_.resize(index + 1, {nullptr, 0});
  auto& s = loop_pool_[index];
  bytes = std::max<size_t>(bytes, 16);
  if (s.second >= bytes) return s.first;
  if (s.first) { cudaDeviceSynchronize(); cudaFree(s.first); s.first = nullptr; s.second = 0; drop_graphs(); }
  if (cudaMalloc(&s.first, bytes) != cudaSuccess) { cudaGetLastError(); s.first = nullptr; err_ = "loop buffer cudaMalloc failed"; return nullptr; }
  s.second = bytes;
  return s.first;
}

int Engine::load_tensor(const char* name, const float* src, const int64_t* dims, int ndim, cudaStream_t st) {
  auto it = slots_.find(name);
  if (it == slots_.end()) { err_ = std::string("unknown tensor ") + name; return -2; }
  WeightSlot& s = it->second;
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) n *= size_t(dims[i]);
  size_t want = 1;
  for (auto d : s.shape) want *= size_t(d);
  if (n != want) { err_ = std::string("shape mismatch for ") + name; return -3; }
  CK(cudaMemcpyAsync(stage_, src, n * sizeof(float), cudaMemcpyDefault, st));
  const int threads = 256;
  const int blocks = int(std::min<size_t>((n + threads - 1) / threads, 4096));
  switch (s.kind) {
    case WeightSlot::F32_COPY:
      CK(cudaMemcpyAsync(reinterpret_cast<float*>(s.dst) + s.dst_off_elems, stage_, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
      break;
    case WeightSlot::BF16_ROWS: {
      const int rows = int(s.shape[0]); const int K = int(n / rows);
      cvt_rows_bf16_kernel<<<blocks, threads, 0, st>>>(stage_, reinterpret_cast<bf16*>(s.dst) + s.dst_off_elems, rows, K, 0, 0);
      break;
    }
    case WeightSlot::BF16_GEGLU_ROWS: {
      const int rows = int(s.shape[0]); const int K = int(n / rows);
      cvt_rows_bf16_kernel<<<blocks, threads, 0, st>>>(stage_, reinterpret_cast<bf16*>(s.dst), rows, K, 1, rows / 2);
      break;
    }
    case WeightSlot::F32_GEGLU_VEC:
      perm_geglu_vec_kernel<<<(int(n) + 255) / 256, 256, 0, st>>>(stage_, reinterpret_cast<float*>(s.dst), int(n), int(n) / 2);
      break;
    case WeightSlot::BF16_CONV3:
      cvt_conv3_bf16_kernel<<<blocks, threads, 0, st>>>(stage_, reinterpret_cast<bf16*>(s.dst), int(s.shape[0]), int(s.shape[1]));
      if (s.dst_up) cvt_upconv_phases_kernel<<<blocks, threads, 0, st>>>(stage_, reinterpret_cast<bf16*>(s.dst_up), int(s.shape[0]), int(s.shape[1]));
      break;
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));   // stage_ is reused by the next tensor
  s.loaded = true;
  return 0;
}

int Engine::finalize_weights(std::string* missing) {
  int n = 0;
  for (auto& kv : slots_)
    if (!kv.second.loaded) { if (missing && n < 8) *missing += kv.first + " "; ++n; }
  if (n) { err_ = "missing weights: " + (missing ? *missing : std::string("?")); return -n; }
  return 0;
}

// ------------------------------------------------------------------------------------------------ GEMM descriptors
static int pick_bn(int M, int N, bool narrow_ok) {
  static const int force = getenv("HEDIT_GEMM_BN") ? atoi(getenv("HEDIT_GEMM_BN")) : 0;     // tuning switch
  if (force == 160 || force == 256) return force;
  // Measured on B200: one 128 x BN x 16 MMA step costs ~ BN/2 + 90 cycles (operand fetch + TMA refill share the SM's shared-
  // memory bandwidth), so wide tiles win unless they add a wave or mostly-empty columns.
  auto cost = [&](int bn) {
    const long tiles = long((M + 127) / 128) * ((N + bn - 1) / bn);
    return double((tiles + 147) / 148) * (bn / 2 + 90);
  };
  int best = cost(256) < cost(160) ? 256 : 160;
  // narrow tiles (round 2): N = 64 / 128 layers of the face-swapping networks (DDPM UNet at 128 channels, IR-SE50, VGG16) waste
  // 20-60 % of a 160-column tile, and launches with fewer tiles than SMs get more CTAs.  Ties keep the wider tile.
  static const bool narrow_off = getenv("HEDIT_GEMM_NARROW") && atoi(getenv("HEDIT_GEMM_NARROW")) == 0;
  if (narrow_ok && !narrow_off) {
    if (cost(128) < cost(best)) best = 128;
    if (cost(64) < cost(best)) best = 64;
  }
  return best;
}

bool make_gemm(GemmParams& g, int& bn, const bf16* A, int lda, int a_mode, const ConvGeom* cg, const bf16* Wt, int M, int N,
               int Ktot, const GemmEpilogue& ep, std::string& err, int ldw) {
  memset(&g, 0, sizeof g);
  bn = pick_bn(M, N, !ep.geglu && !gemm_cluster());
  g.M = M; g.N = N; g.a_mode = a_mode; g.ep = ep;
  g.num_kb = (Ktot + 63) / 64;
  bool ok = true;
  if (a_mode == A_LINEAR) {
    uint64_t dims[2] = {uint64_t(Ktot), uint64_t(M)}; uint64_t str[1] = {uint64_t(lda) * 2}; uint32_t box[2] = {64, 128};
    ok = make_tmap_bf16(&g.tmA, A, 2, dims, str, box);
  } else {
    const int W = cg->W, H = cg->H, C = cg->C;
    const bool wide = W > 128;          // one 128-row tile = part of an image row
    if ((wide ? W % 128 != 0 : 128 % W != 0) || C % 64 != 0) {
      err = "conv geometry unsupported (need W | 128 or 128 | W, Cin % 64 == 0)";
      return false;
    }
    const int BH = wide ? 1 : std::min(H, 128 / W);
    const int BS = wide ? 1 : 128 / (W * BH);
    if (H % BH != 0) { err = "conv geometry unsupported (H)"; return false; }
    g.conv_W = W; g.conv_H = H; g.conv_cin = C; g.cin_blocks = C / 64; g.conv_pad01 = cg->pad01; g.conv_ox = cg->ox; g.conv_oy = cg->oy;
    if (a_mode == A_CONV3X3 || a_mode == A_CONV2X2) {
      uint64_t dims[4] = {uint64_t(C), uint64_t(W), uint64_t(H), uint64_t(cg->S)};
      uint64_t str[3] = {uint64_t(C) * 2, uint64_t(W) * C * 2, uint64_t(H) * W * C * 2};
      uint32_t box[4] = {64, uint32_t(std::min(W, 128)), uint32_t(BH), uint32_t(BS)};
      ok = make_tmap_bf16(&g.tmA, A, 4, dims, str, box);
    } else {
      const int Win = 2 * W, Hin = 2 * H;
      uint64_t dims[5] = {uint64_t(2 * C), uint64_t(W), 2, uint64_t(H), uint64_t(cg->S)};
      uint64_t str[4] = {uint64_t(2 * C) * 2, uint64_t(Win) * C * 2, uint64_t(2) * Win * C * 2, uint64_t(Hin) * Win * C * 2};
      uint32_t box[5] = {64, uint32_t(std::min(W, 128)), 1, uint32_t(BH), uint32_t(BS)};
      ok = make_tmap_bf16(&g.tmA, A, 5, dims, str, box);
    }
  }
  if (!ok) { err = "tensor map (A) encode failed"; return false; }
  uint64_t dimsB[2] = {uint64_t(Ktot), uint64_t(N)}; uint64_t strB[1] = {uint64_t(ldw > 0 ? ldw : Ktot) * 2}; uint32_t boxB[2] = {64, uint32_t(gemm_cluster() ? bn / 2 : bn)};
  g.b_full_box = gemm_cluster() ? 0 : 1;
  if (!make_tmap_bf16(&g.tmB, Wt, 2, dimsB, strB, boxB)) { err = "tensor map (B) encode failed"; return false; }
  // 16-bit outputs without bias / residual on 256-column tiles (q|k|v, q, text k|v projections): bulk tensor stores of the staged tile.
  // HEDIT_GEMM_TMA_STORE=0 keeps the per-lane 16-byte stores.
  static const bool tmd_on = !(getenv("HEDIT_GEMM_TMA_STORE") && atoi(getenv("HEDIT_GEMM_TMA_STORE")) == 0);
  if (tmd_on && bn == 256 && ep.out_bf16 && !ep.out_f32 && !ep.bias && !ep.rowvec && !ep.residual && !ep.geglu && (N % 64) == 0 && (ep.ldob % 8) == 0 &&
      (reinterpret_cast<uintptr_t>(ep.out_bf16) & 15) == 0) {
    uint64_t dimsD[2] = {uint64_t(N), uint64_t(M)}; uint64_t strD[1] = {uint64_t(ep.ldob) * 2}; uint32_t boxD[2] = {64, 32};
    g.use_tmd = make_tmap_bf16(&g.tmD, ep.out_bf16, 2, dimsD, strD, boxD) ? 1 : 0;
  }
  return true;
}

// ---- split-K ------------------------------------------------------------------------------------------------------------------
int gemm_splitk_splits(const GemmParams& g, int bn) {
  static const bool off = getenv("HEDIT_GEMM_SPLITK") && atoi(getenv("HEDIT_GEMM_SPLITK")) == 0;
  const GemmEpilogue& e = g.ep;
  if (off || gemm_cluster() || e.geglu || e.nchw_hw || e.up_W > 0 || e.diag_skip || (g.N & 3) || (e.ldo & 3) || (e.ldob & 3) || (e.ldr & 3)) return 1;
  const int tiles = ((g.M + 127) / 128) * ((g.N + bn - 1) / bn);
  if (tiles * 2 > 148 || g.num_kb < 16) return 1;
  int splits = std::min(std::min(8, 148 / tiles), g.num_kb / 8);
  if (splits < 2) return 1;
  const int per = (g.num_kb + splits - 1) / splits;
  splits = (g.num_kb + per - 1) / per;          // no empty split
  return splits;
}
void enable_splitk(GemmParams& g, int splits, float* ws, SplitKReduceParams& red) {
  memset(&red, 0, sizeof red);
  red.ws = ws; red.split_stride = size_t(g.M) * g.N; red.splits = splits; red.M = g.M; red.N = g.N; red.ep = g.ep;
  if (red.ep.rows_per_group == 0) red.ep.rows_per_group = 1;
  g.splits = splits; g.kb_per_split = (g.num_kb + splits - 1) / splits; g.split_stride = red.split_stride;
  GemmEpilogue pe; memset(&pe, 0, sizeof pe);
  pe.out_f32 = ws; pe.ldo = g.N; pe.rows_per_group = 1;
  g.ep = pe;
  g.use_tmd = 0;
}
cudaError_t launch_splitk_reduce(const SplitKReduceParams& red, cudaStream_t st) {
  return launch_k(splitk_reduce_kernel, dim3((red.N + 31) / 32, (red.M + 31) / 32), dim3(256), 0, st, red);
}

static int g_num_sms = 0;
static int num_sms() {
  if (!g_num_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev); if (g_num_sms <= 0) g_num_sms = 148; }
  return g_num_sms;
}

// tuning switch HEDIT_GEMM_CLUSTER=1 enables the cta_group::2 (CTA pair) variant
// (default off: measured equal to the single-CTA kernel on B200 for every UNet shape -- the kernel is not bound by W traffic)
bool gemm_cluster() { static const bool v = getenv("HEDIT_GEMM_CLUSTER") && atoi(getenv("HEDIT_GEMM_CLUSTER")) != 0; return v; }

template <int BN, bool CL, int EPI = 0>
static cudaError_t launch_gemm_t(const GemmParams& g, cudaStream_t st) {
  using Cfg = GemmCfg<BN, CL, EPI>;
  static bool attr_set = false;
  if (!attr_set) { cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, CL, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES); attr_set = true; }
  const int m_tiles = (g.M + 127) / 128, n_tiles = (g.N + BN - 1) / BN;
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(Cfg::THREADS); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES; cfg.stream = st;
  cudaLaunchAttribute at[2];
  int nat = 0;
  if (CL) {
    const int pairs = ((m_tiles + 1) / 2) * n_tiles;
    cfg.gridDim = dim3(2 * std::min(pairs, num_sms() / 2));
    at[nat].id = cudaLaunchAttributeClusterDimension; at[nat].val.clusterDim.x = 2; at[nat].val.clusterDim.y = 1; at[nat].val.clusterDim.z = 1;
    ++nat;
  } else {
    cfg.gridDim = dim3(std::min(m_tiles * n_tiles * std::max(1, g.splits), num_sms()));
  }
  if (pdl_enabled()) { at[nat].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[nat].val.programmaticStreamSerializationAllowed = 1; ++nat; }
  cfg.attrs = at; cfg.numAttrs = nat;
  return cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05_kernel<BN, CL, EPI>, g);
}

cudaError_t launch_gemm(const GemmParams& g, int bn, cudaStream_t st) {
  const bool cl = gemm_cluster() && g.M > 128;
  if (g.ep.geglu) {          // GEGLU write-back: its own instantiations
    if (bn == 256) return cl ? launch_gemm_t<256, true, 1>(g, st) : launch_gemm_t<256, false, 1>(g, st);
    return cl ? launch_gemm_t<160, true, 1>(g, st) : launch_gemm_t<160, false, 1>(g, st);
  }
  // (EPI == 2, a 16-warp write-back for the bias + fp32 residual -> fp32 projections at K <= 640, was measured at 0.112 vs 0.114 ms on the
  // 163840 x 320 x 320 case: those launches already move 4.7 TB/s, so it is not instantiated)
  if (bn == 256) return cl ? launch_gemm_t<256, true>(g, st) : launch_gemm_t<256, false>(g, st);
  if (bn == 128 && !cl) return launch_gemm_t<128, false>(g, st);
  if (bn == 64 && !cl) return launch_gemm_t<64, false>(g, st);
  return cl ? launch_gemm_t<160, true>(g, st) : launch_gemm_t<160, false>(g, st);
}

// ------------------------------------------------------------------------------------------------ attention descriptors
static bool make_attn_maps(AttnParams& a, const bf16* q, int ldq, int Nq, int Sq, const bf16* k, const bf16* v, int ldkv, int Nkv,
                           int Skv, int H, int d, int bkv, std::string& err) {
  auto mk = [&](CUtensorMap* m, const bf16* base, int ld, int N, int S, int rows) {
    uint64_t dims[4] = {uint64_t(d), uint64_t(H), uint64_t(N), uint64_t(S)};
    uint64_t str[3] = {uint64_t(d) * 2, uint64_t(ld) * 2, uint64_t(N) * ld * 2};
    uint32_t box[4] = {64, 1, uint32_t(rows), 1};
    return make_tmap_bf16(m, base, 4, dims, str, box);
  };
  if (d % 8 != 0 || d > 192) { err = "head dim must be a multiple of 8 and <= 192"; return false; }
  if (!mk(&a.tmQ, q, ldq, Nq, Sq, 128) || !mk(&a.tmK, k, ldkv, Nkv, Skv, bkv) || !mk(&a.tmV, v, ldkv, Nkv, Skv, bkv)) {
    err = "tensor map (attention) encode failed";
    return false;
  }
  return true;
}

template <int DCH, int BKV>
static cudaError_t launch_self_t(const AttnParams& a, int S, cudaStream_t st) {
  static bool set = false;
  if (!set) { cudaFuncSetAttribute(self_attn_kernel<DCH, BKV>, cudaFuncAttributeMaxDynamicSharedMemorySize, SelfAttnCfg<DCH, BKV>::SMEM_BYTES); set = true; }
  dim3 grid((a.Nq + 127) / 128, a.H, S);
  return launch_k(self_attn_kernel<DCH, BKV>, grid, dim3(192), SelfAttnCfg<DCH, BKV>::SMEM_BYTES, st, a);
}
template <int DCH, int NT, int BKV, int POLY = 0>
static cudaError_t launch_self2_t(const AttnParams& a, int S, cudaStream_t st) {
  using Cfg = SelfAttn2Cfg<DCH, NT, BKV>;
  static bool set = false;
  if (!set) { cudaFuncSetAttribute(self_attn2_kernel<DCH, NT, BKV, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES); set = true; }
  constexpr int rows = 128 * NT;
  dim3 grid((a.Nq + rows - 1) / rows, a.H, S);
  return launch_k(self_attn2_kernel<DCH, NT, BKV, POLY>, grid, dim3(Cfg::THREADS), Cfg::SMEM_BYTES, st, a);
}
template <int DCH, bool MMASUM, int POLY16>
static cudaError_t launch_self4_t(const AttnParams& a, int S, cudaStream_t st) {
  using Cfg = SelfAttn4Cfg<DCH>;
  static bool set = false;
  if (!set) { cudaFuncSetAttribute(self_attn4_kernel<DCH, MMASUM, POLY16>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES); set = true; }
  dim3 grid((a.Nq + 255) / 256, a.H, S);
  return launch_k(self_attn4_kernel<DCH, MMASUM, POLY16>, grid, dim3(Cfg::THREADS), Cfg::SMEM_BYTES, st, a);
}
// tuning switch HEDIT_ATTN_V4 (self_attn4_kernel, head dims <= 128; measured on B200 at 40 samples, N = 4096 d = 40 / N = 1024 d = 80):
//   unset = default: d <= 64: tensor-core row sum + 2 of 8 exponential pairs on the FMA pipe (1.405 ms; round-1 kernel 1.67, cuDNN SDPA 1.396),
//           d = 80: plain (0.189 ms; round-1 kernel 0.220);
//   0 = round-1 kernel (self_attn2_kernel); 1 = plain (1.571 / 0.189); 2 = FMA-pipe share only (1.489 / 0.190);
//   3 = tensor-core row sum only (1.586 / 0.197); 4 = row sum + share 2/8 (1.405 / 0.201); 5 = row sum + share 3/8 (1.470 / 0.201)
static int attn_v4() { static const int v = getenv("HEDIT_ATTN_V4") ? atoi(getenv("HEDIT_ATTN_V4")) : -1; return v; }
template <int DCH>
static cudaError_t launch_self4(const AttnParams& a, int S, cudaStream_t st) {
  switch (attn_v4()) {
    case 1: return launch_self4_t<DCH, false, 0>(a, S, st);
    case 2: return launch_self4_t<DCH, false, 2>(a, S, st);
    case 3: return launch_self4_t<DCH, true, 0>(a, S, st);
    case 4: return launch_self4_t<DCH, true, 2>(a, S, st);
    case 5: return launch_self4_t<DCH, true, 3>(a, S, st);
    default: return DCH == 1 ? launch_self4_t<DCH, true, 2>(a, S, st) : launch_self4_t<DCH, false, 0>(a, S, st);
  }
}
// tuning switch HEDIT_ATTN_POLY: how many of every 8 softmax exponentials run on the FMA pipe instead of the MUFU (0, 2 or 4)
static int attn_poly() { static const int v = getenv("HEDIT_ATTN_POLY") ? atoi(getenv("HEDIT_ATTN_POLY")) : 0; return v; }
// tuning switch HEDIT_ATTN_CFG for head dims <= 64 (measured on B200, N=4096, d=40, cycles per 128x128 block):
//   1 (default) = 2 tiles x 64-column blocks, 2 CTAs/SM: 1543;  0 = 2 tiles x 128-column blocks, 1 CTA/SM: 1645;
//   2 = 4 tiles x 64-column blocks, 1 CTA/SM: 2220
static int attn_cfg() { static const int v = getenv("HEDIT_ATTN_CFG") ? atoi(getenv("HEDIT_ATTN_CFG")) : 1; return v; }
cudaError_t launch_self_attn(const AttnParams& a, int dch, int S, cudaStream_t st) {
  const int bkv2 = (dch == 1 && attn_cfg() == 0) ? 128 : 64;
  if (a.Nq >= 256 && a.Nkv % 64 == 0 && attn_v4() != 0 && dch <= 2 && ((a.d + 15) & ~15) + 16 <= dch * 64)
    return dch == 1 ? launch_self4<1>(a, S, st) : launch_self4<2>(a, S, st);
  if (a.Nq >= 256 && a.Nkv % bkv2 == 0) {      // several query tiles per CTA
    if (dch == 1) {
      if (attn_cfg() == 1) {
        if (attn_poly() == 4) return launch_self2_t<1, 2, 64, 4>(a, S, st);
        if (attn_poly() == 2) return launch_self2_t<1, 2, 64, 2>(a, S, st);
        return launch_self2_t<1, 2, 64>(a, S, st);
      }
      if (attn_cfg() == 2) return launch_self2_t<1, 4, 64>(a, S, st);
      return launch_self2_t<1, 2, 128>(a, S, st);
    }
    if (dch == 2) return launch_self2_t<2, 2, 64>(a, S, st);
    return launch_self2_t<3, 2, 64>(a, S, st);
  }
  if (dch == 1) return launch_self_t<1, 128>(a, S, st);
  if (dch == 2) return launch_self_t<2, 128>(a, S, st);
  return launch_self_t<3, 64>(a, S, st);
}
template <int DCH>
static cudaError_t launch_cross_t(const AttnParams& a, int units, cudaStream_t st) {
  static bool set = false;
  if (!set) { cudaFuncSetAttribute(cross_attn_kernel<DCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, CrossAttnCfg<DCH>::SMEM_BYTES); set = true; }
  dim3 grid((a.Nq + 127) / 128, a.H, units);
  return launch_k(cross_attn_kernel<DCH>, grid, dim3(192), CrossAttnCfg<DCH>::SMEM_BYTES, st, a);
}
template <int DCH>
static cudaError_t launch_cross2_t(AttnParams a, int units, cudaStream_t st) {
  static bool set = false;
  if (!set) { cudaFuncSetAttribute(cross_attn2_kernel<DCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, CrossAttn2Cfg<DCH>::SMEM_BYTES); set = true; }
  const int ntiles = (a.Nq + 127) / 128;
  // enough CTAs for ~2 waves, but long enough runs of tiles per CTA to amortise the K/V load and the pipeline fill
  int groups = 1;
  while (groups < ntiles && groups * a.H * units < 2 * 148 && ntiles / (groups * 2) >= 2) groups *= 2;
  a.tiles_per_cta = (ntiles + groups - 1) / groups;
  dim3 grid((ntiles + a.tiles_per_cta - 1) / a.tiles_per_cta, a.H, units);
  return launch_k(cross_attn2_kernel<DCH>, grid, dim3(192), CrossAttn2Cfg<DCH>::SMEM_BYTES, st, a);
}
cudaError_t launch_cross_attn(const AttnParams& a, int dch, int units, cudaStream_t st) {
  static const bool force_v1 = getenv("HEDIT_CROSS_V1") != nullptr;      // tuning switch: single-tile kernel
  if (force_v1) {
    if (dch == 1) return launch_cross_t<1>(a, units, st);
    if (dch == 2) return launch_cross_t<2>(a, units, st);
    return launch_cross_t<3>(a, units, st);
  }
  if (dch == 1) return launch_cross2_t<1>(a, units, st);
  if (dch == 2) return launch_cross2_t<2>(a, units, st);
  return launch_cross_t<3>(a, units, st);
}
static int dch_for(int d) { return d <= 64 ? 1 : (d <= 128 ? 2 : 3); }
static int attn_cfg();
// K/V rows per TMA box: must match the kernel variant launch_self_attn() picks for the same (d, Nq, Nkv)
int self_attn_bkv(int d, int Nq, int Nkv) {
  const int bkv2 = (d <= 64 && attn_cfg() == 0) ? 128 : 64;
  if (Nq >= 256 && Nkv % bkv2 == 0) return bkv2;
  return d <= 128 ? 128 : 64;
}

// ------------------------------------------------------------------------------------------------ contexts / timesteps
int Engine::set_contexts(const float* ctx, int n_ctx, cudaStream_t st) {
  if (n_ctx > maxCtx_) { err_ = "too many contexts"; return -1; }
  const size_t n = size_t(n_ctx) * cfg_.ctx_len * cfg_.ctx_dim;
  if (n > stage_elems_) { err_ = "context batch exceeds staging buffer"; return -1; }
  CK(cudaMemcpyAsync(stage_, ctx, n * sizeof(float), cudaMemcpyDefault, st));
  cast_bf16_kernel<<<int(std::min<size_t>((n / 4 + 255) / 256, 2048)), 256, 0, st>>>(stage_, ctx_bf16_, n / 4);
  CK(cudaGetLastError());
  const int M = n_ctx * cfg_.ctx_len;
  for (auto& t : tfs_) {
    GemmParams g; int bn;
    GemmEpilogue ep; memset(&ep, 0, sizeof ep);
    ep.out_bf16 = t.kv_cache; ep.ldob = 2 * t.C; ep.rows_per_group = 1;
    if (!make_gemm(g, bn, ctx_bf16_, cfg_.ctx_dim, A_LINEAR, nullptr, t.w_kv2, M, 2 * t.C, cfg_.ctx_dim, ep, err_)) return -1;
    CK(launch_gemm(g, bn, st));
  }
  return 0;
}

int Engine::set_timesteps(const float* ts, int n, cudaStream_t st) {
  if (n > maxT_) { err_ = "too many timesteps in one edit (limit " + std::to_string(maxT_ - 1) + " steps)"; return -1; }
  const int temb = cfg_.boc[0] * 4;
  CK(cudaMemcpyAsync(ts_dev_, ts, n * sizeof(float), cudaMemcpyDefault, st));
  float* a1 = temb_act_; float* a2 = temb_act_ + size_t(maxT_) * temb;
  const int wpb = 8;
  small_linear_kernel<<<(temb + wpb - 1) / wpb, wpb * 32, 0, st>>>(ts_dev_, 1, t_w1_, t_b1_, a1, temb, n, temb, cfg_.boc[0], 2, 1);
  small_linear_kernel<<<(temb + wpb - 1) / wpb, wpb * 32, 0, st>>>(a1, temb, t_w2_, t_b2_, a2, temb, n, temb, temb, 0, 0);
  small_linear_kernel<<<(tproj_total_ + wpb - 1) / wpb, wpb * 32, 0, st>>>(a2, temb, tproj_w_, tproj_b_, temb_table_, tproj_total_, n,
                                                                           tproj_total_, temb, 1, 0);
  CK(cudaGetLastError());
  nT_ = n;
  return 0;
}

// ------------------------------------------------------------------------------------------------ plan builder
struct Arena {
  struct Blk { size_t off, size; bool free; };
  std::vector<Blk> blks;
  size_t peak = 0;
  Arena() { blks.push_back({0, size_t(1) << 62, true}); }
  size_t alloc(size_t bytes) {
    bytes = (bytes + 1023) & ~size_t(1023);
    for (size_t i = 0; i < blks.size(); ++i)
      if (blks[i].free && blks[i].size >= bytes) {
        const size_t off = blks[i].off;
        if (blks[i].size > bytes) {
          Blk rest{off + bytes, blks[i].size - bytes, true};
          blks[i].size = bytes;
          blks.insert(blks.begin() + i + 1, rest);
        }
        blks[i].free = false;
        peak = std::max(peak, off + bytes);
        return off;
      }
    return size_t(-1);
  }
  void release(size_t off) {
    for (size_t i = 0; i < blks.size(); ++i)
      if (blks[i].off == off && !blks[i].free) {
        blks[i].free = true;
        if (i + 1 < blks.size() && blks[i + 1].free) { blks[i].size += blks[i + 1].size; blks.erase(blks.begin() + i + 1); }
        if (i > 0 && blks[i - 1].free) { blks[i - 1].size += blks[i].size; blks.erase(blks.begin() + i); }
        return;
      }
  }
};

struct PlanBuilder {
  Engine& E;
  Plan* plan;         // null in the sizing pass
  int S;
  int U = 0;          // > 0: the context-free prefix is built for U distinct latents and broadcast to the S samples (CallCtrl::n_uniq)
  Arena ar;
  bool failed = false;
  double flops = 0;
  float* x_in = nullptr;    // patched per call
  float* eps_out = nullptr;

  PlanBuilder(Engine& e, Plan* p, int s) : E(e), plan(p), S(s) {}
  template <typename T> T* A(size_t n) {
    const size_t off = ar.alloc(n * sizeof(T));
    return reinterpret_cast<T*>(E.arena_ + off);     // arena_ may be null in the sizing pass (pointer arithmetic only)
  }
  template <typename T> void F(T* p) {
    ar.release(size_t(reinterpret_cast<uint8_t*>(p) - E.arena_));
    auto it = colstats_of.find(reinterpret_cast<const void*>(p));
    if (it != colstats_of.end()) { ar.release(size_t(reinterpret_cast<uint8_t*>(it->second) - E.arena_)); colstats_of.erase(it); }
  }
  void push(Op op) { op.nS = S; if (plan) plan->ops.push_back(op); }
  // GroupNorm statistics fused into the producer: fp32 tensors whose every 32-row block lies inside one sample get a
  // [M/32][N] (sum, sumsq) side buffer written by the GEMM epilogue; it lives exactly as long as the tensor.
  std::map<const void*, float2*> colstats_of;
  float2* want_colstats(const float* out, int M, int N, int HW) {
    static const bool off = getenv("HEDIT_GN_FUSED") && atoi(getenv("HEDIT_GN_FUSED")) == 0;      // tuning switch: separate gn_stats pass
    if (off || (HW & 31) || (N & 31)) return nullptr;
    float2* cs = A<float2>(size_t((M + 31) / 32) * N);
    colstats_of[out] = cs;
    return cs;
  }

  void gemm(const char* tag, const bf16* Ain, int lda, int mode, const ConvGeom* cg, const bf16* Wt, int M, int N, int K, GemmEpilogue ep) {
    flops += 2.0 * M * N * K;
    if (!plan) return;
    Op op{}; op.kind = OP_GEMM; op.tag = tag;
    if (ep.rows_per_group == 0) ep.rows_per_group = 1;
    if (!make_gemm(op.gemm, op.gemm_bn, Ain, lda, mode, cg, Wt, M, N, K, ep, E.err_)) { failed = true; return; }
    const int splits = E.splitk() ? gemm_splitk_splits(op.gemm, op.gemm_bn) : 1;
    if (splits > 1) {
      float* ws = A<float>(size_t(splits) * M * N);
      Op red{}; red.kind = OP_SPLITK_REDUCE; red.tag = tag;
      enable_splitk(op.gemm, splits, ws, red.red);
      push(op);
      push(red);
      F(ws);
      return;
    }
    push(op);
  }
  void gn(const float* x1, int C1, const float* x2, int C2, int HW, const float* g, const float* b, float eps, int silu, bf16* out, bf16* raw) {
    auto c1 = colstats_of.find(x1), c2 = colstats_of.find(x2);
    if (c1 != colstats_of.end() && (x2 == nullptr || c2 != colstats_of.end())) {
      float2* stats = A<float2>(size_t(S) * E.cfg_.groups);
      Op f{}; f.kind = OP_GN_FINALIZE; f.cs1 = c1->second; f.cs2 = x2 ? c2->second : nullptr; f.C1 = C1; f.C2 = C2; f.HW = HW; f.eps = eps;
      f.stats = stats; f.tag = "gn_finalize";
      push(f);
      Op a{}; a.kind = OP_GN_APPLY; a.f_in = x1; a.f_in2 = x2; a.C1 = C1; a.C2 = C2; a.HW = HW; a.gamma = g; a.beta = b; a.eps = eps; a.silu = silu;
      a.h_out = out; a.h_out2 = raw; a.stats = stats; a.tag = "gn_apply";
      push(a);
      F(stats);
      return;
    }
    const int chunk = std::max(16, HW / 64);
    const int nch = (HW + chunk - 1) / chunk;
    float2* partial = A<float2>(size_t(S) * nch * E.cfg_.groups);
    Op s{}; s.kind = OP_GN_STATS; s.f_in = x1; s.f_in2 = x2; s.C1 = C1; s.C2 = C2; s.HW = HW; s.chunk = chunk; s.nchunks = nch;
    s.partial = partial; s.tag = "gn_stats";
    push(s);
    Op a = s; a.kind = OP_GN_APPLY; a.gamma = g; a.beta = b; a.eps = eps; a.silu = silu; a.h_out = out; a.h_out2 = raw; a.tag = "gn_apply";
    push(a);
    F(partial);
  }
  void ln(const float* x, const float* g, const float* b, bf16* out, int rows, int C) {
    Op o{}; o.kind = OP_LN; o.f_in = x; o.gamma = g; o.beta = b; o.h_out = out; o.rows = rows; o.C1 = C; o.eps = 1e-5f; o.tag = "layernorm";
    push(o);
  }

  // ResnetBlock2D (oracle/sd_unet.py ResnetBlock2D): returns a new fp32 [S*HW][cout] buffer
  float* resblock(const ResW& w, const float* x1, int C1, const float* x2, int C2, int Hh, int Ww, bool pnp_site = false) {
    const int HW = Hh * Ww, M = S * HW, cin = C1 + C2, cout = w.cout;
    ConvGeom cg{S, Hh, Ww, cin, 1};
    bf16* a1 = A<bf16>(size_t(M) * cin);
    bf16* raw = w.wsc ? A<bf16>(size_t(M) * cin) : nullptr;
    gn(x1, C1, x2, C2, HW, w.n1g, w.n1b, 1e-5f, 1, a1, raw);
    float* h1 = A<float>(size_t(M) * cout);
    GemmEpilogue e1; memset(&e1, 0, sizeof e1);
    e1.bias = w.b1; e1.rowvec = E.temb_rows_ + w.temb_off; e1.ldrv = E.tproj_total_; e1.rows_per_group = HW; e1.out_f32 = h1; e1.ldo = cout;
    e1.colstats = want_colstats(h1, M, cout, HW);
    gemm("res.conv1", a1, cin, A_CONV3X3, &cg, w.w1, M, cout, 9 * cin, e1);
    F(a1);
    bf16* a2 = A<bf16>(size_t(M) * cout);
    gn(h1, cout, nullptr, 0, HW, w.n2g, w.n2b, 1e-5f, 1, a2, nullptr);
    F(h1);
    if (pnp_site) { Op o{}; o.kind = OP_FEAT_COPY; o.h_out = a2; o.count = size_t(HW) * cout * sizeof(bf16) / 16; o.tag = "pnp_feat_copy"; push(o); }
    const float* resid = x1;
    float* sc = nullptr;
    if (w.wsc) {
      sc = A<float>(size_t(M) * cout);
      GemmEpilogue es; memset(&es, 0, sizeof es);
      es.bias = w.bsc; es.out_f32 = sc; es.ldo = cout;
      gemm("res.shortcut", raw, cin, A_LINEAR, nullptr, w.wsc, M, cout, cin, es);
      F(raw);
      resid = sc;
    }
    float* out = A<float>(size_t(M) * cout);
    ConvGeom cg2{S, Hh, Ww, cout, 1};
    GemmEpilogue e2; memset(&e2, 0, sizeof e2);
    e2.bias = w.b2; e2.residual = resid; e2.ldr = cout; e2.out_f32 = out; e2.ldo = cout;
    e2.colstats = want_colstats(out, M, cout, HW);
    gemm("res.conv2", a2, cout, A_CONV3X3, &cg2, w.w2, M, cout, 9 * cout, e2);
    F(a2);
    if (sc) F(sc);
    return out;
  }

  // samples broadcast: dst[s] = src[uniq_of[s]] (fp32 [S][HW][C])
  float* expand(const float* src, int HW, int C) {
    float* dst = A<float>(size_t(S) * HW * C);
    Op o{}; o.kind = OP_EXPAND; o.f_in = src; o.f_out = dst; o.count = size_t(HW) * C * sizeof(float) / 16; o.tag = "prefix_expand";
    push(o);
    return dst;
  }

  // Transformer2DModel with one BasicTransformerBlock: returns a new fp32 buffer.  The context-free head (GroupNorm, proj_in, self-attention
  // incl. its residual add) can be built separately (`transformer_head`) and its result passed in as `t_in`.
  float* transformer_head(int ti, const float* x, int Hh, int Ww) {
    const TfW& w = E.tfs_[ti];
    const int C = w.C, HW = Hh * Ww, M = S * HW, H = E.cfg_.heads, d = C / H;
    bf16* a = A<bf16>(size_t(M) * C);
    gn(x, C, nullptr, 0, HW, w.gng, w.gnb, 1e-6f, 0, a, nullptr);
    float* t = A<float>(size_t(M) * C);
    GemmEpilogue e; memset(&e, 0, sizeof e);
    e.bias = w.b_in; e.out_f32 = t; e.ldo = C;
    gemm("tf.proj_in", a, C, A_LINEAR, nullptr, w.w_in, M, C, C, e);
    // ---- self-attention
    ln(t, w.ln1g, w.ln1b, a, M, C);
    bf16* qkv = A<bf16>(size_t(M) * 3 * C);
    memset(&e, 0, sizeof e); e.out_bf16 = qkv; e.ldob = 3 * C;
    gemm("tf.qkv", a, C, A_LINEAR, nullptr, w.w_qkv, M, 3 * C, C, e);
    bf16* att = A<bf16>(size_t(M) * C);
    flops += 4.0 * S * double(HW) * HW * C;
    if (plan) {
      Op o{}; o.kind = OP_SELF_ATTN; o.tag = "self_attn"; o.tf_index = ti; o.dch = dch_for(d); o.bkv = self_attn_bkv(d, HW, HW);
      AttnParams& p = o.attn;
      if (!make_attn_maps(p, qkv, 3 * C, HW, S, qkv + C, qkv + 2 * C, 3 * C, HW, S, H, d, o.bkv, E.err_)) failed = true;
      p.H = H; p.d = d; p.Nq = HW; p.Nkv = HW; p.scale_log2 = float(1.4426950408889634 / std::sqrt(double(d)));
      p.out = att; p.ldo = C;
      o.cq = qkv; o.ck = qkv + C; o.cv = qkv + 2 * C; o.c_ldq = 3 * C; o.c_ldkv = 3 * C;
      push(o);
    }
    F(qkv);
    memset(&e, 0, sizeof e); e.bias = w.b_o1; e.residual = t; e.ldr = C; e.out_f32 = t; e.ldo = C;
    gemm("tf.attn1.out", att, C, A_LINEAR, nullptr, w.w_o1, M, C, C, e);
    F(att);
    F(a);
    return t;
  }
  float* transformer(int ti, const float* x, int Hh, int Ww, float* t_in = nullptr) {
    float* t = t_in ? t_in : transformer_head(ti, x, Hh, Ww);
    const TfW& w = E.tfs_[ti];
    const int C = w.C, HW = Hh * Ww, M = S * HW, H = E.cfg_.heads, d = C / H;
    GemmEpilogue e;
    bf16* a = A<bf16>(size_t(M) * C);
    bf16* att = A<bf16>(size_t(M) * C);
    // ---- cross-attention
    ln(t, w.ln2g, w.ln2b, a, M, C);
    bf16* qc = A<bf16>(size_t(M) * C);
    memset(&e, 0, sizeof e); e.out_bf16 = qc; e.ldob = C;
    gemm("tf.q2", a, C, A_LINEAR, nullptr, w.w_q2, M, C, C, e);
    flops += 4.0 * S * double(HW) * E.cfg_.ctx_len * C;
    if (plan) {
      Op o{}; o.kind = OP_CROSS_ATTN; o.tag = "cross_attn"; o.tf_index = ti; o.dch = dch_for(d);
      o.blend_layer = -1;
      if (HW == 256 && E.cfg_.sample == 64) {   // LocalBlend layers: 16x16 maps of the down/up path
        int idx = 0;
        for (int j = 0; j < ti; ++j) if (E.tf_tokens_[j] == 256) ++idx;
        // the mid block never has 256 tokens for sample 64 (8x8), so every 256-token layer counts
        o.blend_layer = idx;
      }
      AttnParams& p = o.attn;
      if (!make_attn_maps(p, qc, C, HW, S, w.kv_cache, w.kv_cache + C, 2 * C, E.cfg_.ctx_len, E.maxCtx_, H, d, 80, E.err_)) failed = true;
      p.H = H; p.d = d; p.Nq = HW; p.Nkv = E.cfg_.ctx_len; p.scale_log2 = float(1.4426950408889634 / std::sqrt(double(d)));
      p.out = att; p.ldo = C;
      o.cq = qc; o.ck = w.kv_cache; o.cv = w.kv_cache + C; o.c_ldq = C; o.c_ldkv = 2 * C;
      push(o);
    }
    F(qc);
    memset(&e, 0, sizeof e); e.bias = w.b_o2; e.residual = t; e.ldr = C; e.out_f32 = t; e.ldo = C;
    gemm("tf.attn2.out", att, C, A_LINEAR, nullptr, w.w_o2, M, C, C, e);
    F(att);
    // ---- GEGLU feed-forward
    ln(t, w.ln3g, w.ln3b, a, M, C);
    bf16* ff = A<bf16>(size_t(M) * 4 * C);
    memset(&e, 0, sizeof e); e.bias = w.b_ff1; e.out_bf16 = ff; e.ldob = 4 * C; e.geglu = 1;
    gemm("tf.ff1_geglu", a, C, A_LINEAR, nullptr, w.w_ff1, M, 8 * C, C, e);
    memset(&e, 0, sizeof e); e.bias = w.b_ff2; e.residual = t; e.ldr = C; e.out_bf16 = a; e.ldob = C;
    gemm("tf.ff2", ff, 4 * C, A_LINEAR, nullptr, w.w_ff2, M, C, 4 * C, e);
    F(ff);
    F(t);
    float* out = A<float>(size_t(M) * C);
    memset(&e, 0, sizeof e); e.bias = w.b_out; e.residual = x; e.ldr = C; e.out_f32 = out; e.ldo = C;
    e.colstats = want_colstats(out, M, C, HW);
    gemm("tf.proj_out", a, C, A_LINEAR, nullptr, w.w_out, M, C, C, e);
    F(a);
    return out;
  }

  void build() {
    const UNetCfg& c = E.cfg_;
    int Hh = c.sample, Ww = c.sample;
    x_in = A<float>(size_t(S) * c.in_ch * Hh * Ww);
    eps_out = A<float>(size_t(S) * c.out_ch * Hh * Ww);
    std::vector<std::pair<float*, int>> skips;
    const bool dedup = U > 0 && U < S;
    const int Sfull = S;
    if (dedup) S = U;               // the context-free prefix is built for the distinct latents only
    float* x = A<float>(size_t(S) * Hh * Ww * c.boc[0]);
    float* x_in_u = dedup ? A<float>(size_t(S) * c.in_ch * Hh * Ww) : x_in;
    { Op o{}; o.kind = OP_CONV_IN; o.f_in = x_in_u; o.f_out = x; o.H = Hh; o.W = Ww; o.C1 = c.boc[0]; o.rows = dedup ? 1 : 0; o.tag = "conv_in"; push(o); }
    if (!dedup) skips.push_back({x, c.boc[0]});
    int ri = 0, ti = 0, C = c.boc[0];
    for (int i = 0; i < 4; ++i) {
      for (int l = 0; l < c.layers; ++l) {
        float* y = resblock(E.res_[ri++], x, C, nullptr, 0, Hh, Ww);
        C = c.boc[i];
        if (dedup && i == 0 && l == 0) {
          // prefix done on U samples: resnets[0] output y and the transformer state t after its self-attention; broadcast both (and the
          // conv_in output, a skip connection) to the S samples and continue per sample from the first cross-attention on
          float* t_u = transformer_head(ti, y, Hh, Ww);
          S = Sfull;
          float* x_s = expand(x, Hh * Ww, c.boc[0]); F(x); F(x_in_u);
          float* y_s = expand(y, Hh * Ww, C); F(y);
          float* t_s = expand(t_u, Hh * Ww, C); F(t_u);
          skips.push_back({x_s, c.boc[0]});
          float* z = transformer(ti++, y_s, Hh, Ww, t_s); F(y_s);
          x = z;
          skips.push_back({x, C});
          continue;
        }
        if (i < 3) { float* z = transformer(ti++, y, Hh, Ww); F(y); y = z; }
        x = y;
        skips.push_back({x, C});
      }
      if (i < 3) {
        const int M0 = S * Hh * Ww;
        bf16* xb = A<bf16>(size_t(M0) * C);
        { Op o{}; o.kind = OP_CAST; o.f_in = x; o.h_out = xb; o.count = size_t(M0) * C / 4; o.tag = "cast_bf16"; push(o); }
        Hh /= 2; Ww /= 2;
        float* y = A<float>(size_t(S) * Hh * Ww * C);
        ConvGeom cg{S, Hh, Ww, C, 2};
        GemmEpilogue e; memset(&e, 0, sizeof e); e.bias = E.down_b_[i]; e.out_f32 = y; e.ldo = C;
        e.colstats = want_colstats(y, S * Hh * Ww, C, Hh * Ww);
        gemm("downsample", xb, C, A_CONV3X3S2, &cg, E.down_w_[i], S * Hh * Ww, C, 9 * C, e);
        F(xb);
        x = y;
        skips.push_back({x, C});
      }
    }
    // mid
    { float* y = resblock(E.res_[ri++], x, C, nullptr, 0, Hh, Ww); float* z = transformer(ti++, y, Hh, Ww); F(y);
      float* u = resblock(E.res_[ri++], z, C, nullptr, 0, Hh, Ww); F(z); x = u; }
    bool x_owned = true;
    for (int i = 0; i < 4; ++i) {
      const int oc = c.boc[3 - i];
      for (int l = 0; l < c.layers + 1; ++l) {
        auto sk = skips.back(); skips.pop_back();
        float* y = resblock(E.res_[ri++], x, C, sk.first, sk.second, Hh, Ww, /*pnp_site=*/i == 1 && l == 1);
        if (x_owned) F(x);
        F(sk.first);
        C = oc;
        if (i > 0) { float* z = transformer(ti++, y, Hh, Ww); F(y); y = z; }
        x = y; x_owned = true;
      }
      if (i < 3) {
        static const bool fused_up = !(getenv("HEDIT_UPCONV_FUSED") && atoi(getenv("HEDIT_UPCONV_FUSED")) == 0);   // tuning switch
        if (fused_up && (Hh * Ww) % 32 == 0) {
          // nearest-2x upsample -> 3x3 conv as four 2x2 convs on the COARSE grid (one per output phase, weights pre-summed at load):
          // 2.25x fewer flops and no upsampled operand; every phase GEMM scatters its rows to the fine grid and files its GroupNorm
          // column statistics under [sample][phase].
          const int M0 = S * Hh * Ww;
          bf16* xb = A<bf16>(size_t(M0) * C);
          { Op o{}; o.kind = OP_CAST; o.f_in = x; o.h_out = xb; o.count = size_t(M0) * C / 4; o.tag = "cast_bf16"; push(o); }
          F(x);
          float* y = A<float>(size_t(S) * 4 * Hh * Ww * C);
          float2* cs = want_colstats(y, 4 * M0, C, 4 * Hh * Ww);
          for (int ph = 0; ph < 4; ++ph) {
            ConvGeom cg{S, Hh, Ww, C, 1, 0, (ph & 1) - 1, (ph >> 1) - 1};
            GemmEpilogue e; memset(&e, 0, sizeof e);
            e.bias = E.up_b_[i]; e.out_f32 = y; e.ldo = C; e.colstats = cs;
            e.up_W = Ww; e.up_H = Hh; e.up_py = ph >> 1; e.up_px = ph & 1;
            gemm("upsample.conv", xb, C, A_CONV2X2, &cg, E.up_wp_[i] + size_t(ph) * C * 4 * C, M0, C, 4 * C, e);
          }
          F(xb);
          Hh *= 2; Ww *= 2;
          x = y;
        } else {
        bf16* up = A<bf16>(size_t(S) * 4 * Hh * Ww * C);
        { Op o{}; o.kind = OP_UPSAMPLE; o.f_in = x; o.h_out = up; o.H = Hh; o.W = Ww; o.C1 = C; o.tag = "upsample2x"; push(o); }
        F(x);
        Hh *= 2; Ww *= 2;
        float* y = A<float>(size_t(S) * Hh * Ww * C);
        ConvGeom cg{S, Hh, Ww, C, 1};
        GemmEpilogue e; memset(&e, 0, sizeof e); e.bias = E.up_b_[i]; e.out_f32 = y; e.ldo = C;
        e.colstats = want_colstats(y, S * Hh * Ww, C, Hh * Ww);
        gemm("upsample.conv", up, C, A_CONV3X3, &cg, E.up_w_[i], S * Hh * Ww, C, 9 * C, e);
        F(up);
        x = y;
        }
      }
    }
    bf16* fin = A<bf16>(size_t(S) * Hh * Ww * C);
    gn(x, C, nullptr, 0, Hh * Ww, E.norm_out_g_, E.norm_out_b_, 1e-5f, 1, fin, nullptr);
    {   // conv_out (C0 -> 4) on the tensor-core conv path; the epilogue writes the NCHW latent layout directly
      ConvGeom cg{S, Hh, Ww, C, 1};
      GemmEpilogue e{}; e.bias = E.conv_out_b_; e.out_f32 = eps_out; e.ldo = c.out_ch; e.nchw_hw = Hh * Ww;
      gemm("conv_out", fin, C, A_CONV3X3, &cg, E.conv_out_w_, S * Hh * Ww, c.out_ch, 9 * C, e);
      Op o{}; o.kind = OP_CONV_OUT; o.f_out = eps_out; o.tag = "copy_out"; push(o);
    }
    flops += 2.0 * S * Hh * Ww * 36.0 * c.boc[0];
  }
};

Plan* Engine::get_plan(int S, int U) {
  if (U >= S || U < 0 || cfg_.layers < 1) U = 0;
  const int key = (S * 4096 + U) * 2 + (splitk_ ? 1 : 0);
  auto it = plans_.find(key);
  if (it != plans_.end()) return it->second.get();
  if (S > maxS_) { err_ = "batch exceeds max_samples"; return nullptr; }
  if (!arena_) {   // sizing pass at the maximum batch
    PlanBuilder sz(*this, nullptr, maxS_);
    sz.build();
    size_t peak = sz.ar.peak;
    for (int u : {maxS_ - 1, (maxS_ + 1) / 2}) {        // plans with a de-duplicated prefix hold distinct-latent and per-sample buffers at once
      if (u < 1 || u >= maxS_) continue;
      PlanBuilder szu(*this, nullptr, maxS_);
      szu.U = u;
      szu.build();
      peak = std::max(peak, szu.ar.peak);
    }
    arena_bytes_ = peak + (size_t(1) << 20);
    flops_per_sample_ = sz.flops / maxS_;
    if (cudaMalloc(&arena_, arena_bytes_) != cudaSuccess) { err_ = "arena cudaMalloc failed"; arena_ = nullptr; return nullptr; }
  }
  std::unique_ptr<Plan> p(new Plan());
  p->S = S;
  PlanBuilder b(*this, p.get(), S);
  b.U = U;
  b.build();
  if (b.failed || b.ar.peak > arena_bytes_) { if (err_.empty()) err_ = "plan build failed"; return nullptr; }
  Plan* raw = p.get();
  plans_[key] = std::move(p);
  return raw;
}

// ------------------------------------------------------------------------------------------------ forward
long Engine::launch_op(Op& op, int S_call, const float* x, float* eps, const CallCtrl& cc, cudaStream_t st) {
  const UNetCfg& c = cfg_;
  const size_t lat = size_t(c.in_ch) * c.sample * c.sample;
  const int S = op.nS > 0 ? op.nS : S_call;          // ops of the de-duplicated prefix run on the distinct latents only
  switch (op.kind) {
    case OP_EXPAND:
      CK(launch_k(gather_samples_kernel, dim3(unsigned(std::min<size_t>((op.count + 255) / 256, 64)), S), dim3(256), 0, st,
                  reinterpret_cast<const uint4*>(op.f_in), cc.uniq_of, reinterpret_cast<uint4*>(op.f_out), op.count));
      break;
    case OP_CONV_IN: {
      if (op.rows == 1)    // de-duplicated prefix: one latent per distinct value
        CK(launch_k(gather_samples_kernel, dim3(unsigned(std::min<size_t>((lat * 4 / 16 + 255) / 256, 64)), S), dim3(256), 0, st,
                    reinterpret_cast<const uint4*>(x), cc.uniq_first, reinterpret_cast<uint4*>(const_cast<float*>(op.f_in)), lat * sizeof(float) / 16));
      else
        CK(cudaMemcpyAsync(const_cast<float*>(op.f_in), x, S * lat * sizeof(float), cudaMemcpyDeviceToDevice, st));
      const size_t sm = (36 * size_t(op.C1) + 4 * (kConvInRows + 2) * (op.W + 2)) * sizeof(float);
      static bool set = false;
      if (!set) { cudaFuncSetAttribute(conv_in_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); set = true; }
      CK(launch_k(conv_in_kernel, dim3((op.H + kConvInRows - 1) / kConvInRows, S), dim3(256), sm, st, op.f_in, conv_in_w_, conv_in_b_, op.f_out, op.H, op.W, op.C1));
      break;
    }
    case OP_GN_STATS: {
      GNStatsParams p{op.f_in, op.f_in2, op.C1, op.C2, op.HW, c.groups, op.chunk, op.partial};
      const int quads = (op.C1 + op.C2) / 4;
      const int threads = std::min(640, ((quads + 31) / 32) * 32);
      CK(launch_k(gn_stats_kernel, dim3(op.nchunks, S), dim3(threads), 0, st, p));
      break;
    }
    case OP_GN_FINALIZE: {
      GNFinalizeParams p{op.cs1, op.cs2, op.C1, op.C2, op.HW, c.groups, op.eps, op.stats};
      CK(launch_k(gn_finalize_kernel, dim3(c.groups, S), dim3(op.HW >= 16384 ? 512 : 128), 0, st, p));
      break;
    }
    case OP_GN_APPLY: {
      const int C = op.C1 + op.C2;
      static const int gn_chunk = getenv("HEDIT_GN_CHUNK") ? atoi(getenv("HEDIT_GN_CHUNK")) : 0;      // tuning switch
      const int chunk = gn_chunk > 0 ? gn_chunk : (op.HW >= 4096 ? 32 : 16);
      const int quads_ = C / 4;
      const int threads = std::max(256, quads_ * std::max(1, (256 + quads_ - 1) / quads_));      // quads * nsub (<= 640)
      GNApplyParams p{op.f_in, op.f_in2, op.C1, op.C2, op.HW, c.groups, chunk, op.nchunks, op.partial, op.gamma, op.beta, op.eps, op.silu, op.h_out, op.h_out2, op.stats};
      CK(launch_k(gn_apply_kernel, dim3((op.HW + chunk - 1) / chunk, S), dim3(threads), 0, st, p));
      break;
    }
    case OP_GEMM:
      CK(launch_gemm(op.gemm, op.gemm_bn, st));
      break;
    case OP_SPLITK_REDUCE:
      CK(launch_splitk_reduce(op.red, st));
      break;
    case OP_LN:
      launch_layernorm(op.f_in, op.gamma, op.beta, op.h_out, op.rows, op.C1, op.eps, st);
      break;
    case OP_SELF_ATTN: {
      if (cc.probs_cb || cc.editor_cb) return compat_attention(op, false, S, cc, st);
      AttnParams a = op.attn;
      if ((cc.self_mask & (1u << op.tf_index)) && S == S_call) { a.q_idx = cc.self_q; a.k_idx = cc.self_k; a.v_idx = cc.self_v; }
      CK(launch_self_attn(a, op.dch, S, st));
      break;
    }
    case OP_CROSS_ATTN: {
      if (cc.probs_cb || cc.editor_cb) return compat_attention(op, true, S, cc, st);
      AttnParams a = op.attn;
      a.unit_s0 = cc.unit_s0; a.unit_s1 = cc.unit_s1; a.unit_img = cc.unit_img; a.ctx_idx = cc.ctx_idx;
      a.mapper = cc.mapper; a.c_base = cc.c_base; a.c_tar = cc.c_tar; a.replace_m = cc.replace_m; a.is_replace = cc.is_replace;
      a.map_w = cc.map_w; a.map_rows = cc.map_rows;
      a.blend_alpha = cc.blend_alpha; a.n_blend_layers = n_blend_layers_; a.blend_rows = cc.blend_rows;
      a.blend_layer = op.blend_layer;
      a.blend_acc = (op.blend_layer >= 0) ? cc.blend_acc : nullptr;
      CK(launch_cross_attn(a, op.dch, cc.n_units, st));
      break;
    }
    case OP_UPSAMPLE: {
      const size_t total = size_t(S) * 4 * op.H * op.W * (op.C1 / 4);
      CK(launch_k(upsample2x_bf16_kernel, dim3(unsigned(std::min<size_t>((total + 255) / 256, 8192))), dim3(256), 0, st, op.f_in, op.h_out, S, op.H, op.W, op.C1));
      break;
    }
    case OP_CAST:
      CK(launch_k(cast_bf16_kernel, dim3(unsigned(std::min<size_t>((op.count + 255) / 256, 8192))), dim3(256), 0, st, op.f_in, op.h_out, op.count));
      break;
    case OP_CONV_OUT:
      CK(cudaMemcpyAsync(eps, op.f_out, S_call * lat * sizeof(float), cudaMemcpyDeviceToDevice, st));
      break;
    case OP_FEAT_COPY:
      if (!cc.feat_src) return 0;
      CK(launch_k(copy_samples_kernel, dim3(unsigned(std::min<size_t>((op.count + 255) / 256, 64)), S), dim3(256), 0, st, op.h_out, cc.feat_src, op.count));
      break;
  }
  return 1;
}

// One attention layer the reference's way (p2p/ptp_utils.py:88-107): scores -> softmax (fp32, materialised) -> user hook -> P.V
long Engine::compat_attention(const Op& op, bool is_cross, int S, const CallCtrl& cc, cudaStream_t st) {
  const AttnParams& a = op.attn;
  const size_t need = size_t(S) * a.H * a.Nq * a.Nkv * sizeof(float);
  if (need > compat_probs_bytes_) {
    if (compat_probs_) { CK(cudaStreamSynchronize(st)); cudaFree(compat_probs_); compat_probs_ = nullptr; compat_probs_bytes_ = 0; }
    if (cudaMalloc(&compat_probs_, need) != cudaSuccess) { cudaGetLastError(); err_ = "compat attention: cudaMalloc of the probabilities buffer failed (" + std::to_string(need >> 20) + " MiB)"; return -1; }
    compat_probs_bytes_ = need;
  }
  if (a.d > 160) { err_ = "compat attention: head dim > 160"; return -1; }
  CompatAttnParams p;
  p.q = op.cq; p.ldq = op.c_ldq; p.q_sample = size_t(a.Nq) * op.c_ldq;
  p.k = op.ck; p.v = op.cv; p.ldkv = op.c_ldkv; p.kv_sample = size_t(a.Nkv) * op.c_ldkv;
  p.kv_idx = is_cross ? cc.ctx_idx : nullptr;
  p.probs = compat_probs_; p.out = a.out; p.ldo = a.ldo;
  p.H = a.H; p.d = a.d; p.Nq = a.Nq; p.Nkv = a.Nkv;
  p.scale = float(1.0 / std::sqrt(double(a.d)));
  const int BH = S * a.H;
  const size_t rows = size_t(BH) * a.Nq;
  if (cc.editor_cb) {
    // editor protocol: q, k, v split by head in fp32, scaled scores AND probabilities materialised, output returned by the hook
    const size_t n_sim = size_t(BH) * a.Nq * a.Nkv, n_q = size_t(BH) * a.Nq * a.d, n_kv = size_t(BH) * a.Nkv * a.d, n_out = size_t(S) * a.Nq * a.H * a.d;
    const size_t aux = (n_sim + n_q + 2 * n_kv + n_out) * sizeof(float);
    if (aux > compat_aux_bytes_) {
      if (compat_aux_) { CK(cudaStreamSynchronize(st)); cudaFree(compat_aux_); compat_aux_ = nullptr; compat_aux_bytes_ = 0; }
      if (cudaMalloc(&compat_aux_, aux) != cudaSuccess) { cudaGetLastError(); err_ = "compat attention: cudaMalloc of the editor buffers failed (" + std::to_string(aux >> 20) + " MiB)"; return -1; }
      compat_aux_bytes_ = aux;
    }
    float *sim = compat_aux_, *qf = sim + n_sim, *kf = qf + n_q, *vf = kf + n_kv, *of = vf + n_kv;
    compat_split_heads_kernel<<<dim3((a.Nq * a.d + 255) / 256, BH), 256, 0, st>>>(p.q, p.ldq, p.q_sample, nullptr, qf, a.H, a.d, a.Nq);
    compat_split_heads_kernel<<<dim3((a.Nkv * a.d + 255) / 256, BH), 256, 0, st>>>(p.k, p.ldkv, p.kv_sample, p.kv_idx, kf, a.H, a.d, a.Nkv);
    compat_split_heads_kernel<<<dim3((a.Nkv * a.d + 255) / 256, BH), 256, 0, st>>>(p.v, p.ldkv, p.kv_sample, p.kv_idx, vf, a.H, a.d, a.Nkv);
    p.probs = sim;
    compat_scores_kernel<<<dim3((a.Nkv + 63) / 64, (a.Nq + 63) / 64, BH), 256, 0, st>>>(p);
    CK(cudaMemcpyAsync(compat_probs_, sim, n_sim * sizeof(float), cudaMemcpyDeviceToDevice, st));
    compat_softmax_kernel<<<unsigned((rows + 7) / 8), 256, 0, st>>>(compat_probs_, rows, a.Nkv);
    CK(cudaGetLastError());
    if (cc.editor_cb(cc.editor_user, op.tf_index, is_cross ? 1 : 0, tf_place_[op.tf_index], qf, kf, vf, sim, compat_probs_, of, BH, a.Nq, a.Nkv, a.d) != 0) {
      err_ = "compat attention: the editor hook failed";
      return -1;
    }
    const size_t orow = size_t(S) * a.Nq;
    compat_store_out_kernel<<<unsigned((orow * a.H * a.d + 255) / 256), 256, 0, st>>>(of, a.out, a.ldo, a.H * a.d, orow);
    CK(cudaGetLastError());
    return 8;
  }
  compat_scores_kernel<<<dim3((a.Nkv + 63) / 64, (a.Nq + 63) / 64, BH), 256, 0, st>>>(p);
  compat_softmax_kernel<<<unsigned((rows + 7) / 8), 256, 0, st>>>(compat_probs_, rows, a.Nkv);
  CK(cudaGetLastError());
  if (cc.probs_cb(cc.probs_user, op.tf_index, is_cross ? 1 : 0, tf_place_[op.tf_index], compat_probs_, BH, a.Nq, a.Nkv) != 0) {
    err_ = "compat attention: the probabilities hook failed";
    return -1;
  }
  const dim3 g((a.Nq + 31) / 32, BH);
  if (a.d <= 40) compat_pv_kernel<40><<<g, 256, 0, st>>>(p);
  else if (a.d <= 80) compat_pv_kernel<80><<<g, 256, 0, st>>>(p);
  else compat_pv_kernel<160><<<g, 256, 0, st>>>(p);
  CK(cudaGetLastError());
  return 3;
}

long Engine::forward(const float* x, float* eps, int S, const CallCtrl& cc, cudaStream_t st) {
  NvtxRange nvtx_("hedit.unet_forward S=%d uniq=%d", S, cc.n_uniq);
  const bool dedup = cc.n_uniq > 0 && cc.n_uniq < S && cc.uniq_first && cc.uniq_of && !(cc.self_mask & 1u) && !cc.probs_cb && !cc.editor_cb;
  Plan* plan = get_plan(S, dedup ? cc.n_uniq : 0);
  if (!plan) return -1;
  long launches = 0;
  {   // per-call time-embedding rows
    dim3 grid(std::max(1, tproj_total_ / 4 / 256), S);
    gather_rows_kernel<<<grid, 256, 0, st>>>(temb_table_, cc.time_idx, temb_rows_, tproj_total_ / 4);
    ++launches;
  }
  for (Op& op : plan->ops) {
    const long r = launch_op(op, S, x, eps, cc, st);
    if (r < 0) return -1;
    launches += r;
  }
  CK(cudaGetLastError());
  return launches;
}

static bool loop_graphs_env() { static const bool v = !(getenv("HEDIT_LOOP_GRAPH") && atoi(getenv("HEDIT_LOOP_GRAPH")) == 0); return v; }

long Engine::forward_replayed(const float* x, float* eps, int S, const CallCtrl& cc, cudaStream_t st) {
  if (!graph_replay_ || !loop_graphs_env() || cc.probs_cb || cc.editor_cb) return forward(x, eps, S, cc, st);
  NvtxRange nvtx_("hedit.unet_forward_replayed S=%d uniq=%d", S, cc.n_uniq);
  // identity of the launch: every pointer / flag the kernels' parameters are derived from, packed without struct padding
  std::vector<uint8_t> key;
  {
    const void* ptrs[21] = {cc.uniq_first, cc.uniq_of, cc.map_w, cc.ctx_idx, cc.time_idx, cc.self_q, cc.self_k, cc.self_v, cc.feat_src, cc.unit_s0, cc.unit_s1, cc.unit_img,
                            cc.mapper, cc.c_base, cc.c_tar, cc.replace_m, cc.is_replace, cc.blend_acc, cc.blend_alpha, x, eps};
    const int32_t ints[7] = {int32_t(cc.self_mask), cc.n_units, S, cc.blend_rows, cc.map_rows, cc.n_uniq, int32_t(splitk_)};
    key.assign(reinterpret_cast<const uint8_t*>(ptrs), reinterpret_cast<const uint8_t*>(ptrs) + sizeof ptrs);
    key.insert(key.end(), reinterpret_cast<const uint8_t*>(ints), reinterpret_cast<const uint8_t*>(ints) + sizeof ints);
  }
  FwdGraph* ent = nullptr;
  for (auto& g : fwd_graphs_) if (g.key == key) { ent = &g; break; }
  if (ent && ent->exec) {
    if (cudaGraphLaunch(ent->exec, st) == cudaSuccess) return ent->launches;
    cudaGetLastError(); cudaGraphExecDestroy(ent->exec); ent->exec = nullptr; ent->bad = true;
  }
  if (ent && ent->bad) return forward(x, eps, S, cc, st);
  if (!ent) {
    if (fwd_graphs_.size() >= 64) drop_graphs();
    fwd_graphs_.push_back(FwdGraph{key, nullptr, 0, false});
    return forward(x, eps, S, cc, st);
  }
  if (!get_plan(S)) return -1;                                 // plan building (tensor maps, allocation) must not happen under capture
  if (!cap_stream_ && cudaStreamCreateWithFlags(&cap_stream_, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); ent->bad = true; return forward(x, eps, S, cc, st); }
  if (cudaStreamBeginCapture(cap_stream_, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); ent->bad = true; return forward(x, eps, S, cc, st); }
  const long r = forward(x, eps, S, cc, cap_stream_);
  cudaGraph_t graph = nullptr;
  const cudaError_t ec = cudaStreamEndCapture(cap_stream_, &graph);
  if (r < 0 || ec != cudaSuccess || !graph || cudaGraphInstantiate(&ent->exec, graph, 0) != cudaSuccess) {
    cudaGetLastError();
    if (graph) cudaGraphDestroy(graph);
    ent->exec = nullptr; ent->bad = true;
    if (r < 0) return -1;
    return forward(x, eps, S, cc, st);
  }
  cudaGraphDestroy(graph);
  ent->launches = r;
  CK(cudaGraphLaunch(ent->exec, st));
  return r;
}

long Engine::forward_profiled(const float* x, float* eps, int S, const CallCtrl& cc, cudaStream_t st, std::map<std::string, std::pair<double, long>>& acc) {
  Plan* plan = get_plan(S);
  if (!plan) return -1;
  dim3 grid(std::max(1, tproj_total_ / 4 / 256), S);
  gather_rows_kernel<<<grid, 256, 0, st>>>(temb_table_, cc.time_idx, temb_rows_, tproj_total_ / 4);
  std::vector<cudaEvent_t> ev(plan->ops.size() + 1);
  for (auto& e : ev) cudaEventCreate(&e);
  cudaEventRecord(ev[0], st);
  for (size_t i = 0; i < plan->ops.size(); ++i) {
    if (launch_op(plan->ops[i], S, x, eps, cc, st) < 0) return -1;
    cudaEventRecord(ev[i + 1], st);
  }
  CK(cudaStreamSynchronize(st));
  for (size_t i = 0; i < plan->ops.size(); ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
    std::string tag = plan->ops[i].tag;
    static const bool shapes = getenv("HEDIT_PROFILE_SHAPES") && atoi(getenv("HEDIT_PROFILE_SHAPES")) != 0;
    if (shapes && plan->ops[i].kind == OP_GEMM)      // per-shape records (diagnostics)
      tag += "[" + std::to_string(plan->ops[i].gemm.M) + "x" + std::to_string(plan->ops[i].gemm.N) + "x" + std::to_string(plan->ops[i].gemm.num_kb * 64) + "]";
    auto& a = acc[tag];
    a.first += ms; a.second += 1;
  }
  for (auto& e : ev) cudaEventDestroy(e);
  return long(plan->ops.size()) + 1;
}

bool Engine::tensor_info(int i, std::string& name, std::vector<int64_t>& shape) const {
  if (i < 0 || i >= int(slots_.size())) return false;
  auto it = slots_.begin();
  std::advance(it, i);
  name = it->first; shape = it->second.shape;
  return true;
}

}  // namespace hedit
