#include "netexec.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "elementwise.cuh"
#include "tmap.h"
#include "vae.cuh"

namespace hedit {

#define NCK(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      char buf_[512];                                                                              \
      snprintf(buf_, sizeof buf_, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      err_ = buf_;                                                                                 \
      return -1;                                                                                   \
    }                                                                                              \
  } while (0)

NetExec::~NetExec() {
  drop_graphs();
  if (cap_stream_) cudaStreamDestroy(cap_stream_);
  if (arena_) cudaFree(arena_);
}

void NetExec::drop_graphs() {
  for (auto& g : graphs_) if (g.exec) cudaGraphExecDestroy(g.exec);
  graphs_.clear();
}

static bool net_graphs_on() { static const bool v = !(getenv("HEDIT_NET_GRAPH") && atoi(getenv("HEDIT_NET_GRAPH")) == 0); return v; }

bool NetExec::replay(const GraphKey& key, cudaStream_t st) {
  for (auto& g : graphs_)
    if (g.key == key && g.exec) {
      if (cudaGraphLaunch(g.exec, st) != cudaSuccess) { cudaGetLastError(); cudaGraphExecDestroy(g.exec); g.exec = nullptr; g.bad = true; return false; }
      launches_ = g.launches; flops_ = g.flops;
      return true;
    }
  return false;
}

int NetExec::run_or_capture(const GraphKey& key, cudaStream_t st, const std::function<int()>& body) {
  GraphEntry* ent = nullptr;
  for (auto& g : graphs_) if (g.key == key) ent = &g;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (st) cudaStreamIsCapturing(st, &cs);                      // the caller is itself capturing: just add our launches to its graph
  cudaGetLastError();
  if (!net_graphs_on() || cs != cudaStreamCaptureStatusNone || (ent && ent->bad)) { st_ = st; return body(); }
  if (!ent) {                                                  // first sighting: remember the key, launch directly
    if (graphs_.size() >= 16) drop_graphs();
    graphs_.push_back(GraphEntry{key, nullptr, 0, 0.0, false});
    st_ = st;
    return body();
  }
  if (!cap_stream_ && cudaStreamCreateWithFlags(&cap_stream_, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); ent->bad = true; st_ = st; return body(); }
  cudaGraph_t graph = nullptr;
  if (cudaStreamBeginCapture(cap_stream_, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); ent->bad = true; st_ = st; return body(); }
  st_ = cap_stream_;
  const int r = body();
  st_ = st;
  const cudaError_t ec = cudaStreamEndCapture(cap_stream_, &graph);
  if (r) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return r; }
  if (ec != cudaSuccess || !graph || cudaGraphInstantiate(&ent->exec, graph, 0) != cudaSuccess) {
    cudaGetLastError();
    if (graph) cudaGraphDestroy(graph);
    ent->exec = nullptr; ent->bad = true;
    launches_ = 0; flops_ = 0;
    return body();
  }
  cudaGraphDestroy(graph);
  ent->launches = launches_; ent->flops = flops_;
  NCK(cudaGraphLaunch(ent->exec, st));
  return 0;
}

int NetExec::reserve(size_t need, const char* what) {
  if (need <= arena_bytes_) return 0;
  drop_graphs();
  if (arena_) cudaFree(arena_);
  arena_ = nullptr; arena_bytes_ = 0;
  if (cudaMalloc(&arena_, need) != cudaSuccess) { err_ = std::string(what) + " arena cudaMalloc failed (" + std::to_string(need >> 20) + " MiB)"; return -1; }
  arena_bytes_ = need;
  return 0;
}

int NetExec::gemm(const op_t* Ain, int lda, int mode, const ConvGeom* cg, const op_t* Wt, int M, int N, int K, const GemmEpilogue& ep, int ldw) {
  flops_ += 2.0 * M * N * K;
  if (dry_) return 0;
  GemmParams g; int bn;
  GemmEpilogue e = ep;
  if (e.rows_per_group == 0) e.rows_per_group = 1;
  if (!make_gemm(g, bn, Ain, lda, mode, cg, Wt, M, N, K, e, err_, ldw)) return -1;
  NCK(launch_gemm(g, bn, st_));
  ++launches_;
  return 0;
}

int NetExec::conv3(const op_t* x, const op_t* w, int S, int H, int W, int cin, int cout, GemmEpilogue ep, int stride, int pad01) {
  ConvGeom cg{S, H, W, cin, stride, pad01};
  return gemm(x, cin, stride == 2 ? A_CONV3X3S2 : A_CONV3X3, &cg, w, S * H * W, cout, 9 * cin, ep);
}

float2* NetExec::colstats_for(int M, int N, int HW) { return ((HW & 31) || (N & 31)) ? nullptr : A<float2>(size_t((M + 31) / 32) * N); }

int NetExec::gn_fwd(const float* x1, const float2* cs1, int C1, const float* x2, const float2* cs2, int C2, int S, int HW, const float* g,
                    const float* b, float eps, int silu, op_t* out, op_t* raw, float2** stats_out) {
  const int C = C1 + C2;
  float2* stats = A<float2>(size_t(S) * groups_);
  if (stats_out) *stats_out = stats;
  const bool fused = cs1 != nullptr && (x2 == nullptr || cs2 != nullptr);
  float2* partial = nullptr; int nch = 0;
  const int schunk = std::max(16, HW / 256);
  if (!fused) { nch = (HW + schunk - 1) / schunk; partial = A<float2>(size_t(S) * nch * groups_); }
  if (dry_) return 0;
  if (fused) {
    GNFinalizeParams p{cs1, cs2, C1, C2, HW, groups_, eps, stats};
    gn_finalize_kernel<<<dim3(groups_, S), (HW >= 16384 ? 512 : 128), 0, st_>>>(p);
  } else {
    GNStatsParams sp{x1, x2, C1, C2, HW, groups_, schunk, partial};
    gn_stats_kernel<<<dim3(nch, S), std::min(640, ((C / 4 + 31) / 32) * 32), 0, st_>>>(sp);
    gn_partial_finalize_kernel<<<dim3(groups_, S), 128, 0, st_>>>(partial, stats, nch, groups_, 1.0 / (double(HW) * (C / groups_)), eps);
    ++launches_;
  }
  const int chunk = HW >= 4096 ? 32 : 16, quads = C / 4;
  const int threads = std::max(256, quads * std::max(1, (256 + quads - 1) / quads));
  GNApplyParams ap{x1, x2, C1, C2, HW, groups_, chunk, 0, nullptr, g, b, eps, silu, out, raw, stats};
  gn_apply_kernel<<<dim3((HW + chunk - 1) / chunk, S), threads, 0, st_>>>(ap);
  launches_ += 2;
  NCK(cudaGetLastError());
  return 0;
}

int NetExec::upconv_fused(const float* x, const op_t* w_phases, const float* bias, int S, int H, int W, int C, float** out, float2** cs_out) {
  const int M0 = S * H * W;
  op_t* xb = A<op_t>(size_t(M0) * C);
  if (!dry_) { vae_cast_kernel<<<4096, 256, 0, st_>>>(x, xb, size_t(M0) * C / 4); ++launches_; }
  float* y = A<float>(size_t(4) * M0 * C);
  float2* cs = colstats_for(4 * M0, C, 4 * H * W);
  for (int ph = 0; ph < 4; ++ph) {
    ConvGeom cg{S, H, W, C, 1, 0, (ph & 1) - 1, (ph >> 1) - 1};
    GemmEpilogue e; memset(&e, 0, sizeof e);
    e.bias = bias; e.out_f32 = y; e.ldo = C; e.colstats = cs;
    e.up_W = W; e.up_H = H; e.up_py = ph >> 1; e.up_px = ph & 1;
    if (gemm(xb, C, A_CONV2X2, &cg, w_phases + size_t(ph) * C * 4 * C, M0, C, 4 * C, e)) return -1;
  }
  *out = y; *cs_out = cs;
  return 0;
}

int NetExec::attn1h_fwd(const float* x, const float2* cs_x, int S, int N, int C, const float* gng, const float* gnb, float eps, const op_t* w_qkv,
                        const float* b_qkv, const op_t* w_o, const float* b_o, float** out, float2** cs_out, float2** gn_stats_out,
                        const op_t** qkv_out, const op_t** P_out) {
  const int M = S * N;
  op_t* y = A<op_t>(size_t(M) * C);
  if (gn_fwd(x, cs_x, C, nullptr, nullptr, 0, S, N, gng, gnb, eps, 0, y, nullptr, gn_stats_out)) return -1;
  op_t* qkv = A<op_t>(size_t(M) * 3 * C);
  GemmEpilogue e; memset(&e, 0, sizeof e);
  e.bias = b_qkv; e.out_bf16 = qkv; e.ldob = 3 * C;
  if (gemm(y, C, A_LINEAR, nullptr, w_qkv, M, 3 * C, C, e)) return -1;
  float* Sc = A<float>(size_t(N) * N);                       // scores of one sample (reused)
  op_t* P = A<op_t>(size_t(S) * N * N);                      // probabilities of every sample
  op_t* vt = A<op_t>(size_t(S) * C * N);
  op_t* o = A<op_t>(size_t(M) * C);
  if (!dry_) {
    transpose_h16_kernel<<<dim3((C + 31) / 32, (N + 31) / 32, S), dim3(32, 8), 0, st_>>>(qkv + 2 * C, size_t(N) * 3 * C, 3 * C, vt, size_t(C) * N, N, N, C);
    ++launches_;
  }
  const float scale = 1.0f / sqrtf(float(C));
  for (int s = 0; s < S; ++s) {
    const op_t* q = qkv + size_t(s) * N * 3 * C;
    memset(&e, 0, sizeof e); e.out_f32 = Sc; e.ldo = N;
    if (gemm(q, 3 * C, A_LINEAR, nullptr, q + C, N, N, C, e, 3 * C)) return -1;
    if (!dry_) { attn_softmax_rows_kernel<<<dim3(N, 1), 256, 0, st_>>>(Sc, P + size_t(s) * N * N, N, scale * 1.4426950408889634f); ++launches_; }
    memset(&e, 0, sizeof e); e.out_bf16 = o + size_t(s) * N * C; e.ldob = C;
    if (gemm(P + size_t(s) * N * N, N, A_LINEAR, nullptr, vt + size_t(s) * C * N, N, C, N, e)) return -1;
  }
  float* res = A<float>(size_t(M) * C);
  memset(&e, 0, sizeof e); e.bias = b_o; e.residual = x; e.ldr = C; e.out_f32 = res; e.ldo = C; e.colstats = colstats_for(M, C, N);
  if (gemm(o, C, A_LINEAR, nullptr, w_o, M, C, C, e)) return -1;
  if (qkv_out) *qkv_out = qkv;
  if (P_out) *P_out = P;
  *out = res; *cs_out = e.colstats;
  return 0;
}

}  // namespace hedit
