// UNet engine: owns device weights (bf16 GEMM operands, fp32 norms/biases), a static activation arena, and one
// launch plan per batch size (every tensor map is encoded once at plan-build time).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "attention.cuh"
#include "gemm.cuh"

namespace hedit {

typedef op_t bf16;   // historical alias: "the 16-bit operand type"

struct ConvGeom { int S, H, W, C; int stride; int pad01; int ox, oy; };   // H,W = OUTPUT dims; C = input channels; pad01: stride-2 convs padded (0,1,0,1); ox, oy: first-tap offsets of A_CONV2X2
// D[M][N] = A W^T launch description (TMA maps encoded here).  ldw = row stride of W in elements (0: dense [N][Ktot]).
bool make_gemm(GemmParams& g, int& bn, const op_t* A, int lda, int a_mode, const ConvGeom* cg, const op_t* Wt, int M, int N, int Ktot,
               const GemmEpilogue& ep, std::string& err, int ldw = 0);
cudaError_t launch_gemm(const GemmParams& g, int bn, cudaStream_t st);
// Split-K plan for a prepared launch: 1 = launch as is.  > 1: call enable_splitk with a workspace of splits * M * N floats; it rewrites
// g to write fp32 partials and fills `red` for launch_splitk_reduce (which applies the original epilogue).  HEDIT_GEMM_SPLITK=0: never.
int gemm_splitk_splits(const GemmParams& g, int bn);
void enable_splitk(GemmParams& g, int splits, float* ws, SplitKReduceParams& red);
cudaError_t launch_splitk_reduce(const SplitKReduceParams& red, cudaStream_t st);
bool gemm_cluster();      // HEDIT_GEMM_CLUSTER=1: cta_group::2 pair variant (W boxes of BN/2 rows)

struct UNetCfg {
  int in_ch = 4, out_ch = 4, sample = 64;
  int boc[4] = {320, 640, 1280, 1280};
  int layers = 2, heads = 8, ctx_dim = 768, groups = 32;
  int ctx_len = 77;
};

struct ResW {
  int cin = 0, cout = 0, temb_off = 0;
  float *n1g = 0, *n1b = 0, *n2g = 0, *n2b = 0, *b1 = 0, *b2 = 0, *bsc = 0;
  bf16 *w1 = 0, *w2 = 0, *wsc = 0;
};
struct TfW {
  int C = 0, cross_index = 0;
  float *gng = 0, *gnb = 0, *b_in = 0, *b_out = 0, *ln1g = 0, *ln1b = 0, *ln2g = 0, *ln2b = 0, *ln3g = 0, *ln3b = 0;
  float *b_o1 = 0, *b_o2 = 0, *b_ff1 = 0, *b_ff2 = 0;
  bf16 *w_in = 0, *w_out = 0, *w_qkv = 0, *w_o1 = 0, *w_q2 = 0, *w_kv2 = 0, *w_o2 = 0, *w_ff1 = 0, *w_ff2 = 0;
  bf16* kv_cache = 0;      // [max_ctx][77][2C]
};

// How one named reference tensor is converted into the engine's layout.
struct WeightSlot {
  enum Kind { F32_COPY, BF16_ROWS, BF16_CONV3, BF16_GEGLU_ROWS, F32_GEGLU_VEC } kind;
  void* dst = nullptr;
  void* dst_up = nullptr;       // BF16_CONV3 of an upsampler: additionally the 4 pre-summed 2x2 phase kernels [4][O][2][2][I] (fused upsample conv)
  size_t dst_off_elems = 0;     // element offset inside dst (row-concatenated tensors)
  std::vector<int64_t> shape;   // expected source shape
  bool loaded = false;
};

enum OpKind { OP_CONV_IN, OP_GN_STATS, OP_GN_FINALIZE, OP_GN_APPLY, OP_GEMM, OP_LN, OP_SELF_ATTN, OP_CROSS_ATTN, OP_UPSAMPLE, OP_CAST, OP_CONV_OUT, OP_FEAT_COPY, OP_EXPAND, OP_SPLITK_REDUCE };

struct Op {
  OpKind kind;
  // generic payloads (only the ones relevant to `kind` are used)
  GemmParams gemm; int gemm_bn = 0;
  SplitKReduceParams red;
  AttnParams attn; int dch = 0, bkv = 0, tf_index = 0, blend_layer = -1;
  const bf16 *cq = nullptr, *ck = nullptr, *cv = nullptr; int c_ldq = 0, c_ldkv = 0;   // raw operand pointers for the compat attention path
  const float* f_in = nullptr; const float* f_in2 = nullptr; float* f_out = nullptr;
  const bf16* h_in = nullptr; bf16* h_out = nullptr; bf16* h_out2 = nullptr;
  const float* gamma = nullptr; const float* beta = nullptr; const float* w = nullptr; const float* b = nullptr;
  float2* partial = nullptr;
  const float2 *cs1 = nullptr, *cs2 = nullptr; float2* stats = nullptr;     // fused GroupNorm statistics (gemm colstats -> gn_finalize)
  int C1 = 0, C2 = 0, HW = 0, H = 0, W = 0, rows = 0, chunk = 0, nchunks = 0, silu = 0;
  float eps = 0.f;
  size_t count = 0;
  int nS = 0;              // samples this op runs on (the de-duplicated prefix runs on fewer than the launch's S)
  const char* tag = "";
};

// Compatibility hook: called on the host between softmax and P.V of every attention layer with the materialised fp32 probabilities
// [(S*heads)][n_query][n_key] (device memory, edited in place on the call's stream).  place: 0 down, 1 mid, 2 up.  Non-zero = abort.
typedef int (*AttnProbsFn)(void* user, int tf_index, int is_cross, int place, float* probs, int batch_heads, int n_query, int n_key);

// Editor form of the hook (MasaCtrl's protocol): q [(S*heads)][n_query][d], k / v [(S*heads)][n_key][d], sim = scaled scores and
// attn = softmax(sim), both [(S*heads)][n_query][n_key]; the hook writes the layer output [S][n_query][heads*d] into `out`.
typedef int (*AttnEditorFn)(void* user, int tf_index, int is_cross, int place, float* q, float* k, float* v, float* sim, float* attn,
                            float* out, int batch_heads, int n_query, int n_key, int d);

// Per-call attention control (device arrays prepared by the edit loop).
struct CallCtrl {
  const int* ctx_idx = nullptr;        // [S] index into the text K/V cache
  const int* time_idx = nullptr;       // [S] row of the timestep-embedding table
  // self-attention injection (applied on transformer blocks whose bit is set in self_mask)
  uint32_t self_mask = 0;
  const int *self_q = nullptr, *self_k = nullptr, *self_v = nullptr;
  // Plug-and-Play feature injection at up_blocks[1].resnets[1] (pnp_utils.py:138-146): [S] source sample per sample, or null
  const int* feat_src = nullptr;
  // cross-attention work units + P2P edit tables for the current step
  const int *unit_s0 = nullptr, *unit_s1 = nullptr, *unit_img = nullptr;
  int n_units = 0;
  const int* mapper = nullptr; const float* c_base = nullptr; const float* c_tar = nullptr;
  const float* replace_m = nullptr; const int* is_replace = nullptr;
  const float* map_w = nullptr; int map_rows = 1;
  float* blend_acc = nullptr; const float* blend_alpha = nullptr; int blend_rows = 2;
  // Context-free prefix de-duplication: samples of one launch that share a LATENT (e.g. [x,null] [x,src] [x,tar]) compute identical
  // activations until the first cross-attention (conv_in, down_blocks[0].resnets[0], the first transformer block up to and including its
  // self-attention).  With n_uniq > 0 that prefix runs once per distinct latent and is broadcast: uniq_first[u] = a sample holding
  // distinct latent u, uniq_of[s] = the distinct latent of sample s (device arrays).  Requires one timestep for the whole launch and no
  // self-attention control on transformer block 0.  Bit-identical to the full evaluation (every kernel is batch-invariant).
  const int *uniq_first = nullptr, *uniq_of = nullptr; int n_uniq = 0;
  // compat path (compat_attn.cuh): materialised probabilities + host callback instead of the fused attention kernels
  AttnProbsFn probs_cb = nullptr; void* probs_user = nullptr;
  AttnEditorFn editor_cb = nullptr; void* editor_user = nullptr;
};

struct Plan {
  int S = 0;
  std::vector<Op> ops;
  size_t launches = 0;
};

class Engine {
 public:
  Engine(const UNetCfg& cfg, int max_samples, int max_ctx);
  ~Engine();
  bool ok() const { return err_.empty(); }
  const std::string& error() const { return err_; }

  int load_tensor(const char* name, const float* src, const int64_t* dims, int ndim, cudaStream_t st);
  int finalize_weights(std::string* missing);

  // text contexts -> cached K/V for the 16 cross-attention layers; ctx fp32 [n_ctx][77][ctx_dim] (host or device)
  int set_contexts(const float* ctx, int n_ctx, cudaStream_t st);
  // timestep-embedding table for a list of timesteps (host array)
  int set_timesteps(const float* ts, int n, cudaStream_t st);

  // x [S][4][h][w] fp32 NCHW (device) -> eps [S][4][h][w] (device).  Returns kernels launched, <0 on error.
  long forward(const float* x, float* eps, int S, const CallCtrl& cc, cudaStream_t st);
  // Same launches with a CUDA-event pair around every op; accumulates milliseconds per op tag (diagnostics / bench roofline).
  long forward_profiled(const float* x, float* eps, int S, const CallCtrl& cc, cudaStream_t st, std::map<std::string, std::pair<double, long>>& acc);
  int tensor_count() const { return int(slots_.size()); }
  bool tensor_info(int i, std::string& name, std::vector<int64_t>& shape) const;

  const UNetCfg& cfg() const { return cfg_; }
  int max_samples() const { return maxS_; }
  int n_blend_layers() const { return n_blend_layers_; }
  int n_tf() const { return int(tfs_.size()); }
  int tf_tokens(int i) const { return tf_tokens_[i]; }
  int tf_place(int i) const { return tf_place_[i]; }
  int latent_elems() const { return cfg_.in_ch * cfg_.sample * cfg_.sample; }
  double flops_per_sample() const { return flops_per_sample_; }
  void* scratch_alloc(size_t bytes);     // persistent device allocations owned by the engine
  // Slot `i` of the sampling loop's buffer pool: same request sequence => same addresses across edits (no per-edit cudaMalloc /
  // cudaFree, and the CUDA graphs below stay valid); a slot is re-allocated only when a larger size is asked for.
  void* loop_slot(size_t index, size_t bytes);
  // forward() replayed from a CUDA graph: a launch is identified by (x, eps, S, every field of cc); its first occurrence launches
  // directly, its second is captured on a private stream and instantiated, later ones are a single cudaGraphLaunch.  Per-step tables must
  // therefore sit at FIXED addresses (the loop stages the step's rows).  Calls with a probabilities hook are never captured.
  long forward_replayed(const float* x, float* eps, int S, const CallCtrl& cc, cudaStream_t st);
  void set_graph_replay(bool on) { graph_replay_ = on; }
  void set_prefix_dedup(bool on) { prefix_dedup_ = on; }
  // split-K for launches with far fewer tiles than SMs (single-image use).  Off by default: a split changes the fp32 summation order of
  // the affected layers, so results stop being bit-identical ACROSS batch sizes (they stay run-to-run reproducible)
  void set_splitk(bool on) { splitk_ = on; }
  bool splitk() const { return splitk_; }
  bool prefix_dedup() const { return prefix_dedup_; }
  void drop_graphs();

  std::string err_;

 private:
  friend struct PlanBuilder;
  Plan* get_plan(int S, int U = 0);
  long launch_op(Op& op, int S, const float* x, float* eps, const CallCtrl& cc, cudaStream_t st);
  void reg(const std::string& name, WeightSlot::Kind k, void* dst, size_t off, std::vector<int64_t> shape);
  template <typename T> T* dalloc(size_t n);

  UNetCfg cfg_;
  int maxS_, maxCtx_;
  std::map<std::string, WeightSlot> slots_;
  std::vector<void*> owned_;
  // weights
  float *conv_in_w_ = 0, *conv_in_b_ = 0, *conv_out_b_ = 0, *norm_out_g_ = 0, *norm_out_b_ = 0;
  bf16* conv_out_w_ = 0;
  float *t_w1_ = 0, *t_b1_ = 0, *t_w2_ = 0, *t_b2_ = 0, *tproj_w_ = 0, *tproj_b_ = 0;
  int tproj_total_ = 0;
  std::vector<ResW> res_;       // in forward order
  std::vector<TfW> tfs_;        // in forward order (== controller layer order / 2)
  std::vector<int> tf_tokens_, tf_place_;
  float* compat_probs_ = nullptr; size_t compat_probs_bytes_ = 0;
  float* compat_aux_ = nullptr; size_t compat_aux_bytes_ = 0;        // editor form: sim | q | k | v | out
  long compat_attention(const Op& op, bool is_cross, int S, const CallCtrl& cc, cudaStream_t st);
  std::vector<bf16*> down_w_, up_w_, up_wp_;
  std::vector<float*> down_b_, up_b_;
  int n_blend_layers_ = 0;
  double flops_per_sample_ = 0;
  // runtime buffers
  uint8_t* arena_ = 0; size_t arena_bytes_ = 0;
  float* temb_act_ = 0;         // [maxT][temb_dim]
  float* temb_table_ = 0;       // [maxT][tproj_total]
  float* temb_rows_ = 0;        // [maxS][tproj_total] gathered per call
  float* ts_dev_ = 0;
  int maxT_ = 1008, nT_ = 0;      // timesteps per edit (incl. the final previous timestep): the full 1000-step training schedule fits (81 MB of fp32 time projections)
  bf16* ctx_bf16_ = 0;          // [max_ctx*77][ctx_dim]
  float* stage_ = 0; size_t stage_elems_ = 0;   // fp32 staging for weight upload
  std::map<int, std::unique_ptr<Plan>> plans_;
  std::vector<std::pair<void*, size_t>> loop_pool_;
  struct FwdGraph { std::vector<uint8_t> key; cudaGraphExec_t exec; long launches; bool bad; };
  std::vector<FwdGraph> fwd_graphs_;
  cudaStream_t cap_stream_ = nullptr;
  bool graph_replay_ = true;
  bool prefix_dedup_ = true;
  bool splitk_ = false;
};

}  // namespace hedit
