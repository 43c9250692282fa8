// The reverse-time bridge sampling loop (reference: text-guided/inversion/p2p_h_edit.py:529-701 implicit,
// :380-520 explicit) for a batch of B independent images, with zero host synchronisation inside the loop.
//
// Sample layouts inside one UNet launch (per image b):
//   schedule 0 (reference call pattern, K=1: 9 sample-forwards / step)
//     A : [xo,∅] [xe,∅] [xo,src] [xe,src]          at t   (P2P off)            p2p_h_edit.py:613
//     B : [x_opt,src]                               at tt  (P2P off)            :644
//     C : [xo',∅] [x_opt,∅] [xo',src] [x_opt,tar]   at tt  (P2P on)             :652
//   schedule 1 (exact reuse, 7 / step): the [xo',∅] and [xo',src] outputs of C are bit-for-bit the [xo,∅], [xo,src]
//     inputs of the next step's A (P2P never modifies the source branch, LocalBlend leaves row 0 untouched), and B
//     shares C's launch:
//     A': [xe,∅] [xe,src]                           at t
//     BC: [xo',∅] [x_opt,∅] [xo',src] [x_opt,tar] [x_opt,src]   at tt  (P2P pair = samples 2,3)
//   schedule 2 (opt-in, cfg_src == 1: 5 / step): u + 1 * (c - u) == c, so the unconditional forwards that only feed the
//     source-guided combine are dropped:
//     A": [xe,src]                                  at t
//     BC": [x_opt,∅] [xo',src] [x_opt,tar] [x_opt,src]              at tt  (P2P pair = samples 1,2)
//   explicit form (one launch / step, 5 sample-forwards instead of the reference's 9):
//     E : [xo,∅] [xe,∅] [xo,src] [xe,src] [xe,tar]  at t   (P2P pair = samples 2,4)    :459,484,492
#include <vector>

#include "../../include/hedit_b200.h"
#include "engine.h"
#include "hstep.cuh"
#include "nvtx.h"

namespace hedit {

#define CKE(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      char buf_[512];                                                                              \
      snprintf(buf_, sizeof buf_, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      E.err_ = buf_;                                                                               \
      return -1;                                                                                   \
    }                                                                                              \
  } while (0)

__global__ void gather_latents_kernel(const float* __restrict__ pool, const int* __restrict__ idx, float* __restrict__ dst, int n4) {
  const int s = blockIdx.y;
  const float4* src = reinterpret_cast<const float4*>(pool) + size_t(idx[s]) * n4;
  float4* d = reinterpret_cast<float4*>(dst) + size_t(s) * n4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) d[i] = src[i];
}

// One UNet launch description (host side) + its device mirrors.
struct CallDesc {
  int S = 0;
  int pool_off = 0;                 // first sample of this call inside the eps pool
  std::vector<int> lat, ctx, us0, us1, uimg, sq, sk, ufirst, uof;
  int *d_lat = 0, *d_ctx = 0, *d_us0 = 0, *d_us1 = 0, *d_uimg = 0, *d_sq = 0, *d_sk = 0, *d_ufirst = 0, *d_uof = 0;
  int n_units = 0, n_uniq = 0;
  bool p2p = false;      // attention control active in this launch (P2P edit / MasaCtrl)
  void add(int l, int c) { lat.push_back(l); ctx.push_back(c); sq.push_back(S); sk.push_back(S); ++S; }
  void unit(int a, int b, int img) { us0.push_back(a); us1.push_back(b); uimg.push_back(img); ++n_units; }
};

// per-edit device allocations, released when the edit returns
// The loop's device buffers live in engine-owned slots, requested in a fixed order: edits of the same shape get the same addresses
// (no cudaMalloc / cudaFree per edit; the engine's CUDA graphs of the UNet launches stay valid from one edit to the next).
struct TempPool {
  Engine* E = nullptr;
  size_t next = 0;
  void* get(size_t bytes) { return E->loop_slot(next++, bytes); }
};

static int upload_ints(Engine& E, TempPool& tp, const std::vector<int>& v, int** out, cudaStream_t st) {
  *out = reinterpret_cast<int*>(tp.get(std::max<size_t>(v.size(), 1) * sizeof(int)));
  if (!*out) return -1;
  if (!v.empty()) CKE(cudaMemcpyAsync(*out, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  return 0;
}
static int finish_call(Engine& E, TempPool& tp, CallDesc& c, cudaStream_t st) {
  // distinct latents of this launch (samples that differ only in their text context share the context-free prefix of the UNet)
  c.ufirst.clear(); c.uof.assign(c.S, 0);
  for (int s = 0; s < c.S; ++s) {
    int u = -1;
    for (size_t k = 0; k < c.ufirst.size(); ++k) if (c.lat[c.ufirst[k]] == c.lat[s]) { u = int(k); break; }
    if (u < 0) { u = int(c.ufirst.size()); c.ufirst.push_back(s); }
    c.uof[s] = u;
  }
  c.n_uniq = int(c.ufirst.size());
  if (upload_ints(E, tp, c.ufirst, &c.d_ufirst, st) || upload_ints(E, tp, c.uof, &c.d_uof, st)) return -1;
  if (upload_ints(E, tp, c.lat, &c.d_lat, st) || upload_ints(E, tp, c.ctx, &c.d_ctx, st) || upload_ints(E, tp, c.us0, &c.d_us0, st) ||
      upload_ints(E, tp, c.us1, &c.d_us1, st) || upload_ints(E, tp, c.uimg, &c.d_uimg, st) || upload_ints(E, tp, c.sq, &c.d_sq, st) ||
      upload_ints(E, tp, c.sk, &c.d_sk, st))
    return -1;
  return 0;
}

struct LoopBuffers {
  float *lat = 0, *eps = 0, *corr = 0, *xin = 0, *zs = 0, *blend_acc = 0, *c_base = 0, *c_tar = 0, *replace_m = 0, *blend_alpha = 0, *map_w = 0;
  float2* partial = 0;
  int *mapper = 0, *is_replace = 0, *has_blend = 0, *tidx = 0, *tidx_cur = 0;
  float *c_base_cur = 0, *c_tar_cur = 0;
  int *iuA = 0, *icA = 0, *iuA0 = 0, *icA0 = 0, *iu = 0, *ics = 0, *ict = 0;
};

int run_edit(Engine& E, hedit_edit_args& a, cudaStream_t st) {
  NvtxRange nvtx_edit_("hedit.edit B=%d steps=%d", a.B, a.steps);
  const UNetCfg& c = E.cfg();
  const int B = a.B, T = a.steps, K = a.explicit_form ? 1 : std::max(1, a.opt_steps);
  const int n = E.latent_elems();
  const bool masa = a.masa != 0;
  const bool baseline = a.variant == 2;        // ef_or_pnp_inv_w_p2p / ef_or_pnp_inv_w_masactrl
  const bool p2p = a.use_p2p != 0 && (a.variant == 0 || baseline) && !masa;
  const bool pnp = a.pnp != 0;
  const bool ctrl = p2p || masa || pnp;   // launches C / BC / E run with attention control
  const bool blend = p2p && a.has_blend != nullptr && a.blend_alpha != nullptr;
  if (blend && (E.n_blend_layers() == 0 || c.sample != 64)) {
    // LocalBlend reads the 16x16 cross-attention maps of a 64x64 latent (ptp_classes.py:54-58 reshapes to 16x16); any other geometry makes
    // the reference fail in that reshape, so it is an error here too rather than a silently unblended edit
    E.err_ = "LocalBlend needs a 64x64 latent with 16x16 cross-attention layers (ptp_classes.py:54-58)";
    return -1;
  }
  const int brows = (blend && a.blend_rows == 4) ? 4 : 2;     // 4: LocalBlend substruct_words maps ride along
  const cudaMemcpyKind kIn = a.buffers_on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  const cudaMemcpyKind kOut = a.buffers_on_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  if (B < 1 || T < 1) { E.err_ = "bad batch/steps"; return -1; }
  if (pnp && (a.explicit_form || (a.variant != 0 && a.variant != 2) || a.use_p2p || masa || !a.pnp_qk_on || !a.pnp_feat_on || (a.xt_is_pair && a.variant != 2))) {
    E.err_ = "Plug-and-Play runs the implicit form only (pnp_h_edit.py:33), without P2P / MasaCtrl, and needs both per-step flag arrays";
    return -1;
  }
  if (masa && !a.masa_step_on) { E.err_ = "masa needs masa_step_on[steps * opt_steps]"; return -1; }
  if (baseline && (a.guidance || a.pre_step)) { E.err_ = "variant 2 (baseline samplers) runs without reward guidance or pre_step"; return -1; }
  if (a.guidance && (a.explicit_form || !a.x0_coef || !a.guid_x0 || !a.guid_grad)) {
    E.err_ = "reward guidance runs in the implicit form and needs x0_coef, guid_x0 and guid_grad";
    return -1;
  }
  if (a.xt_is_pair && !a.explicit_form && a.variant == 0 && a.schedule != 0) {
    E.err_ = "xt_is_pair (single-step use) needs schedule 0: the exact-reuse schedule carries UNet outputs across timesteps";
    return -1;
  }
  TempPool tp; tp.E = &E;

  // ---- call descriptors.  latent pool slots: xt[b][row] = 2b+row ; xprev[b][row] = 2B+2b+row ; xopt[b] = 4B+b
  // contexts: 0 = "", 1+2b = src_b, 2+2b = tar_b
  std::vector<CallDesc> calls;
  auto XT = [&](int b, int r) { return 2 * b + r; };
  auto XP = [&](int b, int r) { return 2 * B + 2 * b + r; };
  auto XO = [&](int b) { return 4 * B + b; };
  std::vector<int> iuA(2 * B), icA(2 * B), iuA0(2 * B), icA0(2 * B), iu(B), ics(B), ict(B);
  int pool = 0;
  if (a.variant == 1) {
    // h_Edit_R_* : only the edit row is denoised; no attention control
    CallDesc A; A.pool_off = 0;
    for (int b = 0; b < B; ++b) {
      const int s = A.S;
      A.add(XT(b, 1), 0); A.add(XT(b, 1), 1 + 2 * b);
      if (a.explicit_form) A.add(XT(b, 1), 2 + 2 * b);
      for (int j = s; j < A.S; ++j) A.unit(j, -1, b);
      iuA[2 * b] = iuA[2 * b + 1] = iuA0[2 * b] = iuA0[2 * b + 1] = s;
      icA[2 * b] = icA[2 * b + 1] = icA0[2 * b] = icA0[2 * b + 1] = s + 1;
      if (a.explicit_form) { iu[b] = s; ics[b] = s + 1; ict[b] = s + 2; }
    }
    calls.push_back(A);
    pool = A.S;
    if (!a.explicit_form) {
      CallDesc C; C.pool_off = pool;
      for (int b = 0; b < B; ++b) {
        const int s = C.S;
        C.add(XO(b), 0); C.add(XO(b), 1 + 2 * b); C.add(XO(b), 2 + 2 * b);
        for (int j = 0; j < 3; ++j) C.unit(s + j, -1, b);
        iu[b] = C.pool_off + s; ics[b] = C.pool_off + s + 1; ict[b] = C.pool_off + s + 2;
      }
      pool += C.S;
      calls.push_back(C);
    }
  } else if (pnp && !baseline) {
    // h_Edit_PnP_implicit (pnp_h_edit.py:104-160).  Reference pattern per step: A (4) at t, then at tt [x_opt,src], [x_opt,null] and
    // the injected pair ([xo',src],[x_opt,tar]) = 8 sample-forwards.  schedule 1 reuses the pair's untouched source sample
    // [xo',src] as the next step's [xo,src] (injection only ever writes the target sample): 7 per step.
    CallDesc A, BC; BC.p2p = true;
    A.pool_off = 0;
    for (int b = 0; b < B; ++b) {
      const int s = A.S;
      A.add(XT(b, 0), 0); A.add(XT(b, 1), 0); A.add(XT(b, 1), 1 + 2 * b);
      if (a.schedule == 0) A.add(XT(b, 0), 1 + 2 * b);
      for (int j = s; j < A.S; ++j) A.unit(j, -1, b);
      iuA[2 * b] = iuA0[2 * b] = s; iuA[2 * b + 1] = iuA0[2 * b + 1] = s + 1;
      icA[2 * b + 1] = icA0[2 * b + 1] = s + 2;
      icA0[2 * b] = (a.schedule == 0) ? s + 3 : s + 2;       // step 0: xo == xe
      icA[2 * b] = s + 3;                                     // schedule 1: patched below to the pair's source sample
    }
    BC.pool_off = A.S;
    for (int b = 0; b < B; ++b) {
      const int s = BC.S;
      BC.add(XP(b, 0), 1 + 2 * b); BC.add(XO(b), 2 + 2 * b); BC.add(XO(b), 1 + 2 * b); BC.add(XO(b), 0);
      for (int j = 0; j < 4; ++j) BC.unit(s + j, -1, b);
      BC.sq[s + 1] = s;                                       // target <- source (q, k and the resnet features)
      if (a.schedule != 0) icA[2 * b] = BC.pool_off + s;
      ict[b] = BC.pool_off + s + 1; ics[b] = BC.pool_off + s + 2; iu[b] = BC.pool_off + s + 3;
    }
    pool = A.S + BC.S;
    calls.push_back(A); calls.push_back(BC);
  } else if (baseline) {
    // p2p_baselines.py:153-165: torch.cat([xt] * 2) with [uncond, uncond, src, tar] -> [xo,null] [xe,null] [xo,src] [xe,tar], P2P pair = (2, 3)
    CallDesc e; e.p2p = ctrl; e.pool_off = 0;
    for (int b = 0; b < B; ++b) {
      const int s = e.S;
      e.add(XT(b, 0), 0); e.add(XT(b, 1), 0); e.add(XT(b, 0), 1 + 2 * b); e.add(XT(b, 1), 2 + 2 * b);
      e.unit(s, -1, b); e.unit(s + 1, -1, b);
      if (p2p) { e.unit(s + 2, s + 3, b); e.sq[s + 3] = s + 2; } else { e.unit(s + 2, -1, b); e.unit(s + 3, -1, b); }
      if (pnp) e.sq[s + 3] = s + 2;             // Plug-and-Play baselines (pnp_baselines.py:317): the target of the pair takes the source's q, k and features
      e.sk[s + 1] = s; e.sk[s + 3] = s + 2;
      iuA[2 * b] = iuA0[2 * b] = s; iuA[2 * b + 1] = iuA0[2 * b + 1] = s + 1;
      icA[2 * b] = icA0[2 * b] = s + 2; icA[2 * b + 1] = icA0[2 * b + 1] = s + 3;
      iu[b] = s + 1; ics[b] = s + 2; ict[b] = s + 3;
    }
    calls.push_back(e);
    pool = e.S;
  } else if (a.explicit_form) {
    CallDesc e; e.p2p = ctrl; e.pool_off = 0;
    for (int b = 0; b < B; ++b) {
      const int s = e.S;
      e.add(XT(b, 0), 0); e.add(XT(b, 1), 0); e.add(XT(b, 0), 1 + 2 * b); e.add(XT(b, 1), 1 + 2 * b); e.add(XT(b, 1), 2 + 2 * b);
      e.unit(s, -1, b); e.unit(s + 1, -1, b); e.unit(s + 3, -1, b);
      if (p2p) { e.unit(s + 2, s + 4, b); e.sq[s + 4] = s + 2; } else { e.unit(s + 2, -1, b); e.unit(s + 4, -1, b); }
      e.sk[s + 1] = s; e.sk[s + 4] = s + 2;
      iuA[2 * b] = iuA0[2 * b] = s; iuA[2 * b + 1] = iuA0[2 * b + 1] = s + 1;
      icA[2 * b] = icA0[2 * b] = s + 2; icA[2 * b + 1] = icA0[2 * b + 1] = s + 3;
      iu[b] = s + 1; ics[b] = s + 3; ict[b] = s + 4;
    }
    calls.push_back(e);
    pool = e.S;
  } else if (a.schedule == 0) {
    CallDesc A, Bc, C; C.p2p = ctrl;
    A.pool_off = 0;
    for (int b = 0; b < B; ++b) {
      const int s = A.S;
      A.add(XT(b, 0), 0); A.add(XT(b, 1), 0); A.add(XT(b, 0), 1 + 2 * b); A.add(XT(b, 1), 1 + 2 * b);
      for (int j = 0; j < 4; ++j) A.unit(s + j, -1, b);
      iuA[2 * b] = iuA0[2 * b] = s; iuA[2 * b + 1] = iuA0[2 * b + 1] = s + 1;
      icA[2 * b] = icA0[2 * b] = s + 2; icA[2 * b + 1] = icA0[2 * b + 1] = s + 3;
    }
    Bc.pool_off = A.S;
    for (int b = 0; b < B; ++b) { Bc.add(XO(b), 1 + 2 * b); Bc.unit(b, -1, b); ics[b] = Bc.pool_off + b; }
    C.pool_off = A.S + Bc.S;
    for (int b = 0; b < B; ++b) {
      const int s = C.S;
      C.add(XP(b, 0), 0); C.add(XO(b), 0); C.add(XP(b, 0), 1 + 2 * b); C.add(XO(b), 2 + 2 * b);
      C.unit(s, -1, b); C.unit(s + 1, -1, b);
      if (p2p) { C.unit(s + 2, s + 3, b); C.sq[s + 3] = s + 2; } else { C.unit(s + 2, -1, b); C.unit(s + 3, -1, b); }
      C.sk[s + 1] = s; C.sk[s + 3] = s + 2;
      iu[b] = C.pool_off + s + 1; ict[b] = C.pool_off + s + 3;
    }
    pool = A.S + Bc.S + C.S;
    calls.push_back(A); calls.push_back(Bc); calls.push_back(C);
  } else if (a.schedule == 2) {
    // schedule 2 (w_src == 1 only, 5 / step): with cfg_src = 1 the source-guided noise u + 1 * (c - u) IS the conditional prediction
    // c (in exact arithmetic; the reference's fp32 evaluation differs from c by at most one rounding), so the unconditional
    // forwards that only feed that combine -- [xe,null] at t and [xo',null] at tt -- are not needed.  The reverse-step kernel reads
    // its "u" and "c" terms from the same sample, which makes eps = c + 1 * (c - c) = c exactly.
    if (a.w_src != 1.0f || masa) { E.err_ = "schedule 2 needs cfg_src == 1 and no MasaCtrl (whose unconditional target attends to the unconditional source)"; return -1; }
    CallDesc A, BC; BC.p2p = ctrl;
    A.pool_off = 0;
    for (int b = 0; b < B; ++b) {
      const int s = A.S;
      A.add(XT(b, 1), 1 + 2 * b);
      A.unit(s, -1, b);
      iuA0[2 * b] = iuA0[2 * b + 1] = icA0[2 * b] = icA0[2 * b + 1] = s;        // step 0: xo == xe
      iuA[2 * b + 1] = icA[2 * b + 1] = s;
    }
    BC.pool_off = A.S;
    for (int b = 0; b < B; ++b) {
      const int s = BC.S;
      BC.add(XO(b), 0); BC.add(XP(b, 0), 1 + 2 * b); BC.add(XO(b), 2 + 2 * b); BC.add(XO(b), 1 + 2 * b);
      BC.unit(s, -1, b); BC.unit(s + 3, -1, b);
      if (p2p) { BC.unit(s + 1, s + 2, b); BC.sq[s + 2] = s + 1; } else { BC.unit(s + 1, -1, b); BC.unit(s + 2, -1, b); }
      iuA[2 * b] = icA[2 * b] = BC.pool_off + s + 1;                               // reused next step for the orig row
      iu[b] = BC.pool_off + s; ict[b] = BC.pool_off + s + 2; ics[b] = BC.pool_off + s + 3;
    }
    pool = A.S + BC.S;
    calls.push_back(A); calls.push_back(BC);
  } else {
    CallDesc A, BC; BC.p2p = ctrl;
    A.pool_off = 0;
    for (int b = 0; b < B; ++b) {
      const int s = A.S;
      A.add(XT(b, 1), 0); A.add(XT(b, 1), 1 + 2 * b);
      A.unit(s, -1, b); A.unit(s + 1, -1, b);
      iuA0[2 * b] = iuA0[2 * b + 1] = s; icA0[2 * b] = icA0[2 * b + 1] = s + 1;   // step 0: xo == xe
      iuA[2 * b + 1] = s; icA[2 * b + 1] = s + 1;
    }
    BC.pool_off = A.S;
    for (int b = 0; b < B; ++b) {
      const int s = BC.S;
      BC.add(XP(b, 0), 0); BC.add(XO(b), 0); BC.add(XP(b, 0), 1 + 2 * b); BC.add(XO(b), 2 + 2 * b); BC.add(XO(b), 1 + 2 * b);
      BC.unit(s, -1, b); BC.unit(s + 1, -1, b); BC.unit(s + 4, -1, b);
      if (p2p) { BC.unit(s + 2, s + 3, b); BC.sq[s + 3] = s + 2; } else { BC.unit(s + 2, -1, b); BC.unit(s + 3, -1, b); }
      BC.sk[s + 1] = s; BC.sk[s + 3] = s + 2;
      iuA[2 * b] = BC.pool_off + s; icA[2 * b] = BC.pool_off + s + 2;              // reused next step for the orig row
      iu[b] = BC.pool_off + s + 1; ict[b] = BC.pool_off + s + 3; ics[b] = BC.pool_off + s + 4;
    }
    pool = A.S + BC.S;
    calls.push_back(A); calls.push_back(BC);
  }
  int maxS = 0;
  for (auto& cd : calls) { maxS = std::max(maxS, cd.S); if (finish_call(E, tp, cd, st)) return -1; }
  if (maxS > E.max_samples()) { E.err_ = "batch needs more samples per launch than the engine was created for"; return -1; }
  if (1 + 2 * B > 4096) { E.err_ = "too many contexts"; return -1; }

  // ---- device buffers: engine-owned slots requested in a fixed order (an edit of the same shape reuses them at the same addresses)
  LoopBuffers L;
  auto fa = [&](size_t nf) { return reinterpret_cast<float*>(tp.get(nf * sizeof(float))); };
  L.lat = fa(size_t(5) * B * n); L.eps = fa(size_t(pool) * n); L.corr = fa(size_t(B) * n); L.xin = fa(size_t(maxS) * n);
  L.zs = fa(size_t(B) * T * n);
  const int nparts = 16;
  L.partial = reinterpret_cast<float2*>(tp.get(size_t(B) * nparts * sizeof(float2)));
  L.tidx = reinterpret_cast<int*>(tp.get(size_t(T + 1) * maxS * sizeof(int)));
  // fixed-address copies of the CURRENT launch's time-index row and P2P coefficient rows (what the replayed graphs read)
  L.tidx_cur = reinterpret_cast<int*>(tp.get(size_t(maxS) * sizeof(int)));
  L.c_base_cur = fa(size_t(B) * 80); L.c_tar_cur = fa(size_t(B) * 80);
  if (!L.lat || !L.eps || !L.corr || !L.xin || !L.zs || !L.partial || !L.tidx || !L.tidx_cur || !L.c_base_cur || !L.c_tar_cur) return -1;
  if (upload_ints(E, tp, iuA, &L.iuA, st) || upload_ints(E, tp, icA, &L.icA, st) || upload_ints(E, tp, iuA0, &L.iuA0, st) ||
      upload_ints(E, tp, icA0, &L.icA0, st) || upload_ints(E, tp, iu, &L.iu, st) || upload_ints(E, tp, ics, &L.ics, st) || upload_ints(E, tp, ict, &L.ict, st))
    return -1;
  {   // time-index rows: row j = all samples at timestep index j
    std::vector<int> t(size_t(T + 1) * maxS);
    for (int j = 0; j <= T; ++j) for (int s = 0; s < maxS; ++s) t[size_t(j) * maxS + s] = j;
    CKE(cudaMemcpyAsync(L.tidx, t.data(), t.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    CKE(cudaStreamSynchronize(st));
  }
  if (p2p) {
    const int mrows = (a.map_w != nullptr && a.map_rows > 1) ? a.map_rows : 1;
    L.mapper = reinterpret_cast<int*>(tp.get(size_t(B) * mrows * 80 * sizeof(int)));
    if (a.map_w) {
      L.map_w = fa(size_t(B) * mrows * 80);
      CKE(cudaMemcpyAsync(L.map_w, a.map_w, size_t(B) * mrows * 80 * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    L.is_replace = reinterpret_cast<int*>(tp.get(size_t(B) * sizeof(int)));
    L.c_base = fa(size_t(T + 1) * B * 80); L.c_tar = fa(size_t(T + 1) * B * 80);
    CKE(cudaMemcpyAsync(L.mapper, a.mapper, size_t(B) * mrows * 80 * sizeof(int), cudaMemcpyHostToDevice, st));
    CKE(cudaMemcpyAsync(L.is_replace, a.is_replace, size_t(B) * sizeof(int), cudaMemcpyHostToDevice, st));
    CKE(cudaMemcpyAsync(L.c_base, a.c_base, size_t(T + 1) * B * 80 * sizeof(float), cudaMemcpyHostToDevice, st));
    CKE(cudaMemcpyAsync(L.c_tar, a.c_tar, size_t(T + 1) * B * 80 * sizeof(float), cudaMemcpyHostToDevice, st));
    if (a.replace_m) {
      L.replace_m = fa(size_t(B) * 77 * 80);
      CKE(cudaMemcpyAsync(L.replace_m, a.replace_m, size_t(B) * 77 * 80 * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    if (blend) {
      L.has_blend = reinterpret_cast<int*>(tp.get(size_t(B) * sizeof(int)));
      L.blend_alpha = fa(size_t(B) * brows * 80);
      const size_t accn = size_t(B) * brows * E.n_blend_layers() * c.heads * 256;
      CKE(cudaMemcpyAsync(L.has_blend, a.has_blend, size_t(B) * sizeof(int), cudaMemcpyHostToDevice, st));
      CKE(cudaMemcpyAsync(L.blend_alpha, a.blend_alpha, size_t(B) * brows * 80 * sizeof(float), cudaMemcpyHostToDevice, st));
      if (a.blend_state) {
        L.blend_acc = a.blend_state;               // caller-owned, carried across single-step calls
      } else {
        L.blend_acc = fa(accn);
        CKE(cudaMemsetAsync(L.blend_acc, 0, accn * sizeof(float), st));
      }
    }
  }
  // ---- inputs
  if (E.set_contexts(a.ctx, 1 + 2 * B, st)) return -1;
  if (E.set_timesteps(a.timesteps, T + 1, st)) return -1;
  CKE(cudaMemcpyAsync(L.zs, a.zs, size_t(B) * T * n * sizeof(float), kIn, st));
  if (a.xt_is_pair) {
    CKE(cudaMemcpyAsync(L.lat, a.xT, size_t(2) * B * n * sizeof(float), kIn, st));
  } else {   // xt rows 0 and 1 <- xT
    CKE(cudaMemcpy2DAsync(L.lat, size_t(2) * n * sizeof(float), a.xT, size_t(n) * sizeof(float), size_t(n) * sizeof(float), B, kIn, st));
    CKE(cudaMemcpy2DAsync(L.lat + n, size_t(2) * n * sizeof(float), a.xT, size_t(n) * sizeof(float), size_t(n) * sizeof(float), B, kIn, st));
  }

  uint32_t mask_small = 0;    // transformer blocks whose token count allows self-attention replacement
  for (int i = 0; i < E.n_tf(); ++i) if (E.tf_tokens(i) <= a.self_max_tokens) mask_small |= 1u << i;

  long launches = 0;
  long long fwd = 0;
  float* xt = L.lat; float* xprev = L.lat + size_t(2) * B * n; float* xopt = L.lat + size_t(4) * B * n;

  int masa_step = 0;      // MasaCtrl's editor counts every controlled launch (masactrl_utils.py:15-23)
  auto run_call = [&](CallDesc& cd, int tindex, int ctrl_step, bool save) -> int {
    dim3 g(std::max(1, n / 4 / 256), cd.S);
    gather_latents_kernel<<<g, 256, 0, st>>>(L.lat, cd.d_lat, L.xin, n / 4);
    ++launches;
    CallCtrl cc;
    CKE(cudaMemcpyAsync(L.tidx_cur, L.tidx + size_t(tindex) * maxS, size_t(cd.S) * sizeof(int), cudaMemcpyDeviceToDevice, st));
    cc.ctx_idx = cd.d_ctx; cc.time_idx = L.tidx_cur;
    static const bool prefix_dedup = !(getenv("HEDIT_PREFIX_DEDUP") && atoi(getenv("HEDIT_PREFIX_DEDUP")) == 0);
    if (prefix_dedup && E.prefix_dedup() && cd.n_uniq < cd.S) { cc.uniq_first = cd.d_ufirst; cc.uniq_of = cd.d_uof; cc.n_uniq = cd.n_uniq; }   // one timestep per launch here
    cc.unit_s0 = cd.d_us0; cc.unit_s1 = cd.d_us1; cc.unit_img = cd.d_uimg; cc.n_units = cd.n_units;
    if (cd.p2p && p2p) {
      if (a.self_lo <= a.ctrl_step0 + ctrl_step && a.ctrl_step0 + ctrl_step < a.self_hi) { cc.self_mask = mask_small; cc.self_q = cd.d_sq; cc.self_k = cd.d_sq; cc.self_v = nullptr; }
      cc.mapper = L.mapper; cc.is_replace = L.is_replace; cc.replace_m = L.replace_m;
      cc.map_w = L.map_w; cc.map_rows = (a.map_w != nullptr && a.map_rows > 1) ? a.map_rows : 1;
      CKE(cudaMemcpyAsync(L.c_base_cur, L.c_base + size_t(ctrl_step) * B * 80, size_t(B) * 80 * sizeof(float), cudaMemcpyDeviceToDevice, st));
      CKE(cudaMemcpyAsync(L.c_tar_cur, L.c_tar + size_t(ctrl_step) * B * 80, size_t(B) * 80 * sizeof(float), cudaMemcpyDeviceToDevice, st));
      cc.c_base = L.c_base_cur; cc.c_tar = L.c_tar_cur;
      if (blend && save) { cc.blend_acc = L.blend_acc; cc.blend_alpha = L.blend_alpha; cc.blend_rows = brows; }
    } else if (cd.p2p && pnp) {
      if (a.pnp_qk_on[ctrl_step]) { cc.self_mask = a.pnp_self_mask; cc.self_q = cd.d_sq; cc.self_k = cd.d_sq; cc.self_v = nullptr; }
      if (a.pnp_feat_on[ctrl_step]) cc.feat_src = cd.d_sq;
    } else if (cd.p2p && masa && a.masa_step_on[masa_step]) {
      cc.self_mask = a.masa_layer_mask; cc.self_q = nullptr; cc.self_k = cd.d_sk; cc.self_v = cd.d_sk;
    }
    const long r = E.forward_replayed(L.xin, L.eps + size_t(cd.pool_off) * n, cd.S, cc, st);
    if (r < 0) return -1;
    launches += r;
    fwd += cd.S;
    if (cd.p2p && masa) ++masa_step;
    return 0;
  };

  if (a.pre_step) {
    // h_Edit_R_implicit after skipped steps (p2p_h_edit.py:239-267): one editing move of the edit row at the first timestep
    if (a.variant != 1 || a.explicit_form) { E.err_ = "pre_step belongs to h_Edit_R_implicit (variant 1, implicit form)"; return -1; }
    CKE(cudaMemcpy2DAsync(xopt, size_t(n) * sizeof(float), xt + n, size_t(2) * n * sizeof(float), size_t(n) * sizeof(float), B, cudaMemcpyDeviceToDevice, st));
    if (run_call(calls[1], 0, 0, false)) return -1;
    CorrParams cp;
    cp.eps = L.eps; cp.iu = L.iu; cp.ics = L.ics; cp.ict = L.ict; cp.w_src_edit = a.w_src_edit; cp.w_tar = a.w_tar;
    cp.corr = L.corr; cp.x_opt = xopt; cp.x_stride = n; cp.x_base = xopt; cp.xb_stride = n; cp.partial = nullptr; cp.n = n;
    hstep_corr_kernel<<<dim3(nparts, B), 256, 0, st>>>(cp);
    UpdateParams up;
    up.x_opt = xopt; up.x_stride = n; up.x_base = xopt; up.xb_stride = n; up.corr = L.corr;
    up.partial = nullptr; up.nparts = nparts; up.coeff = a.pre_coeff; up.w_rec = 0.f; up.n = n;
    hstep_update_kernel<<<dim3(nparts, B), 256, 0, st>>>(up);
    launches += 2;
    CKE(cudaMemcpy2DAsync(xt + n, size_t(2) * n * sizeof(float), xopt, size_t(n) * sizeof(float), size_t(n) * sizeof(float), B, cudaMemcpyDeviceToDevice, st));
  }

  for (int i = 0; i < T; ++i) {
    const int idx = T - 1 - i;                     // zs index (p2p_h_edit.py:599)
    NvtxRange nvtx_step_("hedit.edit.step %d of %d", i, T);
    const hedit_step_coef& hc = a.coef[i];
    // ---- call A (or the single explicit-form call) at t
    if (run_call(calls[0], i, i, true)) return -1;
    ReverseParams rp;
    rp.xt = xt; rp.z = L.zs + size_t(idx) * n; rp.z_stride = size_t(T) * n;
    rp.eps_u = L.eps; rp.eps_c = L.eps;
    rp.iu = (i == 0) ? L.iuA0 : L.iuA; rp.ic = (i == 0) ? L.icA0 : L.icA;
    rp.w_src = a.w_src;
    rp.k.sqrt_1m_at = hc.sqrt_1m_at; rp.k.sqrt_at = hc.sqrt_at; rp.k.sqrt_ap = hc.sqrt_ap; rp.k.dir = hc.dir; rp.k.noise = hc.noise; rp.k.coeff = hc.coeff;
    rp.x_prev = xprev; rp.n = n;
    rp.per_row = 0; rp.w_row1 = a.w_src; rp.k1 = rp.k;
    if (baseline) {
      const hedit_step_coef& he = a.coef_edit ? a.coef_edit[i] : hc;
      rp.per_row = 1; rp.w_row1 = a.w_tar;
      rp.k1.sqrt_1m_at = he.sqrt_1m_at; rp.k1.sqrt_at = he.sqrt_at; rp.k1.sqrt_ap = he.sqrt_ap; rp.k1.dir = he.dir; rp.k1.noise = he.noise; rp.k1.coeff = he.coeff;
    }
    hstep_reverse_kernel<<<dim3(std::max(1, n / 4 / 256), B, 2), 256, 0, st>>>(rp);
    ++launches;
    // x_opt <- x_base
    CKE(cudaMemcpy2DAsync(xopt, size_t(n) * sizeof(float), xprev + n, size_t(2) * n * sizeof(float), size_t(n) * sizeof(float), B, cudaMemcpyDeviceToDevice, st));
    for (int k = 0; k < (baseline ? 0 : K); ++k) {
      const bool save = (k == K - 1);
      if (!a.explicit_form) {
        if (a.variant == 0 && a.schedule == 0 && !pnp) { if (run_call(calls[1], i + 1, i, false)) return -1; if (run_call(calls[2], i + 1, i, save)) return -1; }
        else if (run_call(calls[1], i + 1, i, save)) return -1;
      }
      CorrParams cp;
      cp.eps = L.eps; cp.iu = L.iu; cp.ics = L.ics; cp.ict = L.ict; cp.w_src_edit = a.w_src_edit; cp.w_tar = a.w_tar;
      cp.corr = L.corr; cp.x_opt = xopt; cp.x_stride = n; cp.x_base = xprev + n; cp.xb_stride = size_t(2) * n;
      const bool pull = (k > 0) && a.mos_pull != 0;
      cp.partial = pull ? L.partial : nullptr; cp.n = n;
      hstep_corr_kernel<<<dim3(nparts, B), 256, 0, st>>>(cp);
      UpdateParams up;
      up.x_opt = xopt; up.x_stride = n; up.x_base = xprev + n; up.xb_stride = size_t(2) * n; up.corr = L.corr;
      up.partial = pull ? L.partial : nullptr; up.nparts = nparts; up.coeff = hc.coeff; up.w_rec = a.weight_reconstruction; up.n = n;
      hstep_update_kernel<<<dim3(nparts, B), 256, 0, st>>>(up);
      launches += 2;
      if (a.guidance) {
        // reward-guided Langevin move on the Tweedie prediction of the just-updated x_opt (h_edit.py:150-172); eps_tar is the
        // target-guided noise of THIS iteration's UNet call (evaluated before the text move), as in the reference
        const float s1m = a.x0_coef[2 * i], sa = a.x0_coef[2 * i + 1];
        X0PredParams xp;
        xp.eps = L.eps; xp.iu = L.iu; xp.ict = L.ict; xp.w_tar = a.w_tar; xp.sqrt_1m_att = s1m; xp.sqrt_att = sa;
        xp.x_opt = xopt; xp.x_stride = n; xp.x0 = a.guid_x0; xp.n = n;
        hstep_x0pred_kernel<<<dim3(nparts, B), 256, 0, st>>>(xp);
        if (a.guidance(a.guidance_user, i, k) != 0) { E.err_ = "guidance callback failed"; return -1; }
        GuidNormParams gn;
        gn.corr = L.corr; gn.grad_x0 = a.guid_grad; gn.inv_sqrt_att = 1.f / sa; gn.partial = L.partial; gn.n = n;
        hstep_guid_norm_kernel<<<dim3(nparts, B), 256, 0, st>>>(gn);
        GuidUpdateParams gu;
        gu.x_opt = xopt; gu.x_stride = n; gu.grad_x0 = a.guid_grad; gu.inv_sqrt_att = 1.f / sa; gu.weight = a.guidance_weight;
        gu.partial = L.partial; gu.nparts = nparts; gu.n = n;
        hstep_guid_update_kernel<<<dim3(nparts, B), 256, 0, st>>>(gu);
        launches += 3;
      }
    }
    // xt <- [x_orig_{t-1}, x_opt]
    CKE(cudaMemcpy2DAsync(xt, size_t(2) * n * sizeof(float), xprev, size_t(2) * n * sizeof(float), size_t(n) * sizeof(float), B, cudaMemcpyDeviceToDevice, st));
    CKE(cudaMemcpy2DAsync(xt + n, size_t(2) * n * sizeof(float), xopt, size_t(n) * sizeof(float), size_t(n) * sizeof(float), B, cudaMemcpyDeviceToDevice, st));
    if (blend && (a.ctrl_step0 + i + 1) > a.start_blend) {
      BlendParams bp;
      bp.acc = L.blend_acc; bp.has_blend = L.has_blend; bp.L = E.n_blend_layers(); bp.H = c.heads; bp.th = a.blend_th;
      bp.rows = brows; bp.th_sub = a.blend_th_sub;
      bp.xt = xt; bp.C = c.in_ch; bp.hh = c.sample; bp.ww = c.sample;
      local_blend_kernel<<<B, 256, 0, st>>>(bp);
      ++launches;
    }
    if (a.trace) CKE(cudaMemcpyAsync(a.trace + size_t(i) * B * 2 * n, xt, size_t(B) * 2 * n * sizeof(float), kOut, st));
  }
  CKE(cudaMemcpy2DAsync(a.edited, size_t(n) * sizeof(float), xt + n, size_t(2) * n * sizeof(float), size_t(n) * sizeof(float), B, kOut, st));
  CKE(cudaMemcpy2DAsync(a.recon, size_t(n) * sizeof(float), xt, size_t(2) * n * sizeof(float), size_t(n) * sizeof(float), B, kOut, st));
  CKE(cudaStreamSynchronize(st));
  CKE(cudaGetLastError());
  a.n_sample_forwards = fwd;
  a.n_kernel_launches = launches;
  return 0;
}

}  // namespace hedit
