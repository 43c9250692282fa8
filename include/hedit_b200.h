/*
 * hedit_b200 -- C ABI of the B200-native h-Edit hot path (reverse-time bridge sampling loop).
 *
 * The reference (nktoan/h-edit) has no FFI: its boundary is Python duck typing.  Each entry point below names the
 * reference interface it replaces (paths relative to the reference root); INTEGRATION.md shows the ctypes binding
 * a reference maintainer would add.  All pointers are plain device or host addresses (cudaMemcpyDefault semantics
 * where noted); no torch types cross this boundary.  Every function returns 0 / a non-negative count on success
 * and a negative code on failure, with the reason available from hedit_last_error().
 */
#ifndef HEDIT_B200_H
#define HEDIT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hedit_engine hedit_engine;

/* SD-1.x UNet geometry (diffusers UNet2DConditionModel config the reference loads at
 * text-guided/main_p2p.py:106; restated in oracle/sd_unet.py). */
typedef struct hedit_unet_config {
  int32_t in_channels, out_channels, sample_size;
  int32_t block_out_channels[4];
  int32_t layers_per_block, heads, cross_attention_dim, norm_groups, ctx_len;
} hedit_unet_config;

/* Per-step scalars of reverse_step / compute_full_coeff (text-guided/inversion/inversion_utils.py:58-126,168-195
 * and the coeff line text-guided/inversion/p2p_h_edit.py:664-665), precomputed by the host. */
typedef struct hedit_step_coef {
  float sqrt_1m_at, sqrt_at, sqrt_ap, dir, noise, coeff;
} hedit_step_coef;

/* Reward-model hook of the guided samplers (text-guided-n-style/inversion/h_edit.py:150-172): called once per (timestep, MOS
 * iteration) on the launching stream with x0 = the Tweedie prediction [B][C][h][w] already written to guid_x0 (device); must
 * enqueue, on the same stream, work that leaves dLoss/dx0 [B][C][h][w] in guid_grad (device).  Returns 0 on success. */
typedef int (*hedit_guidance_fn)(void* user, int step, int opt_step);

/* One batched edit = B independent images, each with prompts [src, tar].
 * Replaces the loop-level callables h_Edit_p2p_implicit / h_Edit_p2p_explicit
 * (text-guided/inversion/p2p_h_edit.py:529,380) including the P2P controller hook surface
 * (text-guided/p2p/ptp_utils.py:31-123, p2p/ptp_classes.py:17-283) compiled into per-step tables. */
typedef struct hedit_edit_args {
  int32_t B;                 /* images */
  int32_t steps;             /* after_skip_steps (number of timesteps executed) */
  int32_t opt_steps;         /* optimization_steps K (implicit form) */
  int32_t explicit_form;     /* 0: h_Edit_p2p_implicit, 1: h_Edit_p2p_explicit */
  int32_t schedule;          /* 0: the reference's UNet call pattern (9 sample-forwards / step at K=1);
                                1: exact-reuse merged pattern (7 / step): source-branch outputs of call C are reused
                                   as the next step's call-A source inputs, calls B and C share one launch;
                                2: (opt-in, requires w_src == 1) additionally drops the unconditional forwards that only feed
                                   u + 1 * (c - u) == c (5 / step); equal to schedule 1 up to one fp32 rounding per element */
  int32_t buffers_on_host;   /* 1: xT, zs, ctx, edited, recon, trace are host pointers (copies are part of the call) */
  int32_t variant;           /* 0: P2P-family samplers (orig and edit rows both denoised: h_Edit_p2p_*, h_Edit_masactrl_implicit);
                                1: h_Edit_R_implicit / h_Edit_R_explicit (p2p_h_edit.py:162,21): no attention control, both rows are
                                   stepped with the edit row's source-guided noise prediction;
                                2: the baseline samplers ef_or_pnp_inv_w_p2p / ef_or_pnp_inv_w_masactrl (inversion/p2p_baselines.py:103,
                                   masactrl_baselines.py:15; Edit Friendly and PnP Inversion): one attention-controlled launch
                                   [xo,null] [xe,null] [xo,src] [xe,tar] per timestep, then the ORIG row steps with the source-guided
                                   noise (coef) and the EDIT row with the target-guided noise (w_tar, coef_edit); no h-term.  Accepts
                                   xt_is_pair / ctrl_step0 / blend_state for single-step use (the gradient-guided baselines nmg_p2p,
                                   nmg_pnp, nulltext_pnp call it once per timestep) */
  const float* xT;           /* [B][C][h][w] */
  const float* zs;           /* [B][steps][C][h][w]; zs[b][idx] as in the reference (idx = steps-1-i at step i) */
  const float* ctx;          /* [1+2B][ctx_len][cross_dim]: row 0 = "", then (src_b, tar_b) pairs (encode_text) */
  const float* timesteps;    /* [steps+1]: t_0 .. t_{steps-1}, then the final previous timestep (0) */
  const hedit_step_coef* coef; /* [steps] (host) */
  float w_src, w_src_edit, w_tar;   /* cfg_scales */
  float weight_reconstruction;
  /* ---- Prompt-to-Prompt tables (host pointers); use_p2p = 0 runs every UNet call with use_controller=False */
  int32_t use_p2p;
  const int32_t* mapper;     /* [B][map_rows][80]  source-token index lists: row 0 = AttentionRefine.mapper (clamped to [0,77)) */
  const float* map_w;        /* [B][map_rows][80] or NULL: weights of those indices.  With map_w the cross edit's base term is
                                sum_k map_w[k][j] * P_src[mapper[k][j]]: the non-zeros of AttentionReplace's 77x77 mapper, column by column
                                (seq_aligner.py:157-190: 1-3 per column), instead of the dense product with replace_m */
  int32_t map_rows;          /* R (<= 4 served from shared memory); 0 / 1 without map_w */
  const int32_t* is_replace; /* [B]      1: AttentionReplace matrix form */
  const float* replace_m;    /* [B][77][80] or NULL */
  const float* c_base;       /* [steps+1][B][80]  coefficient on the mapped source probability at controller step s */
  const float* c_tar;        /* [steps+1][B][80]  coefficient on the target's own probability */
  int32_t self_lo, self_hi;  /* AttentionControlEdit.num_self_replace */
  int32_t self_max_tokens;   /* 32*32 (ptp_classes.py:196) */
  const int32_t* has_blend;  /* [B] or NULL */
  const float* blend_alpha;  /* [B][blend_rows][80]  rows 0,1 = LocalBlend.alpha_layers (src, tar); rows 2,3 = substruct_layers */
  int32_t start_blend;       /* LocalBlend.start_blend */
  float blend_th;            /* LocalBlend.th[0]: threshold of the max-pooled word mask */
  int32_t blend_rows;        /* 0 / 2: no substruct_words; 4: mask *= ~mask(substruct word maps, no pooling, th[1]) (ptp_classes.py:28-38,66-67) */
  float blend_th_sub;        /* LocalBlend.th[1] */
  /* ---- MasaCtrl mutual self-attention (masactrl/masactrl.py:11-69).  masa = 1: in the c-th attention-controlled UNet launch of this
   * call (the editor's cur_step advances once per controlled launch, masactrl_utils.py:15-23, i.e. steps * opt_steps launches in the
   * implicit form) the edit samples attend to the K/V of their source samples in the transformer blocks of masa_layer_mask
   * (bit l = block l in forward order = the editor's layer_idx) iff masa_step_on[c] != 0 (= `cur_step in step_idx`, which also ends
   * the injection once cur_step reaches total_steps).  masa = 0: off */
  int32_t masa;
  uint32_t masa_layer_mask;
  const int32_t* masa_step_on;   /* host [steps * opt_steps] */
  int32_t mos_pull;          /* 1: apply the L1 reconstruction pull on MOS iterations k>0 (p2p_h_edit.py:670-686); 0: masactrl_h_edit.py */
  /* ---- Plug-and-Play (text-guided/plug_n_play/pnp_utils.py:29-164 driven by inversion/pnp_h_edit.py:33): pnp = 1 runs
   * h_Edit_PnP_implicit.  The attention-controlled call of a step is the pair ([x_orig,src],[x_opt,tar]) at the previous timestep tt;
   * when pnp_qk_on[i] != 0 (tt in the qk injection schedule) the target takes the source's self-attention q and k in the transformer
   * blocks of pnp_self_mask (bit = block index in forward order), and when pnp_feat_on[i] != 0 it takes the source's conv2 output at
   * up_blocks[1].resnets[1].  Requires explicit_form = 0, use_p2p = 0 and variant = 0 -- or variant = 2 (the Plug-and-Play baselines
   * ef_or_pnp_inv_w_pnp / negative_prompt_pnp / nmg_pnp / nulltext_pnp, inversion/pnp_baselines.py), where the injected pair is samples
   * 2, 3 of the step's single launch at the CURRENT timestep and the flags are indexed by that timestep. */
  int32_t pnp;
  uint32_t pnp_self_mask;
  const int32_t* pnp_qk_on;   /* host [steps] */
  const int32_t* pnp_feat_on; /* host [steps] */
  /* ---- h_Edit_R_implicit on a skipped schedule (p2p_h_edit.py:214-267): pre_step = 1 first moves the edit row at the first
   * executed timestep by pre_coeff * (eps_tar - eps_src_edit) evaluated at that timestep (3 extra sample-forwards per image);
   * pre_coeff = compute_full_coeff(time_ahead, t) - sqrt(1-abar[time_ahead]) * sqrt(abar[t]) / sqrt(abar[time_ahead]).
   * Requires variant = 1, explicit_form = 0. */
  int32_t pre_step;
  float pre_coeff;
  /* ---- reward guidance: after the text-guided move of every MOS iteration, x_opt <- x_opt - rho * dLoss/dx with
   * dLoss/dx = guid_grad / sqrt(abar_tt), rho = rms(corr) / rms(dLoss/dx) * guidance_weight per image (h_edit.py:150-172).
   * guidance = NULL: off.  x0_coef[steps][2] (host) = (sqrt(1 - abar_tt), sqrt(abar_tt)) of every step's previous timestep. */
  hedit_guidance_fn guidance;
  void* guidance_user;
  float guidance_weight;
  const float* x0_coef;
  float* guid_x0;            /* device [B][C][h][w] */
  float* guid_grad;          /* device [B][C][h][w] */
  /* ---- single-step use (h_edit_step): run `steps` timesteps of a longer schedule and carry the controller state outside */
  int32_t xt_is_pair;        /* 1: xT is [B][2][C][h][w] = (x_orig, x_edit) rows of an edit in progress (requires schedule 0) */
  int32_t ctrl_step0;        /* controller step (AttentionControl.cur_step / LocalBlend.counter) before the first executed timestep;
                                c_base / c_tar row 0 corresponds to controller step ctrl_step0 */
  float* blend_state;        /* device [B][2][n_blend_layers][heads][256] accumulated word maps carried across calls, or NULL */
  /* ---- outputs */
  float* edited;             /* [B][C][h][w] */
  float* recon;              /* [B][C][h][w] */
  float* trace;              /* [steps][B][2][C][h][w] or NULL: xt after every timestep */
  int64_t n_sample_forwards; /* out */
  int64_t n_kernel_launches; /* out */
  /* ---- variant 2 only: reverse-step scalars of the EDIT row (host [steps]); NULL = coef.  PnP Inversion steps the edit row with
   * eta = 0 while the orig row keeps eta = 1 (p2p_baselines.py:180-184) */
  const hedit_step_coef* coef_edit;
} hedit_edit_args;

const char* hedit_last_error(void);
/* sizeof() of the argument / configuration structs of this header by name ("hedit_edit_args", "hedit_face_args", "hedit_unet_config",
 * ...), -1 for an unknown name: lets a binding (ctypes, cgo, JNI) verify its struct layout against the library it loaded. */
int hedit_abi_sizeof(const char* struct_name);
int hedit_device_count(void);

/* engine lifetime: replaces copy.deepcopy(pipeline).to(device) per image (text-guided/main_p2p.py:119) with
 * persistent device weights */
hedit_engine* hedit_engine_create(const hedit_unet_config* cfg, int max_samples, int max_contexts, int device);
void hedit_engine_destroy(hedit_engine* e);
/* state-dict ingestion by diffusers parameter name (fp32, host or device pointer) */
int hedit_engine_load_tensor(hedit_engine* e, const char* name, const float* data, const int64_t* dims, int ndim);
int hedit_engine_finalize(hedit_engine* e);
double hedit_engine_flops_per_sample(hedit_engine* e);
/* number of floats of hedit_edit_args.blend_state for a batch of B images */
int hedit_engine_blend_state_elems(hedit_engine* e, int B);
/* "fp16" (default) or "bf16": the 16-bit tensor-core operand type this build uses for activations/weights */
const char* hedit_operand_dtype(void);
/* enumerate the state-dict tensors the engine expects (diffusers parameter names): returns ndim, fills dims4 */
int hedit_engine_tensor_count(hedit_engine* e);
int hedit_engine_tensor_info(hedit_engine* e, int index, char* name_buf, int name_len, int64_t* dims4);
/* The sampling loop replays each distinct UNet launch (same buffers, batch and control tables) from a CUDA graph from its third
 * occurrence on (on by default; HEDIT_LOOP_GRAPH=0 or on=0 launches every kernel directly).  Results are bit-identical either way. */
int hedit_engine_set_graph_replay(hedit_engine* e, int on);
/* Samples of one UNet launch that share a latent and differ only in their text context (the [x,null] [x,src] [x,tar] evaluations of
 * classifier-free guidance, p2p_h_edit.py:606-652) compute identical activations up to the first cross-attention; the loop evaluates that
 * context-free prefix (conv_in, down_blocks[0].resnets[0], the first transformer block's self-attention) once per distinct latent and
 * broadcasts it.  On by default (HEDIT_PREFIX_DEDUP=0 or on=0: every sample evaluates it).  Bit-identical results either way. */
int hedit_engine_set_prefix_dedup(hedit_engine* e, int on);
/* split-K for the launches of the deep UNet levels at 1-5 samples (8 / 20 tiles on 148 SMs otherwise): K is cut into <= 8 ranges whose
 * fp32 partial tiles are added in a fixed order by a second kernel that applies the epilogue.  Off by default -- with it the low bits
 * of a result depend on the batch size; the one-image samplers (h_Edit_p2p_implicit(model, xT, ...), main_p2p.py:224) switch it on. */
int hedit_engine_set_splitk(hedit_engine* e, int on);
/* diagnostics: one UNet forward of S samples with CUDA events around every kernel; writes "tag:ms:launches;" records */
int hedit_engine_profile_forward(hedit_engine* e, int S, int reps, char* out, int out_len);

/* model.unet(sample, t, encoder_hidden_states=ctx, cross_attention_kwargs={'use_controller': False}).sample
 * (text-guided/inversion/p2p_h_edit.py:613).  x/eps: [S][C][h][w] device fp32; timesteps: [S] host; ctx:
 * [S][ctx_len][cross_dim] host or device.  stream: cudaStream_t or NULL. */
int hedit_unet_forward(hedit_engine* e, const float* x, const float* timesteps, const float* ctx, int S, float* eps, void* stream);
/* same, with n_ctx distinct contexts shared by the samples through ctx_idx[S] (host): the inversion callables
 * (text-guided/inversion/ddpm_inversion.py:130-132, ddim_inversion.py:31-52) evaluate many timesteps against two prompts */
int hedit_unet_forward_indexed(hedit_engine* e, const float* x, const float* timesteps, const float* ctx, int n_ctx, const int32_t* ctx_idx,
                               int S, float* eps, void* stream);

/* Compatibility path for ARBITRARY controller objects: the same UNet call with every attention layer run the reference processor's
 * way (text-guided/p2p/ptp_utils.py:88-107): scores -> fp32 softmax, MATERIALISED -> probs_hook -> probs . V.  The hook is invoked on the
 * calling thread, in layer order (attn1 then attn2 of each transformer block, down -> mid -> up = the order in which the reference's
 * processors fire), with the device buffer probs[(S*heads)][n_query][n_key] (head_to_batch_dim order: sample-major) which it may edit
 * in place with work queued on `stream`; place: 0 "down", 1 "mid", 2 "up"; return non-zero to abort.  This is what a maintainer binds
 * `P2PCrossAttnProcessor.__call__`'s `self.controller(attention_probs, is_cross, self.place_in_unet, save_attn)` line to when the
 * controller is not one of the stock classes (which compile to the fused kernels instead).  Slow by construction (CUDA-core kernels,
 * 2.1 GB of probabilities for 4 samples at 64x64). */
typedef int (*hedit_attn_probs_fn)(void* user, int tf_index, int is_cross, int place, float* probs, int batch_heads, int n_query, int n_key);
int hedit_unet_forward_compat(hedit_engine* e, const float* x, const float* timesteps, const float* ctx, int S, float* eps,
                              hedit_attn_probs_fn probs_hook, void* user, void* stream);

/* The same for MasaCtrl's EDITOR protocol (text-guided/masactrl/masactrl_utils.py:40-89: `out = editor(q, k, v, sim, attn, is_cross,
 * place_in_unet, self.heads, scale=self.scale)` replaces `attn @ v`): per attention layer the hook receives, as fp32 device buffers,
 * q [(S*heads)][n_query][d], k and v [(S*heads)][n_key][d] (the 'b n (h d) -> (b h) n d' split), sim = q k^T * d^-1/2 and attn =
 * softmax(sim), both [(S*heads)][n_query][n_key], and must write the layer output [S][n_query][heads*d] ('(b h) n d -> b n (h d)') into
 * `out` with work queued on `stream`.  Serves editor objects that are not the stock MutualSelfAttentionControl. */
typedef int (*hedit_attn_editor_fn)(void* user, int tf_index, int is_cross, int place, float* q, float* k, float* v, float* sim, float* attn,
                                    float* out, int batch_heads, int n_query, int n_key, int d);
int hedit_unet_forward_editor(hedit_engine* e, const float* x, const float* timesteps, const float* ctx, int S, float* eps,
                              hedit_attn_editor_fn editor_hook, void* user, void* stream);

/* the whole bridge-sampling loop for a batch of images */
int hedit_edit_p2p(hedit_engine* e, hedit_edit_args* args, void* stream);

/* ---- VAE decoder: `model.vae.decode(z).sample` (diffusers AutoencoderKL decode direction; text-guided/main_p2p.py:262-275) and the
 * gradient of a scalar loss on the decoded image with respect to z, which the reference's style path obtains with torch.autograd
 * through that decode (text-guided-n-style/inversion/h_edit.py:158-164).  Weights are ingested by diffusers parameter name
 * ("post_quant_conv.weight", "decoder.conv_in.weight", ...), fp32, host or device pointer. */
typedef struct hedit_vae hedit_vae;
typedef struct hedit_vae_config {
  int32_t latent_channels, out_channels;
  int32_t block_out_channels[4];
  int32_t layers_per_block, norm_groups;
} hedit_vae_config;
hedit_vae* hedit_vae_create(const hedit_vae_config* cfg, int device);
void hedit_vae_destroy(hedit_vae* v);
int hedit_vae_load_tensor(hedit_vae* v, const char* name, const float* data, const int64_t* dims, int ndim);
int hedit_vae_finalize(hedit_vae* v);
int hedit_vae_tensor_count(hedit_vae* v);
int hedit_vae_tensor_info(hedit_vae* v, int index, char* name_buf, int name_len, int64_t* dims4);
/* z [B][latent][h][w] (device fp32, already divided by the scaling factor 0.18215) -> img [B][out][8h][8w] (device fp32).
 * Returns the number of kernels launched; keeps the activations decode_backward needs until the next decode. */
int hedit_vae_decode(hedit_vae* v, const float* z, float* img, int B, int h, int w, void* stream);
/* dimg [B][out][8h][8w] = dLoss/dimg of the last decode -> dz [B][latent][h][w] = dLoss/dz (device fp32) */
int hedit_vae_decode_backward(hedit_vae* v, const float* dimg, float* dz, void* stream);
/* floating-point operations of the last decode (+ backward) call, for roofline accounting */
double hedit_vae_last_flops(hedit_vae* v);

/* VAE encoder: `model.vae.encode(x).latent_dist` (text-guided/main_p2p.py:154-159); weights "encoder.*" and "quant_conv.*" */
typedef struct hedit_vae_enc hedit_vae_enc;
hedit_vae_enc* hedit_vae_enc_create(const hedit_vae_config* cfg, int device);
void hedit_vae_enc_destroy(hedit_vae_enc* v);
int hedit_vae_enc_load_tensor(hedit_vae_enc* v, const char* name, const float* data, const int64_t* dims, int ndim);
int hedit_vae_enc_finalize(hedit_vae_enc* v);
/* img [B][3][H][W] (device fp32, in [-1,1]) -> moments [B][2*latent][H/8][W/8] = (mean | logvar) (device fp32); latent_dist.mode() = mean */
int hedit_vae_encode(hedit_vae_enc* v, const float* img, float* moments, int B, int H, int W, void* stream);

/* ---- CLIP-Gram style reward: `torch.linalg.norm(image_encoder.get_gram_matrix_residual(img))` and its gradient with respect to
 * img (text-guided-n-style/clip_guidance/base_clip.py:55-66 over clip/model.py:339-359; differentiated by torch.autograd at
 * text-guided-n-style/inversion/h_edit.py:161-164).  Weights = state dict of the CLIP image tower (`clip_model.visual`: "conv1.weight",
 * "class_embedding", "positional_embedding", "ln_pre.*", "transformer.resblocks.{i}.*" for i < layers), fp32.  ViT variants with head
 * dim 64 and at most 256 tokens (ViT-B/16 at 224 px: 197). */
typedef struct hedit_clip hedit_clip;
typedef struct hedit_clip_config {
  int32_t resolution, patch, width, heads, layers;   /* layers = blocks evaluated: features[2] -> 3 */
} hedit_clip_config;
hedit_clip* hedit_clip_create(const hedit_clip_config* cfg, int device);
void hedit_clip_destroy(hedit_clip* c);
/* returns 1 (ignored) for tensors this path does not use (later blocks, ln_post, proj) */
int hedit_clip_load_tensor(hedit_clip* c, const char* name, const float* data, const int64_t* dims, int ndim);
int hedit_clip_finalize(hedit_clip* c);
/* ref [1][3][resolution][resolution]: the CLIP-normalised style image (CLIPEncoder.ref, base_clip.py:43-52), device fp32 */
int hedit_clip_set_reference(hedit_clip* c, const float* ref, void* stream);
/* img [B][3][H][W] in [-1,1] (device fp32) -> loss[B] (device fp32); keeps the activations gram_backward needs */
int hedit_clip_gram_loss(hedit_clip* c, const float* img, int B, int H, int W, float* loss, void* stream);
/* dimg [B][3][H][W] = d loss[b] / d img[b] of the last gram_loss call */
int hedit_clip_gram_backward(hedit_clip* c, float* dimg, void* stream);

/* ---- CLIP text tower: `model.text_encoder(ids)[0]` as called by encode_text (text-guided/inversion/inversion_utils.py:13-36).
 * Weights by transformers' CLIPTextModel parameter names ("text_model.embeddings.token_embedding.weight", ...); head dim must be 64. */
typedef struct hedit_text hedit_text;
typedef struct hedit_text_config { int32_t vocab, width, heads, layers, ffn, tokens; } hedit_text_config;
hedit_text* hedit_text_create(const hedit_text_config* cfg, int device);
void hedit_text_destroy(hedit_text* t);
int hedit_text_load_tensor(hedit_text* t, const char* name, const float* data, const int64_t* dims, int ndim);
int hedit_text_finalize(hedit_text* t);
/* ids [B][tokens] int32 (host) -> last_hidden_state [B][tokens][width] (device fp32) */
int hedit_text_encode(hedit_text* t, const int32_t* ids, int B, float* out, void* stream);

/* ---- face swapping: the pixel-space DDPM denoiser `model(x, t)` (face-swapping/diffusion/diffusion.py:193-341; weights by the
 * reference's parameter names: "temb.dense.0.weight", "conv_in.weight", "down.0.block.0.norm1.weight", ...) and the reward-guided
 * sampler `h_Edit_R` (face-swapping/inversion/h_edit_R.py:7-137). */
typedef struct hedit_face hedit_face;
typedef struct hedit_face_config {
  int32_t ch, n_levels, ch_mult[8], num_res_blocks, attn_resolution, image_size, in_channels, out_ch;
} hedit_face_config;
hedit_face* hedit_face_create(const hedit_face_config* cfg, int device);
void hedit_face_destroy(hedit_face* f);
int hedit_face_load_tensor(hedit_face* f, const char* name, const float* data, const int64_t* dims, int ndim);
int hedit_face_finalize(hedit_face* f);
int hedit_face_tensor_count(hedit_face* f);
int hedit_face_tensor_info(hedit_face* f, int index, char* name_buf, int name_len, int64_t* dims4);
/* eps = model(x, t): x, eps [S][3][R][R] device fp32; t [S] host.  Returns kernels launched. */
int hedit_face_unet_forward(hedit_face* f, const float* x, const float* t, int S, float* eps, void* stream);
/* floating-point operations (GEMM / conv) of the last denoiser call, for roofline accounting */
double hedit_face_last_flops(hedit_face* f);

/* reward hook of h_Edit_R: which = 0 -> idloss.get_cosine_loss, 1 -> lpipsloss.get_lpips_loss (arcface/arcface_model.py:62,91).
 * Called on the launching stream with the Tweedie prediction x0 [B][3][R][R] in reward_x0 (device); must leave d loss / d x0 (per
 * image) in reward_grad (device), enqueued on the same stream.  Returns 0 on success. */
typedef int (*hedit_reward_fn)(void* user, int which, int step, int opt_step);
typedef struct hedit_face_step_coef {
  float t, tm1;                 /* timestep and previous timestep (0 after the last step) */
  float sqrt_1m_at, sqrt_at;    /* of alpha_bar[t] */
  float sqrt_1m_atm1, sqrt_atm1;/* of alpha_bar[tm1] */
  float c2, noise;              /* c2 = sqrt(1-abar[tm1]) * sqrt(1 - 0.5^2);  noise = etas[idx] * sqrt(1-abar[tm1]) * 0.5  (h_edit_R.py:82-86) */
} hedit_face_step_coef;
typedef struct hedit_face_args {
  int32_t B, steps, opt_steps;
  const float* xT;              /* device [B][3][R][R] */
  const float* zs;              /* device [B][steps][3][R][R]; zs[b][idx], idx = steps-1-i at step i */
  const hedit_face_step_coef* coef;   /* host [steps] */
  float weight;                 /* weight_edit_face: rho = sqrt(abar[tm1]) * weight (h_edit_R.py:106) */
  const float* mask;            /* device [B][3][R][R] soft face mask or NULL (h_edit_R.py:113-116; applies to the ID move only) */
  int32_t use_id, use_lpips;
  hedit_reward_fn reward; void* reward_user;
  float* reward_x0;             /* device [B][3][R][R] */
  float* reward_grad;           /* device [B][3][R][R] */
  float* edited;                /* device [B][3][R][R] out */
  int64_t n_sample_forwards, n_kernel_launches;   /* out */
} hedit_face_args;
int hedit_face_edit(hedit_face* f, hedit_face_args* args, void* stream);

/* ---- face-swapping reward networks, forward + input gradient (no autograd) ------------------------------------
 * hedit_arcface replaces `IDLoss.get_cosine_loss` + `torch.autograd.grad` (face-swapping/arcface/arcface_model.py:41-70; IR-SE50
 * backbone arcface/facial_recognition/model_irse.py:9-56, helpers.py:47-125; call site inversion/h_edit_R.py:109-110).
 * Tensors are loaded under the state_dict keys of the reference's `Backbone(112, 50, mode='ir_se')` ("input_layer.0.weight",
 * "body.3.res_layer.4.running_var", "output_layer.3.weight", ...); BatchNorm folding happens in hedit_arcface_finalize.
 * hedit_lpips replaces `LPIPS_Loss.get_lpips_loss` + autograd (arcface_model.py:72-95; lpips package 0.1, net='vgg'; call site
 * h_edit_R.py:128-129).  Tensor names: "conv0".."conv12" .weight/.bias (the 13 VGG16 convs in order), "lin0".."lin4" .weight ([C]),
 * "shift", "scale" ([3], the ScalingLayer). */
typedef struct hedit_arcface hedit_arcface;
hedit_arcface* hedit_arcface_create(int device);
void hedit_arcface_destroy(hedit_arcface* a);
int hedit_arcface_load_tensor(hedit_arcface* a, const char* name, const float* data, const int64_t* dims, int ndim);
int hedit_arcface_finalize(hedit_arcface* a);
/* img [B][3][256][256] device fp32 in [-1, 1] -> unit-norm embedding [B][512] (device) */
int hedit_arcface_features(hedit_arcface* a, const float* img, int B, float* feat, void* stream);
/* the face whose identity is transferred (IDLoss.ref): img [1][3][256][256] device */
int hedit_arcface_set_reference(hedit_arcface* a, const float* img, void* stream);
/* loss[b] = 1 - cos(ref, f(img[b])) (device [B] or NULL); grad [B][3][256][256] = d loss[b] / d img[b].  Returns kernels launched. */
int hedit_arcface_loss_grad(hedit_arcface* a, const float* img, int B, float* loss, float* grad, void* stream);
double hedit_arcface_last_flops(hedit_arcface* a);
typedef struct hedit_lpips hedit_lpips;
hedit_lpips* hedit_lpips_create(int device);
void hedit_lpips_destroy(hedit_lpips* l);
int hedit_lpips_load_tensor(hedit_lpips* l, const char* name, const float* data, const int64_t* dims, int ndim);
int hedit_lpips_finalize(hedit_lpips* l);
/* anchor image(s) (LPIPS_Loss.src): img [n][3][R][R] device, n = 1 (shared) or the batch size; R = 128, 256 or 512 */
int hedit_lpips_set_source(hedit_lpips* l, const float* img, int n, int R, void* stream);
/* loss[b] = LPIPS(img[b], src) (device [B] or NULL); grad [B][3][R][R] = d loss[b] / d img[b].  Returns kernels launched. */
int hedit_lpips_loss_grad(hedit_lpips* l, const float* img, int B, float* loss, float* grad, void* stream);
double hedit_lpips_last_flops(hedit_lpips* l);

/* ---- single operators, exposed for parity tests (device pointers) ------------------------------------------- */
/* D[M][N] = A[M][K] W[N][K]^T (+bias) (+residual) -> fp32 and/or 16-bit; A, W in the operand dtype (hedit_operand_dtype) */
int hedit_op_linear(const void* A_h16, const void* W_h16, const float* bias, const float* residual, float* out_f32, void* out_h16,
                    int M, int N, int K, void* stream);
/* 3x3 conv, pad 1, stride 1 or 2, NHWC 16-bit in [S][Hin][Win][C], weights 16-bit [Cout][3][3][C] -> fp32 NHWC */
/* GEGLU feed-forward projection (diffusers GEGLU: hidden, gate = proj(x).chunk(2); hidden * gelu(gate)) with the gate fused into the GEMM
 * epilogue.  W [N2][K] and bias [N2] are in the engine's interleaved order: every 32 rows = 16 value rows followed by their 16 gate rows
 * (row 32c+i <- value row 16c+i, row 32c+16+i <- gate row 16c+i).  out [M][N2/2] 16-bit. */
int hedit_op_linear_geglu(const void* A_h16, const void* W_h16, const float* bias, void* out_h16, int M, int N2, int K, void* stream);
int hedit_op_conv3x3(const void* x_h16, const void* w_h16, const float* bias, float* out_f32, int S, int Hin, int Win, int C, int Cout,
                     int stride, void* stream);
/* softmax(Q K^T/sqrt(d)) V per (sample, head); q/k/v 16-bit [S][N][H*d] with row strides ldq/ldkv; idx arrays may be NULL */
int hedit_op_self_attention(const void* q, const void* k, const void* v, int ldq, int ldkv, int S, int Nq, int Nkv, int H, int d,
                            const int32_t* q_idx, const int32_t* k_idx, const int32_t* v_idx, void* out_h16, void* stream);
/* N x 77 cross attention with the fused Prompt-to-Prompt edit (replaces P2PCrossAttnProcessor.__call__ +
 * AttentionControlEdit.forward for is_cross=True: text-guided/p2p/ptp_utils.py:38-123, p2p/ptp_classes.py:202-283).
 * q 16-bit [S][Nq][H*d]; kv 16-bit [n_ctx][77][2*H*d] (K | V); work units = single samples (s1 = -1) or (source, target) pairs. */
int hedit_op_cross_attention_p2p(const void* q, const void* kv, int S, int n_ctx, int Nq, int H, int d, int n_units, const int32_t* unit_s0,
                                 const int32_t* unit_s1, const int32_t* unit_img, const int32_t* ctx_idx, const int32_t* mapper,
                                 const float* c_base, const float* c_tar, const float* replace_m, const int32_t* is_replace,
                                 float* blend_acc, const float* blend_alpha, int blend_layer, int n_blend_layers, void* out_h16, void* stream);
/* GroupNorm(32 groups)(+SiLU): fp32 NHWC [S][HW][C] -> 16-bit operand dtype */
int hedit_op_group_norm(const float* x, const float* gamma, const float* beta, void* out_h16, int S, int HW, int C, int groups, float eps,
                        int silu, void* stream);
int hedit_op_layer_norm(const float* x, const float* gamma, const float* beta, void* out_h16, int rows, int C, float eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HEDIT_B200_H */
