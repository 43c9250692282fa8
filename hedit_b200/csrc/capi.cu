// extern "C" boundary (include/hedit_b200.h).  Plain pointers and sizes only.
#include <cstring>
#include <mutex>
#include <map>
#include <string>
#include <vector>

#include "../../include/hedit_b200.h"
#include "elementwise.cuh"
#include "engine.h"
#include "tmap.h"
#include "vae.h"
#include "vae_enc.h"
#include "clip.h"
#include "clip_text.h"
#include "face.h"
#include "reward.h"

// Every entry point runs on its handle's device and restores the caller's current device on return: torch reads the current device from
// the runtime, so leaving it switched would silently redirect the caller's later allocations and launches.
struct DeviceGuard {
  int prev = -1;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) err = cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur = -1;
    cudaGetDevice(&cur);
    if (prev >= 0 && cur != prev) cudaSetDevice(prev);
  }
};

namespace hedit {
int run_edit(Engine& E, hedit_edit_args& a, cudaStream_t st);
int run_face_edit(FaceUNet& U, hedit_face_args& a, cudaStream_t st);
cudaError_t launch_gemm(const GemmParams& g, int bn, cudaStream_t st);
cudaError_t launch_self_attn(const AttnParams& a, int dch, int S, cudaStream_t st);
cudaError_t launch_cross_attn(const AttnParams& a, int dch, int units, cudaStream_t st);
int self_attn_bkv(int d, int Nq, int Nkv);
}  // namespace hedit

using namespace hedit;

static thread_local std::string g_err;
static int fail(const std::string& m, int code = -1) { g_err = m; return code; }
static int cuda_fail(cudaError_t e, const char* what) { return fail(std::string(what) + ": " + cudaGetErrorString(e)); }

struct hedit_engine {
  Engine* E;
  int device;
  int *d_ctx_idx = nullptr, *d_tidx = nullptr, *d_unit0 = nullptr, *d_unit1 = nullptr, *d_uimg = nullptr;
  int cap = 0;
};

struct hedit_face {
  FaceUNet* U;
  int device;
};

struct hedit_text {
  ClipText* T;
  int device;
};

struct hedit_clip {
  ClipGram* C;
  int device;
};

struct hedit_vae_enc {
  VaeEncoder* E;
  int device;
};

struct hedit_vae {
  VaeDecoder* D;
  int device;
};

extern "C" {

const char* hedit_last_error(void) { return g_err.c_str(); }

int hedit_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

hedit_engine* hedit_engine_create(const hedit_unet_config* cfg, int max_samples, int max_contexts, int device) {
  if (!cfg) { fail("null config"); return nullptr; }
  if (hedit_device_count() <= device) { fail("hedit_b200 requires a CUDA device (sm_100a); none visible"); return nullptr; }
  DeviceGuard guard_(device);
  if (guard_.err != cudaSuccess) { cuda_fail(guard_.err, "cudaSetDevice"); return nullptr; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) { fail("hedit_b200 kernels are built for sm_100a only; device is sm_" + std::to_string(prop.major * 10 + prop.minor)); return nullptr; }
  UNetCfg c;
  c.in_ch = cfg->in_channels; c.out_ch = cfg->out_channels; c.sample = cfg->sample_size;
  for (int i = 0; i < 4; ++i) c.boc[i] = cfg->block_out_channels[i];
  c.layers = cfg->layers_per_block; c.heads = cfg->heads; c.ctx_dim = cfg->cross_attention_dim; c.groups = cfg->norm_groups; c.ctx_len = cfg->ctx_len;
  if (c.in_ch != 4 || c.out_ch != 4 || c.layers != 2 || c.ctx_len != 77 || c.groups != 32) { fail("unsupported UNet config (need in/out 4, 2 layers per block, 77 tokens, 32 groups)"); return nullptr; }
  for (int i = 0; i < 4; ++i)
    if (c.boc[i] % 64 != 0 || (c.boc[i] / c.heads) % 8 != 0) { fail("block_out_channels must be multiples of 64 with head_dim % 8 == 0"); return nullptr; }
  if (c.ctx_dim % 8 != 0) { fail("cross_attention_dim must be a multiple of 8"); return nullptr; }
  hedit_engine* h = new hedit_engine();
  h->device = device;
  h->E = new Engine(c, max_samples, max_contexts);
  if (!h->E->ok()) { fail(h->E->error()); delete h->E; delete h; return nullptr; }
  h->cap = max_samples;
  cudaMalloc(&h->d_ctx_idx, max_samples * sizeof(int)); cudaMalloc(&h->d_tidx, max_samples * sizeof(int));
  cudaMalloc(&h->d_unit0, max_samples * sizeof(int)); cudaMalloc(&h->d_unit1, max_samples * sizeof(int)); cudaMalloc(&h->d_uimg, max_samples * sizeof(int));
  return h;
}

void hedit_engine_destroy(hedit_engine* h) {
  if (!h) return;
  DeviceGuard guard_(h->device);
  cudaFree(h->d_ctx_idx); cudaFree(h->d_tidx); cudaFree(h->d_unit0); cudaFree(h->d_unit1); cudaFree(h->d_uimg);
  delete h->E;
  delete h;
}

int hedit_engine_load_tensor(hedit_engine* h, const char* name, const float* data, const int64_t* dims, int ndim) {
  if (!h) return fail("null engine");
  DeviceGuard guard_(h->device);
  const int r = h->E->load_tensor(name, data, dims, ndim, 0);
  if (r) return fail(h->E->error(), r);
  return 0;
}

int hedit_engine_finalize(hedit_engine* h) {
  if (!h) return fail("null engine");
  std::string missing;
  const int r = h->E->finalize_weights(&missing);
  if (r) return fail(h->E->error(), -1);
  return 0;
}

double hedit_engine_flops_per_sample(hedit_engine* h) { return h ? h->E->flops_per_sample() : 0.0; }

int hedit_engine_blend_state_elems(hedit_engine* h, int B) {
  if (!h) return fail("null engine");
  return B * 4 * h->E->n_blend_layers() * h->E->cfg().heads * 256;      // sized for blend_rows = 4 (substruct words)
}

const char* hedit_operand_dtype(void) { return HEDIT_OPERAND_NAME; }

int hedit_engine_tensor_count(hedit_engine* h) { return h ? h->E->tensor_count() : fail("null engine"); }

int hedit_engine_tensor_info(hedit_engine* h, int index, char* name_buf, int name_len, int64_t* dims4) {
  if (!h) return fail("null engine");
  std::string name; std::vector<int64_t> shape;
  if (!h->E->tensor_info(index, name, shape)) return fail("tensor index out of range");
  if (int(name.size()) + 1 > name_len || shape.size() > 4) return fail("tensor_info buffer too small");
  memcpy(name_buf, name.c_str(), name.size() + 1);
  for (size_t i = 0; i < shape.size(); ++i) dims4[i] = shape[i];
  return int(shape.size());
}

int hedit_engine_set_graph_replay(hedit_engine* h, int on) {
  if (!h) return fail("null engine");
  h->E->set_graph_replay(on != 0);
  if (!on) h->E->drop_graphs();
  return 0;
}
int hedit_engine_set_prefix_dedup(hedit_engine* h, int on) {
  if (!h) return fail("null engine");
  h->E->set_prefix_dedup(on != 0);
  if (!on) h->E->drop_graphs();
  return 0;
}

int hedit_engine_set_splitk(hedit_engine* h, int on) {
  if (!h) return fail("null engine");
  h->E->set_splitk(on != 0);
  return 0;
}

int hedit_engine_profile_forward(hedit_engine* h, int S, int reps, char* out, int out_len) {
  if (!h) return fail("null engine");
  DeviceGuard guard_(h->device);
  Engine& E = *h->E;
  if (S > h->cap) return fail("S exceeds max_samples");
  const size_t lat = size_t(E.latent_elems());
  float *x = nullptr, *eps = nullptr, *ctx = nullptr;
  const size_t ctx_n = size_t(77) * E.cfg().ctx_dim;
  if (cudaMalloc(&x, S * lat * sizeof(float)) != cudaSuccess || cudaMalloc(&eps, S * lat * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&ctx, ctx_n * sizeof(float)) != cudaSuccess)
    return fail("cudaMalloc");
  cudaMemset(x, 0, S * lat * sizeof(float)); cudaMemset(ctx, 0, ctx_n * sizeof(float));
  const float t = 501.f;
  std::vector<int> zeros(S, 0), ident(S), minus1(S, -1);
  for (int s = 0; s < S; ++s) ident[s] = s;
  if (E.set_timesteps(&t, 1, 0) || E.set_contexts(ctx, 1, 0)) return fail(E.error());
  cudaMemcpy(h->d_tidx, zeros.data(), S * sizeof(int), cudaMemcpyHostToDevice);
  cudaMemcpy(h->d_ctx_idx, zeros.data(), S * sizeof(int), cudaMemcpyHostToDevice);
  cudaMemcpy(h->d_unit0, ident.data(), S * sizeof(int), cudaMemcpyHostToDevice);
  cudaMemcpy(h->d_unit1, minus1.data(), S * sizeof(int), cudaMemcpyHostToDevice);
  cudaMemcpy(h->d_uimg, zeros.data(), S * sizeof(int), cudaMemcpyHostToDevice);
  CallCtrl cc;
  cc.ctx_idx = h->d_ctx_idx; cc.time_idx = h->d_tidx; cc.unit_s0 = h->d_unit0; cc.unit_s1 = h->d_unit1; cc.unit_img = h->d_uimg; cc.n_units = S;
  std::map<std::string, std::pair<double, long>> acc;
  if (E.forward(x, eps, S, cc, 0) < 0) return fail(E.error());     // warm-up
  for (int r = 0; r < reps; ++r)
    if (E.forward_profiled(x, eps, S, cc, 0, acc) < 0) return fail(E.error());
  // the whole forward, launched kernel by kernel vs replayed from a CUDA graph ("@" records; how much of it is launch gaps)
  double ms_imm = 0, ms_graph = 0;
  {
    cudaStream_t st; cudaStreamCreate(&st);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    E.forward(x, eps, S, cc, st);
    cudaEventRecord(e0, st);
    for (int r = 0; r < reps; ++r) E.forward(x, eps, S, cc, st);
    cudaEventRecord(e1, st); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); ms_imm = ms / reps;
    cudaGraph_t g = nullptr; cudaGraphExec_t ge = nullptr;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      E.forward(x, eps, S, cc, st);
      if (cudaStreamEndCapture(st, &g) == cudaSuccess && g && cudaGraphInstantiate(&ge, g, 0) == cudaSuccess) {
        cudaGraphLaunch(ge, st);
        cudaEventRecord(e0, st);
        for (int r = 0; r < reps; ++r) cudaGraphLaunch(ge, st);
        cudaEventRecord(e1, st); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); ms_graph = ms / reps;
      }
    }
    cudaGetLastError();
    if (ge) cudaGraphExecDestroy(ge);
    if (g) cudaGraphDestroy(g);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaStreamDestroy(st);
  }
  cudaFree(x); cudaFree(eps); cudaFree(ctx);
  std::string js = "@immediate:" + std::to_string(ms_imm) + ":0;@graph:" + std::to_string(ms_graph) + ":0;";
  for (auto& kv : acc) js += kv.first + ":" + std::to_string(kv.second.first / reps) + ":" + std::to_string(kv.second.second / reps) + ";";
  if (int(js.size()) + 1 > out_len) return fail("profile buffer too small");
  memcpy(out, js.c_str(), js.size() + 1);
  return 0;
}

static int unet_forward_impl(hedit_engine* h, const float* x, const float* timesteps, const float* ctx, int n_ctx, const int32_t* ctx_idx,
                             int S, float* eps, void* stream, hedit_attn_probs_fn probs_cb, void* probs_user,
                             hedit_attn_editor_fn editor_cb = nullptr, void* editor_user = nullptr) {
  if (!h) return fail("null engine");
  DeviceGuard guard_(h->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  Engine& E = *h->E;
  if (S > h->cap) return fail("S exceeds max_samples");
  // distinct timesteps -> table rows
  std::vector<float> uniq; std::vector<int> tidx(S), ident(S), minus1(S, -1), zeros(S, 0), cidx(S);
  for (int s = 0; s < S; ++s) {
    int j = -1;
    for (size_t k = 0; k < uniq.size(); ++k) if (uniq[k] == timesteps[s]) j = int(k);
    if (j < 0) { uniq.push_back(timesteps[s]); j = int(uniq.size()) - 1; }
    tidx[s] = j; ident[s] = s;
    cidx[s] = ctx_idx ? ctx_idx[s] : s;
    if (cidx[s] < 0 || cidx[s] >= n_ctx) return fail("ctx_idx out of range");
  }
  if (E.set_timesteps(uniq.data(), int(uniq.size()), st)) return fail(E.error());
  if (E.set_contexts(ctx, n_ctx, st)) return fail(E.error());
  cudaMemcpyAsync(h->d_tidx, tidx.data(), S * sizeof(int), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(h->d_ctx_idx, cidx.data(), S * sizeof(int), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(h->d_unit0, ident.data(), S * sizeof(int), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(h->d_unit1, minus1.data(), S * sizeof(int), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(h->d_uimg, zeros.data(), S * sizeof(int), cudaMemcpyHostToDevice, st);
  cudaStreamSynchronize(st);     // host vectors go out of scope
  CallCtrl cc;
  cc.ctx_idx = h->d_ctx_idx; cc.time_idx = h->d_tidx; cc.unit_s0 = h->d_unit0; cc.unit_s1 = h->d_unit1; cc.unit_img = h->d_uimg; cc.n_units = S;
  cc.probs_cb = probs_cb; cc.probs_user = probs_user;
  cc.editor_cb = editor_cb; cc.editor_user = editor_user;
  const long r = E.forward(x, eps, S, cc, st);
  if (r < 0) return fail(E.error());
  cudaError_t e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return cuda_fail(e, "unet forward");
  return int(r);
}

int hedit_unet_forward_indexed(hedit_engine* h, const float* x, const float* timesteps, const float* ctx, int n_ctx, const int32_t* ctx_idx,
                               int S, float* eps, void* stream) {
  return unet_forward_impl(h, x, timesteps, ctx, n_ctx, ctx_idx, S, eps, stream, nullptr, nullptr);
}

int hedit_unet_forward_compat(hedit_engine* h, const float* x, const float* timesteps, const float* ctx, int S, float* eps,
                              hedit_attn_probs_fn probs_hook, void* user, void* stream) {
  if (!probs_hook) return fail("hedit_unet_forward_compat needs a probabilities hook");
  return unet_forward_impl(h, x, timesteps, ctx, S, nullptr, S, eps, stream, probs_hook, user);
}

int hedit_unet_forward_editor(hedit_engine* h, const float* x, const float* timesteps, const float* ctx, int S, float* eps,
                              hedit_attn_editor_fn editor_hook, void* user, void* stream) {
  if (!editor_hook) return fail("hedit_unet_forward_editor needs an editor hook");
  return unet_forward_impl(h, x, timesteps, ctx, S, nullptr, S, eps, stream, nullptr, nullptr, editor_hook, user);
}

int hedit_unet_forward(hedit_engine* h, const float* x, const float* timesteps, const float* ctx, int S, float* eps, void* stream) {
  return hedit_unet_forward_indexed(h, x, timesteps, ctx, S, nullptr, S, eps, stream);
}

// ------------------------------------------------------------------------------------------------ VAE decoder
hedit_vae* hedit_vae_create(const hedit_vae_config* cfg, int device) {
  if (!cfg) { fail("null config"); return nullptr; }
  if (hedit_device_count() <= device) { fail("hedit_b200 requires a CUDA device (sm_100a); none visible"); return nullptr; }
  DeviceGuard guard_(device);
  if (guard_.err != cudaSuccess) { cuda_fail(guard_.err, "cudaSetDevice"); return nullptr; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) { fail("hedit_b200 kernels are built for sm_100a only"); return nullptr; }
  VaeCfg c;
  c.latent_ch = cfg->latent_channels; c.out_ch = cfg->out_channels; c.layers = cfg->layers_per_block; c.groups = cfg->norm_groups;
  for (int i = 0; i < 4; ++i) c.boc[i] = cfg->block_out_channels[i];
  if (c.latent_ch != 4 || c.out_ch > 4 || c.groups > 32) { fail("unsupported VAE config (need 4 latent channels, <= 4 output channels, <= 32 groups)"); return nullptr; }
  for (int i = 0; i < 4; ++i)
    if (c.boc[i] % 64 != 0 || c.boc[i] > 2048 || c.boc[i] % c.groups != 0) { fail("VAE block_out_channels must be multiples of 64 (and of the group count), <= 2048"); return nullptr; }
  hedit_vae* v = new hedit_vae();
  v->device = device;
  v->D = new VaeDecoder(c);
  if (!v->D->ok()) { fail(v->D->error()); delete v->D; delete v; return nullptr; }
  return v;
}
void hedit_vae_destroy(hedit_vae* v) {
  if (!v) return;
  DeviceGuard guard_(v->device);
  delete v->D;
  delete v;
}
int hedit_vae_load_tensor(hedit_vae* v, const char* name, const float* data, const int64_t* dims, int ndim) {
  if (!v) return fail("null vae");
  DeviceGuard guard_(v->device);
  const int r = v->D->load_tensor(name, data, dims, ndim, 0);
  if (r) return fail(v->D->error(), r);
  return 0;
}
int hedit_vae_finalize(hedit_vae* v) {
  if (!v) return fail("null vae");
  std::string missing;
  if (v->D->finalize(&missing)) return fail(v->D->error());
  return 0;
}
int hedit_vae_tensor_count(hedit_vae* v) { return v ? v->D->tensor_count() : fail("null vae"); }
int hedit_vae_tensor_info(hedit_vae* v, int index, char* name_buf, int name_len, int64_t* dims4) {
  if (!v) return fail("null vae");
  std::string name; std::vector<int64_t> shape;
  if (!v->D->tensor_info(index, name, shape)) return fail("tensor index out of range");
  if (int(name.size()) + 1 > name_len || shape.size() > 4) return fail("tensor_info buffer too small");
  memcpy(name_buf, name.c_str(), name.size() + 1);
  for (size_t i = 0; i < shape.size(); ++i) dims4[i] = shape[i];
  return int(shape.size());
}
int hedit_vae_decode(hedit_vae* v, const float* z, float* img, int B, int h, int w, void* stream) {
  if (!v) return fail("null vae");
  DeviceGuard guard_(v->device);
  if (v->D->decode(z, img, B, h, w, reinterpret_cast<cudaStream_t>(stream))) return fail(v->D->error());
  return int(v->D->launches());
}
int hedit_vae_decode_backward(hedit_vae* v, const float* dimg, float* dz, void* stream) {
  if (!v) return fail("null vae");
  DeviceGuard guard_(v->device);
  if (v->D->backward(dimg, dz, reinterpret_cast<cudaStream_t>(stream))) return fail(v->D->error());
  return int(v->D->launches());
}
double hedit_vae_last_flops(hedit_vae* v) { return v ? v->D->flops() : 0.0; }

hedit_vae_enc* hedit_vae_enc_create(const hedit_vae_config* cfg, int device) {
  if (!cfg) { fail("null config"); return nullptr; }
  if (hedit_device_count() <= device) { fail("hedit_b200 requires a CUDA device (sm_100a); none visible"); return nullptr; }
  DeviceGuard guard_(device);
  if (guard_.err != cudaSuccess) { cuda_fail(guard_.err, "cudaSetDevice"); return nullptr; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) { fail("hedit_b200 kernels are built for sm_100a only"); return nullptr; }
  VaeCfg c;
  c.latent_ch = cfg->latent_channels; c.out_ch = cfg->out_channels; c.layers = cfg->layers_per_block; c.groups = cfg->norm_groups;
  for (int i = 0; i < 4; ++i) c.boc[i] = cfg->block_out_channels[i];
  if (c.latent_ch != 4 || c.out_ch > 4 || c.groups > 32) { fail("unsupported VAE config"); return nullptr; }
  for (int i = 0; i < 4; ++i)
    if (c.boc[i] % 64 != 0 || c.boc[i] > 2048 || c.boc[i] % c.groups != 0) { fail("VAE block_out_channels must be multiples of 64 (and of the group count), <= 2048"); return nullptr; }
  hedit_vae_enc* v = new hedit_vae_enc();
  v->device = device;
  v->E = new VaeEncoder(c);
  if (!v->E->ok()) { fail(v->E->error()); delete v->E; delete v; return nullptr; }
  return v;
}
void hedit_vae_enc_destroy(hedit_vae_enc* v) {
  if (!v) return;
  DeviceGuard guard_(v->device);
  delete v->E;
  delete v;
}
int hedit_vae_enc_load_tensor(hedit_vae_enc* v, const char* name, const float* data, const int64_t* dims, int ndim) {
  if (!v) return fail("null vae encoder");
  DeviceGuard guard_(v->device);
  const int r = v->E->load_tensor(name, data, dims, ndim, 0);
  if (r) return fail(v->E->error(), r);
  return 0;
}
int hedit_vae_enc_finalize(hedit_vae_enc* v) {
  if (!v) return fail("null vae encoder");
  std::string missing;
  if (v->E->finalize(&missing)) return fail(v->E->error());
  return 0;
}
int hedit_vae_encode(hedit_vae_enc* v, const float* img, float* moments, int B, int H, int W, void* stream) {
  if (!v) return fail("null vae encoder");
  DeviceGuard guard_(v->device);
  if (v->E->encode(img, moments, B, H, W, reinterpret_cast<cudaStream_t>(stream))) return fail(v->E->error());
  return int(v->E->launches());
}

// ------------------------------------------------------------------------------------------------ CLIP text tower
hedit_text* hedit_text_create(const hedit_text_config* cfg, int device) {
  if (!cfg) { fail("null config"); return nullptr; }
  if (hedit_device_count() <= device) { fail("hedit_b200 requires a CUDA device (sm_100a); none visible"); return nullptr; }
  DeviceGuard guard_(device);
  if (guard_.err != cudaSuccess) { cuda_fail(guard_.err, "cudaSetDevice"); return nullptr; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) { fail("hedit_b200 kernels are built for sm_100a only"); return nullptr; }
  TextCfg c;
  c.vocab = cfg->vocab; c.width = cfg->width; c.heads = cfg->heads; c.layers = cfg->layers; c.ffn = cfg->ffn; c.tokens = cfg->tokens;
  if (c.heads < 1 || c.width != 64 * c.heads || c.tokens < 1 || c.tokens > 256 || c.layers < 1 || c.ffn % 64 || c.vocab < 1) {
    fail("unsupported CLIP text config (need head dim 64, <= 256 tokens, ffn % 64 == 0)");
    return nullptr;
  }
  hedit_text* h = new hedit_text();
  h->device = device;
  h->T = new ClipText(c);
  if (!h->T->ok()) { fail(h->T->error()); delete h->T; delete h; return nullptr; }
  return h;
}
void hedit_text_destroy(hedit_text* t) {
  if (!t) return;
  DeviceGuard guard_(t->device);
  delete t->T;
  delete t;
}
int hedit_text_load_tensor(hedit_text* t, const char* name, const float* data, const int64_t* dims, int ndim) {
  if (!t) return fail("null text encoder");
  DeviceGuard guard_(t->device);
  const int r = t->T->load_tensor(name, data, dims, ndim, 0);
  if (r < 0) return fail(t->T->error(), r);
  return r;
}
int hedit_text_finalize(hedit_text* t) {
  if (!t) return fail("null text encoder");
  std::string missing;
  if (t->T->finalize(&missing)) return fail(t->T->error());
  return 0;
}
int hedit_text_encode(hedit_text* t, const int32_t* ids, int B, float* out, void* stream) {
  if (!t) return fail("null text encoder");
  DeviceGuard guard_(t->device);
  if (t->T->forward(ids, B, out, reinterpret_cast<cudaStream_t>(stream))) return fail(t->T->error());
  return int(t->T->launches());
}

// ------------------------------------------------------------------------------------------------ face swapping
hedit_face* hedit_face_create(const hedit_face_config* cfg, int device) {
  if (!cfg) { fail("null config"); return nullptr; }
  if (hedit_device_count() <= device) { fail("hedit_b200 requires a CUDA device (sm_100a); none visible"); return nullptr; }
  DeviceGuard guard_(device);
  if (guard_.err != cudaSuccess) { cuda_fail(guard_.err, "cudaSetDevice"); return nullptr; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) { fail("hedit_b200 kernels are built for sm_100a only"); return nullptr; }
  FaceCfg c;
  c.ch = cfg->ch; c.nlevels = cfg->n_levels; c.nres = cfg->num_res_blocks; c.attn_res = cfg->attn_resolution; c.resolution = cfg->image_size;
  c.in_ch = cfg->in_channels; c.out_ch = cfg->out_ch;
  if (c.nlevels < 1 || c.nlevels > 8 || c.ch % 64 || c.in_ch > 4 || c.out_ch > 4 || c.nres < 1) { fail("unsupported face UNet config"); return nullptr; }
  for (int i = 0; i < 8; ++i) c.mult[i] = i < c.nlevels ? cfg->ch_mult[i] : 0;
  const int low = c.resolution >> (c.nlevels - 1);
  if ((c.resolution > 128 ? c.resolution % 128 : 128 % c.resolution) || low < 1 || (low << (c.nlevels - 1)) != c.resolution || (low * low) % 32) {
    fail("face UNet image size must divide 128 or be a multiple of it, and the coarsest level must keep >= 32 pixels");
    return nullptr;
  }
  hedit_face* h = new hedit_face();
  h->device = device;
  h->U = new FaceUNet(c);
  if (!h->U->ok()) { fail(h->U->error()); delete h->U; delete h; return nullptr; }
  return h;
}
void hedit_face_destroy(hedit_face* f) {
  if (!f) return;
  DeviceGuard guard_(f->device);
  delete f->U;
  delete f;
}
int hedit_face_load_tensor(hedit_face* f, const char* name, const float* data, const int64_t* dims, int ndim) {
  if (!f) return fail("null face engine");
  DeviceGuard guard_(f->device);
  const int r = f->U->load_tensor(name, data, dims, ndim, 0);
  if (r) return fail(f->U->error(), r);
  return 0;
}
int hedit_face_finalize(hedit_face* f) {
  if (!f) return fail("null face engine");
  std::string missing;
  if (f->U->finalize(&missing)) return fail(f->U->error());
  return 0;
}
int hedit_face_tensor_count(hedit_face* f) { return f ? f->U->tensor_count() : fail("null face engine"); }
int hedit_face_tensor_info(hedit_face* f, int index, char* name_buf, int name_len, int64_t* dims4) {
  if (!f) return fail("null face engine");
  std::string name; std::vector<int64_t> shape;
  if (!f->U->tensor_info(index, name, shape)) return fail("tensor index out of range");
  if (int(name.size()) + 1 > name_len || shape.size() > 4) return fail("tensor_info buffer too small");
  memcpy(name_buf, name.c_str(), name.size() + 1);
  for (size_t i = 0; i < shape.size(); ++i) dims4[i] = shape[i];
  return int(shape.size());
}
int hedit_face_unet_forward(hedit_face* f, const float* x, const float* t, int S, float* eps, void* stream) {
  if (!f) return fail("null face engine");
  DeviceGuard guard_(f->device);
  if (f->U->forward(x, t, eps, S, reinterpret_cast<cudaStream_t>(stream))) return fail(f->U->error());
  return int(f->U->launches());
}
double hedit_face_last_flops(hedit_face* f) { return f ? f->U->flops() : 0.0; }
int hedit_face_edit(hedit_face* f, hedit_face_args* args, void* stream) {
  if (!f || !args) return fail("null face engine / args");
  DeviceGuard guard_(f->device);
  if (run_face_edit(*f->U, *args, reinterpret_cast<cudaStream_t>(stream))) return fail(f->U->error().empty() ? "face edit failed" : f->U->error());
  return 0;
}

// ------------------------------------------------------------------------------------------------ face-swapping reward networks
struct hedit_arcface { ArcFaceNet* N; int device; };
struct hedit_lpips { LpipsNet* N; int device; };

static bool reward_device_ok(int device) {
  if (hedit_device_count() <= device) { fail("hedit_b200 requires a CUDA device (sm_100a); none visible"); return false; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) { fail("hedit_b200 kernels are built for sm_100a only"); return false; }
  return true;
}
hedit_arcface* hedit_arcface_create(int device) {
  if (!reward_device_ok(device)) return nullptr;
  DeviceGuard guard_(device);
  if (guard_.err != cudaSuccess) { cuda_fail(guard_.err, "cudaSetDevice"); return nullptr; }
  hedit_arcface* a = new hedit_arcface();
  a->device = device; a->N = new ArcFaceNet();
  return a;
}
void hedit_arcface_destroy(hedit_arcface* a) {
  if (!a) return;
  DeviceGuard guard_(a->device);
  delete a->N;
  delete a;
}
int hedit_arcface_load_tensor(hedit_arcface* a, const char* name, const float* data, const int64_t* dims, int ndim) {
  if (!a) return fail("null arcface handle");
  DeviceGuard guard_(a->device);
  if (a->N->load_tensor(name, data, dims, ndim, 0)) return fail(a->N->error());
  return 0;
}
int hedit_arcface_finalize(hedit_arcface* a) {
  if (!a) return fail("null arcface handle");
  DeviceGuard guard_(a->device);
  if (a->N->finalize(0)) return fail(a->N->error());
  return 0;
}
int hedit_arcface_features(hedit_arcface* a, const float* img, int B, float* feat, void* stream) {
  if (!a) return fail("null arcface handle");
  DeviceGuard guard_(a->device);
  if (a->N->features(img, B, feat, reinterpret_cast<cudaStream_t>(stream))) return fail(a->N->error());
  return int(a->N->launches());
}
int hedit_arcface_set_reference(hedit_arcface* a, const float* img, void* stream) {
  if (!a) return fail("null arcface handle");
  DeviceGuard guard_(a->device);
  if (a->N->set_reference(img, reinterpret_cast<cudaStream_t>(stream))) return fail(a->N->error());
  return 0;
}
int hedit_arcface_loss_grad(hedit_arcface* a, const float* img, int B, float* loss, float* grad, void* stream) {
  if (!a) return fail("null arcface handle");
  DeviceGuard guard_(a->device);
  if (a->N->loss_grad(img, B, loss, grad, reinterpret_cast<cudaStream_t>(stream))) return fail(a->N->error());
  return int(a->N->launches());
}
double hedit_arcface_last_flops(hedit_arcface* a) { return a ? a->N->flops() : 0.0; }

hedit_lpips* hedit_lpips_create(int device) {
  if (!reward_device_ok(device)) return nullptr;
  DeviceGuard guard_(device);
  if (guard_.err != cudaSuccess) { cuda_fail(guard_.err, "cudaSetDevice"); return nullptr; }
  hedit_lpips* l = new hedit_lpips();
  l->device = device; l->N = new LpipsNet();
  return l;
}
void hedit_lpips_destroy(hedit_lpips* l) {
  if (!l) return;
  DeviceGuard guard_(l->device);
  delete l->N;
  delete l;
}
int hedit_lpips_load_tensor(hedit_lpips* l, const char* name, const float* data, const int64_t* dims, int ndim) {
  if (!l) return fail("null lpips handle");
  DeviceGuard guard_(l->device);
  if (l->N->load_tensor(name, data, dims, ndim, 0)) return fail(l->N->error());
  return 0;
}
int hedit_lpips_finalize(hedit_lpips* l) {
  if (!l) return fail("null lpips handle");
  DeviceGuard guard_(l->device);
  if (l->N->finalize(0)) return fail(l->N->error());
  return 0;
}
int hedit_lpips_set_source(hedit_lpips* l, const float* img, int n, int R, void* stream) {
  if (!l) return fail("null lpips handle");
  DeviceGuard guard_(l->device);
  if (l->N->set_source(img, n, R, reinterpret_cast<cudaStream_t>(stream))) return fail(l->N->error());
  return 0;
}
int hedit_lpips_loss_grad(hedit_lpips* l, const float* img, int B, float* loss, float* grad, void* stream) {
  if (!l) return fail("null lpips handle");
  DeviceGuard guard_(l->device);
  if (l->N->loss_grad(img, B, loss, grad, reinterpret_cast<cudaStream_t>(stream))) return fail(l->N->error());
  return int(l->N->launches());
}
double hedit_lpips_last_flops(hedit_lpips* l) { return l ? l->N->flops() : 0.0; }

// ------------------------------------------------------------------------------------------------ CLIP-Gram style reward
hedit_clip* hedit_clip_create(const hedit_clip_config* cfg, int device) {
  if (!cfg) { fail("null config"); return nullptr; }
  if (hedit_device_count() <= device) { fail("hedit_b200 requires a CUDA device (sm_100a); none visible"); return nullptr; }
  DeviceGuard guard_(device);
  if (guard_.err != cudaSuccess) { cuda_fail(guard_.err, "cudaSetDevice"); return nullptr; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) { fail("hedit_b200 kernels are built for sm_100a only"); return nullptr; }
  ClipCfg c;
  c.resolution = cfg->resolution; c.patch = cfg->patch; c.width = cfg->width; c.heads = cfg->heads; c.layers = cfg->layers;
  const int G = c.patch > 0 ? c.resolution / c.patch : 0;
  if (c.patch < 1 || c.resolution % c.patch || c.heads < 1 || c.width != 64 * c.heads || G * G + 1 > 256 || c.layers < 1 || c.width % 64 ||
      (3 * c.patch * c.patch) % 8) {
    fail("unsupported CLIP config (need a ViT with head dim 64, <= 256 tokens, width % 64 == 0)");
    return nullptr;
  }
  hedit_clip* h = new hedit_clip();
  h->device = device;
  h->C = new ClipGram(c);
  if (!h->C->ok()) { fail(h->C->error()); delete h->C; delete h; return nullptr; }
  return h;
}
void hedit_clip_destroy(hedit_clip* c) {
  if (!c) return;
  DeviceGuard guard_(c->device);
  delete c->C;
  delete c;
}
int hedit_clip_load_tensor(hedit_clip* c, const char* name, const float* data, const int64_t* dims, int ndim) {
  if (!c) return fail("null clip");
  DeviceGuard guard_(c->device);
  const int r = c->C->load_tensor(name, data, dims, ndim, 0);
  if (r < 0) return fail(c->C->error(), r);
  return r;
}
int hedit_clip_finalize(hedit_clip* c) {
  if (!c) return fail("null clip");
  std::string missing;
  if (c->C->finalize(&missing)) return fail(c->C->error());
  return 0;
}
int hedit_clip_set_reference(hedit_clip* c, const float* ref, void* stream) {
  if (!c) return fail("null clip");
  DeviceGuard guard_(c->device);
  if (c->C->set_reference(ref, reinterpret_cast<cudaStream_t>(stream))) return fail(c->C->error());
  return 0;
}
int hedit_clip_gram_loss(hedit_clip* c, const float* img, int B, int H, int W, float* loss, void* stream) {
  if (!c) return fail("null clip");
  DeviceGuard guard_(c->device);
  if (c->C->forward(img, B, H, W, loss, reinterpret_cast<cudaStream_t>(stream))) return fail(c->C->error());
  return int(c->C->launches());
}
int hedit_clip_gram_backward(hedit_clip* c, float* dimg, void* stream) {
  if (!c) return fail("null clip");
  DeviceGuard guard_(c->device);
  if (c->C->backward(dimg, reinterpret_cast<cudaStream_t>(stream))) return fail(c->C->error());
  return int(c->C->launches());
}

int hedit_edit_p2p(hedit_engine* h, hedit_edit_args* args, void* stream) {
  if (!h || !args) return fail("null engine/args");
  DeviceGuard guard_(h->device);
  const int r = run_edit(*h->E, *args, reinterpret_cast<cudaStream_t>(stream));
  if (r) {
    cudaError_t e = cudaGetLastError();
    return fail(h->E->error().empty() ? std::string("edit failed: ") + cudaGetErrorString(e) : h->E->error());
  }
  return 0;
}

int hedit_abi_sizeof(const char* name) {
  if (!name) return -1;
  const std::string n(name);
#define HEDIT_SZ(T) if (n == #T) return int(sizeof(T));
  HEDIT_SZ(hedit_edit_args) HEDIT_SZ(hedit_step_coef) HEDIT_SZ(hedit_unet_config) HEDIT_SZ(hedit_vae_config) HEDIT_SZ(hedit_clip_config)
  HEDIT_SZ(hedit_text_config) HEDIT_SZ(hedit_face_config) HEDIT_SZ(hedit_face_step_coef) HEDIT_SZ(hedit_face_args)
#undef HEDIT_SZ
  return -1;
}

// ------------------------------------------------------------------------------------------------ single operators
static int pick_bn_op(int M, int N) {
  static const int force = getenv("HEDIT_GEMM_BN") ? atoi(getenv("HEDIT_GEMM_BN")) : 0;     // tuning switch
  if (force == 160 || force == 256) return force;
  // Measured on B200: one 128 x BN x 16 MMA step costs ~ BN/2 + 90 cycles (operand fetch + TMA refill share the SM's shared-
  // memory bandwidth), so wide tiles win unless they add a wave or mostly-empty columns.
  auto cost = [&](int bn) {
    const long tiles = long((M + 127) / 128) * ((N + bn - 1) / bn);
    return double((tiles + 147) / 148) * (bn / 2 + 90);
  };
  return cost(256) < cost(160) ? 256 : 160;
}

// launch a prepared GEMM, through the split-K path when the engine's planner would choose it (operator-level tests cover both)
static cudaError_t launch_gemm_op(GemmParams& g, int bn, cudaStream_t st) {
  const int splits = gemm_splitk_splits(g, bn);
  if (splits <= 1) return launch_gemm(g, bn, st);
  float* ws = nullptr;
  cudaError_t e = cudaMalloc(&ws, size_t(splits) * g.M * g.N * sizeof(float));
  if (e != cudaSuccess) return e;
  SplitKReduceParams red;
  enable_splitk(g, splits, ws, red);
  e = launch_gemm(g, bn, st);
  if (e == cudaSuccess) e = launch_splitk_reduce(red, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(ws);
  return e;
}

int hedit_op_linear(const void* A, const void* W, const float* bias, const float* residual, float* out_f32, void* out_bf16, int M, int N,
                    int K, void* stream) {
  GemmParams g; memset(&g, 0, sizeof g);
  const int bn = pick_bn_op(M, N);
  g.M = M; g.N = N; g.num_kb = (K + 63) / 64; g.a_mode = A_LINEAR;
  uint64_t da[2] = {uint64_t(K), uint64_t(M)}, sa[1] = {uint64_t(K) * 2}; uint32_t ba[2] = {64, 128};
  uint64_t db[2] = {uint64_t(K), uint64_t(N)}, sb[1] = {uint64_t(K) * 2}; uint32_t bb[2] = {64, uint32_t(gemm_cluster() ? bn / 2 : bn)};
  g.b_full_box = gemm_cluster() ? 0 : 1;
  if (!make_tmap_bf16(&g.tmA, A, 2, da, sa, ba) || !make_tmap_bf16(&g.tmB, W, 2, db, sb, bb)) return fail("tensor map encode failed");
  g.ep.bias = bias; g.ep.residual = residual; g.ep.ldr = N; g.ep.out_f32 = out_f32; g.ep.ldo = N;
  g.ep.out_bf16 = reinterpret_cast<op_t*>(out_bf16); g.ep.ldob = N; g.ep.rows_per_group = 1;
  g.ep.diag_skip = getenv("HEDIT_GEMM_DIAG_SKIP") ? atoi(getenv("HEDIT_GEMM_DIAG_SKIP")) : 0;
  cudaError_t e = launch_gemm_op(g, bn, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "linear launch");
  return 0;
}

int hedit_op_linear_geglu(const void* A, const void* W, const float* bias, void* out_h16, int M, int N2, int K, void* stream) {
  if (N2 % 32) return fail("GEGLU width must be a multiple of 32");
  if (!bias) {               // the kernel's GEGLU write-back always adds a bias: hand it zeros
    static float* zeros = nullptr; static int zeros_n = 0;
    if (zeros_n < N2) {
      if (zeros) cudaFree(zeros);
      if (cudaMalloc(&zeros, size_t(N2) * sizeof(float)) != cudaSuccess) { zeros = nullptr; zeros_n = 0; return fail("cudaMalloc"); }
      cudaMemset(zeros, 0, size_t(N2) * sizeof(float)); zeros_n = N2;
    }
    bias = zeros;
  }
  GemmParams g; memset(&g, 0, sizeof g);
  const int bn = pick_bn_op(M, N2);
  g.M = M; g.N = N2; g.num_kb = (K + 63) / 64; g.a_mode = A_LINEAR;
  uint64_t da[2] = {uint64_t(K), uint64_t(M)}, sa[1] = {uint64_t(K) * 2}; uint32_t ba[2] = {64, 128};
  uint64_t db[2] = {uint64_t(K), uint64_t(N2)}, sb[1] = {uint64_t(K) * 2}; uint32_t bb[2] = {64, uint32_t(gemm_cluster() ? bn / 2 : bn)};
  g.b_full_box = gemm_cluster() ? 0 : 1;
  if (!make_tmap_bf16(&g.tmA, A, 2, da, sa, ba) || !make_tmap_bf16(&g.tmB, W, 2, db, sb, bb)) return fail("tensor map encode failed");
  g.ep.bias = bias; g.ep.out_bf16 = reinterpret_cast<op_t*>(out_h16); g.ep.ldob = N2 / 2; g.ep.geglu = 1; g.ep.rows_per_group = 1;
  g.ep.diag_skip = getenv("HEDIT_GEMM_DIAG_SKIP") ? atoi(getenv("HEDIT_GEMM_DIAG_SKIP")) : 0;
  cudaError_t e = launch_gemm(g, bn, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "GEGLU linear launch");
  return 0;
}

int hedit_op_conv3x3(const void* x, const void* w, const float* bias, float* out, int S, int Hin, int Win, int C, int Cout, int stride,
                     void* stream) {
  GemmParams g; memset(&g, 0, sizeof g);
  const int H = Hin / stride, Wd = Win / stride, M = S * H * Wd;
  const int bn = pick_bn_op(M, Cout);
  if (Wd > 128 || 128 % Wd != 0 || C % 64 != 0) return fail("conv geometry unsupported");
  const int BH = std::min(H, 128 / Wd), BS = 128 / (Wd * BH);
  if (H % BH) return fail("conv geometry unsupported (H)");
  g.M = M; g.N = Cout; g.num_kb = 9 * C / 64; g.conv_W = Wd; g.conv_H = H; g.conv_cin = C; g.cin_blocks = C / 64;
  bool ok;
  if (stride == 1) {
    g.a_mode = A_CONV3X3;
    uint64_t d[4] = {uint64_t(C), uint64_t(Wd), uint64_t(H), uint64_t(S)}, s[3] = {uint64_t(C) * 2, uint64_t(Wd) * C * 2, uint64_t(H) * Wd * C * 2};
    uint32_t b[4] = {64, uint32_t(Wd), uint32_t(BH), uint32_t(BS)};
    ok = make_tmap_bf16(&g.tmA, x, 4, d, s, b);
  } else {
    g.a_mode = A_CONV3X3S2;
    uint64_t d[5] = {uint64_t(2 * C), uint64_t(Wd), 2, uint64_t(H), uint64_t(S)};
    uint64_t s[4] = {uint64_t(2 * C) * 2, uint64_t(Win) * C * 2, uint64_t(2) * Win * C * 2, uint64_t(Hin) * Win * C * 2};
    uint32_t b[5] = {64, uint32_t(Wd), 1, uint32_t(BH), uint32_t(BS)};
    ok = make_tmap_bf16(&g.tmA, x, 5, d, s, b);
  }
  uint64_t db[2] = {uint64_t(9 * C), uint64_t(Cout)}, sb[1] = {uint64_t(9 * C) * 2}; uint32_t bb[2] = {64, uint32_t(gemm_cluster() ? bn / 2 : bn)};
  g.b_full_box = gemm_cluster() ? 0 : 1;
  if (!ok || !make_tmap_bf16(&g.tmB, w, 2, db, sb, bb)) return fail("tensor map encode failed");
  g.ep.bias = bias; g.ep.out_f32 = out; g.ep.ldo = Cout; g.ep.rows_per_group = 1;
  cudaError_t e = launch_gemm_op(g, bn, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "conv launch");
  return 0;
}

static bool attn_maps(AttnParams& a, const void* q, int ldq, int Nq, int Sq, const void* k, const void* v, int ldkv, int Nkv, int Skv, int H,
                      int d, int bkv) {
  auto mk = [&](CUtensorMap* m, const void* base, int ld, int N, int S, int rows) {
    uint64_t dims[4] = {uint64_t(d), uint64_t(H), uint64_t(N), uint64_t(S)};
    uint64_t str[3] = {uint64_t(d) * 2, uint64_t(ld) * 2, uint64_t(N) * ld * 2};
    uint32_t box[4] = {64, 1, uint32_t(rows), 1};
    return make_tmap_bf16(m, base, 4, dims, str, box);
  };
  return mk(&a.tmQ, q, ldq, Nq, Sq, 128) && mk(&a.tmK, k, ldkv, Nkv, Skv, bkv) && mk(&a.tmV, v, ldkv, Nkv, Skv, bkv);
}

int hedit_op_self_attention(const void* q, const void* k, const void* v, int ldq, int ldkv, int S, int Nq, int Nkv, int H, int d,
                            const int32_t* q_idx, const int32_t* k_idx, const int32_t* v_idx, void* out, void* stream) {
  if (d % 8 || d > 192) return fail("head dim must be a multiple of 8 and <= 192");
  AttnParams a; memset(&a, 0, sizeof a);
  const int dch = d <= 64 ? 1 : (d <= 128 ? 2 : 3), bkv = self_attn_bkv(d, Nq, Nkv);
  if (!attn_maps(a, q, ldq, Nq, S, k, v, ldkv, Nkv, S, H, d, bkv)) return fail("tensor map encode failed");
  a.H = H; a.d = d; a.Nq = Nq; a.Nkv = Nkv; a.scale_log2 = float(1.4426950408889634 / sqrt(double(d)));
  a.q_idx = q_idx; a.k_idx = k_idx; a.v_idx = v_idx; a.out = reinterpret_cast<op_t*>(out); a.ldo = H * d;
  cudaError_t e = launch_self_attn(a, dch, S, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "self attention launch");
  return 0;
}

int hedit_op_cross_attention_p2p(const void* q, const void* kv, int S, int n_ctx, int Nq, int H, int d, int n_units, const int32_t* unit_s0,
                                 const int32_t* unit_s1, const int32_t* unit_img, const int32_t* ctx_idx, const int32_t* mapper,
                                 const float* c_base, const float* c_tar, const float* replace_m, const int32_t* is_replace,
                                 float* blend_acc, const float* blend_alpha, int blend_layer, int n_blend_layers, void* out, void* stream) {
  if (d % 8 || d > 192) return fail("head dim must be a multiple of 8 and <= 192");
  AttnParams a; memset(&a, 0, sizeof a);
  const int C = H * d;
  const int dch = d <= 64 ? 1 : (d <= 128 ? 2 : 3);
  const op_t* kvp = reinterpret_cast<const op_t*>(kv);
  if (!attn_maps(a, q, C, Nq, S, kvp, kvp + C, 2 * C, 77, n_ctx, H, d, 80)) return fail("tensor map encode failed");
  a.H = H; a.d = d; a.Nq = Nq; a.Nkv = 77; a.scale_log2 = float(1.4426950408889634 / sqrt(double(d)));
  a.out = reinterpret_cast<op_t*>(out); a.ldo = C;
  a.unit_s0 = unit_s0; a.unit_s1 = unit_s1; a.unit_img = unit_img; a.ctx_idx = ctx_idx; a.mapper = mapper; a.c_base = c_base; a.c_tar = c_tar;
  a.replace_m = replace_m; a.is_replace = is_replace; a.blend_acc = blend_acc; a.blend_alpha = blend_alpha; a.blend_layer = blend_layer;
  a.n_blend_layers = n_blend_layers; a.blend_rows = 2; a.map_w = nullptr; a.map_rows = 1;
  cudaError_t e = launch_cross_attn(a, dch, n_units, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "cross attention launch");
  return 0;
}

int hedit_op_group_norm(const float* x, const float* gamma, const float* beta, void* out, int S, int HW, int C, int groups, float eps, int silu,
                        void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (groups > 32 || C % groups || (C / groups) % 2 || C % 4) return fail("group norm geometry unsupported");
  const int chunk = std::max(16, HW / 64), nch = (HW + chunk - 1) / chunk;
  float2* partial = nullptr;
  if (cudaMalloc(&partial, size_t(S) * nch * groups * sizeof(float2)) != cudaSuccess) return fail("cudaMalloc");
  GNStatsParams sp{x, nullptr, C, 0, HW, groups, chunk, partial};
  gn_stats_kernel<<<dim3(nch, S), std::min(640, ((C / 4 + 31) / 32) * 32), 0, st>>>(sp);
  GNApplyParams ap{x, nullptr, C, 0, HW, groups, 16, nch, partial, gamma, beta, eps, silu, reinterpret_cast<op_t*>(out), nullptr};
  gn_apply_kernel<<<dim3((HW + 15) / 16, S), std::max(256, (C / 4) * std::max(1, (256 + C / 4 - 1) / (C / 4))), 0, st>>>(ap);
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(partial);
  if (e != cudaSuccess) return cuda_fail(e, "group norm");
  return 0;
}

int hedit_op_layer_norm(const float* x, const float* gamma, const float* beta, void* out, int rows, int C, float eps, void* stream) {
  if (C % 64 || C > 2048) return fail("layer norm needs C % 64 == 0 and C <= 2048");
  launch_layernorm(x, gamma, beta, reinterpret_cast<op_t*>(out), rows, C, eps, reinterpret_cast<cudaStream_t>(stream));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "layer norm launch");
  return 0;
}

}  // extern "C"
