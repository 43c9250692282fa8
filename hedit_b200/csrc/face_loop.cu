// h_Edit_R for face swapping (reference: face-swapping/inversion/h_edit_R.py:7-137) for a batch of B independent images: per timestep
// one denoiser call at t and the DDPM-style step to x_{t-1}^base (:72-88), then per implicit-loop iteration two reward-guided moves on
// the Tweedie prediction -- identity loss (optionally masked) and LPIPS -- each preceded by a fresh denoiser call at t-1 (:96-131).
// The reward gradients come through the hook (the caller's ArcFace / LPIPS modules); everything else runs here without host syncs.
#include "nvtx.h"
#include "../../include/hedit_b200.h"
#include "face.h"

namespace hedit {

// x_{t-1} = sqrt(abar_tm1) * (x - sqrt(1-abar_t) eps) / sqrt(abar_t) + c2 * eps + noise * z     (h_edit_R.py:78,88)
__global__ void face_reverse_kernel(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ z, size_t z_stride,
                                    float* __restrict__ out, hedit_face_step_coef k, int n) {
  const int b = blockIdx.y;
  const float* zb = z + size_t(b) * z_stride;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const size_t o = size_t(b) * n + i;
    const float e = eps[o];
    const float x0 = (x[o] - k.sqrt_1m_at * e) / k.sqrt_at;
    out[o] = (k.sqrt_atm1 * x0 + k.c2 * e) + k.noise * zb[i];
  }
}
// Tweedie prediction from x_{t-1} (h_edit_R.py:105,126)
__global__ void face_x0_kernel(const float* __restrict__ x, const float* __restrict__ eps, float* __restrict__ x0, float s1m, float sa, size_t n) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) x0[i] = (x[i] - s1m * eps[i]) / sa;
}
// x <- x - rho * dLoss/dx [* mask],  dLoss/dx = grad_x0 / sqrt(abar_tm1),  rho = sqrt(abar_tm1) * weight   (h_edit_R.py:106-116,129-131)
__global__ void face_update_kernel(float* __restrict__ x, const float* __restrict__ grad_x0, const float* __restrict__ mask, float rho, float inv_sa, size_t n) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    float g = rho * (grad_x0[i] * inv_sa);
    if (mask) g *= mask[i];
    x[i] -= g;
  }
}

int run_face_edit(FaceUNet& U, hedit_face_args& a, cudaStream_t st) {
  NvtxRange nvtx_edit_("hedit.face.edit B=%d steps=%d", a.B, a.steps);
  const FaceCfg& c = U.cfg();
  const int B = a.B, T = a.steps, n = c.in_ch * c.resolution * c.resolution;
  if (B < 1 || T < 1 || !a.xT || !a.zs || !a.coef || !a.edited) { U.err_ = "bad face edit arguments"; return -1; }
  if ((a.use_id || a.use_lpips) && (!a.reward || !a.reward_x0 || !a.reward_grad)) { U.err_ = "rewards need the hook and its two buffers"; return -1; }
  float *x = nullptr, *xn = nullptr, *eps = nullptr;
  auto cleanup = [&]() { cudaFree(x); cudaFree(xn); cudaFree(eps); };
  if (cudaMalloc(&x, size_t(B) * n * 4) != cudaSuccess || cudaMalloc(&xn, size_t(B) * n * 4) != cudaSuccess || cudaMalloc(&eps, size_t(B) * n * 4) != cudaSuccess) {
    cleanup(); U.err_ = "cudaMalloc failed"; return -1;
  }
  cudaMemcpyAsync(x, a.xT, size_t(B) * n * 4, cudaMemcpyDeviceToDevice, st);
  std::vector<float> tv(B);
  long launches = 0; long long fwd = 0;
  const size_t N = size_t(B) * n;
  const int eb = int(std::min<size_t>((N + 255) / 256, 8192));
  auto unet = [&](const float* in, float t) -> int {
    for (int b = 0; b < B; ++b) tv[b] = t;
    if (U.forward(in, tv.data(), eps, B, st)) return -1;
    launches += U.launches(); fwd += B;
    return 0;
  };
  for (int i = 0; i < T; ++i) {
    const hedit_face_step_coef& k = a.coef[i];
    const int idx = T - 1 - i;
    if (unet(x, k.t)) { cleanup(); return -1; }
    face_reverse_kernel<<<dim3(std::max(1, n / 256 / 4), B), 256, 0, st>>>(x, eps, a.zs + size_t(idx) * n, size_t(T) * n, xn, k, n);
    ++launches;
    std::swap(x, xn);
    const int K = (k.tm1 == 0.f) ? 0 : a.opt_steps;                      // h_edit_R.py:90-91
    const float rho = k.sqrt_atm1 * a.weight, inv_sa = 1.f / k.sqrt_atm1;
    for (int it = 0; it < K; ++it) {
      for (int which = 0; which < 2; ++which) {
        // the denoiser is re-evaluated before each reward (h_edit_R.py:100-101,120-123), whether or not that reward is switched on
        if (unet(x, k.tm1)) { cleanup(); return -1; }
        if (!(which == 0 ? a.use_id : a.use_lpips)) continue;
        face_x0_kernel<<<eb, 256, 0, st>>>(x, eps, a.reward_x0, k.sqrt_1m_atm1, k.sqrt_atm1, N);
        if (a.reward(a.reward_user, which, i, it) != 0) { cleanup(); U.err_ = "reward hook failed"; return -1; }
        face_update_kernel<<<eb, 256, 0, st>>>(x, a.reward_grad, which == 0 ? a.mask : nullptr, rho, inv_sa, N);
        launches += 2;
      }
    }
  }
  cudaMemcpyAsync(a.edited, x, size_t(B) * n * 4, cudaMemcpyDeviceToDevice, st);
  cudaError_t e = cudaStreamSynchronize(st);
  cleanup();
  if (e != cudaSuccess) { U.err_ = std::string("face edit: ") + cudaGetErrorString(e); return -1; }
  a.n_sample_forwards = fwd; a.n_kernel_launches = launches;
  return 0;
}

}  // namespace hedit
