// Fused element kernels of the h-Edit step (NCHW fp32 latents, the reference's layout), driven by host-precomputed
// per-step scalar tables so the loop has no host synchronisation
// (reference: text-guided/inversion/p2p_h_edit.py:616-622,658-692; inversion_utils.py:58-126,168-195).
#pragma once
#include "ptx.cuh"

namespace hedit {

// per-step scalars (computed on the host in the reference's fp32 order)
struct StepCoef {
  float sqrt_1m_at;   // sqrt(1 - abar_t)
  float sqrt_at;      // sqrt(abar_t)
  float sqrt_ap;      // sqrt(abar_prev)
  float dir;          // sqrt(1 - abar_prev - eta^2 var)   (or sqrt(1 - abar_prev) for DDIM inversion)
  float noise;        // eta * sqrt(var)                    (or eta)
  float coeff;        // h-term coefficient (p2p_h_edit.py:664-665)
};

// Call-A combine + reverse step for both rows of every image (p2p_h_edit.py:616-622):
//   eps = u + w_src (c - u);  x0 = (x - s1 eps)/s2;  x_prev = s3 x0 + dir eps + noise z
// eps_* index arrays select the UNet output sample that holds each term (supports the exact-reuse schedule).
struct ReverseParams {
  const float* xt;        // [B][2][n]  (orig, edit)
  const float* z;         // [B][n]     (pointer already offset to this step) with image stride z_stride
  size_t z_stride;
  const float* eps_u;     // base pointers of the eps pools
  const float* eps_c;
  const int* iu;          // [B][2] sample index of uncond eps for (orig, edit) inside eps_u
  const int* ic;          // [B][2]
  float w_src;
  StepCoef k;
  float* x_prev;          // [B][2][n]  (orig_{t-1}, base_{t-1})
  int n;
  int per_row;            // 1: row 1 (edit) uses its own guidance weight and scalars (the baseline samplers, p2p_baselines.py:172-184)
  float w_row1;
  StepCoef k1;
};

static __global__ void hstep_reverse_kernel(const ReverseParams p) {
  const int b = blockIdx.y, row = blockIdx.z;
  const float* eu = p.eps_u + size_t(p.iu[b * 2 + row]) * p.n;
  const float* ec = p.eps_c + size_t(p.ic[b * 2 + row]) * p.n;
  const float* x = p.xt + (size_t(b) * 2 + row) * p.n;
  const float* z = p.z + size_t(b) * p.z_stride;
  float* o = p.x_prev + (size_t(b) * 2 + row) * p.n;
  const bool own = p.per_row && row == 1;
  const float w = own ? p.w_row1 : p.w_src;
  const StepCoef k = own ? p.k1 : p.k;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4; i < p.n; i += gridDim.x * blockDim.x * 4) {
    const float4 u = *reinterpret_cast<const float4*>(eu + i), c = *reinterpret_cast<const float4*>(ec + i);
    const float4 xv = *reinterpret_cast<const float4*>(x + i), zv = *reinterpret_cast<const float4*>(z + i);
    float4 r;
#define HEDIT_REV(f)                                                          \
    {                                                                         \
      const float eps = u.f + w * (c.f - u.f);                                \
      const float x0 = (xv.f - k.sqrt_1m_at * eps) / k.sqrt_at;               \
      r.f = (k.sqrt_ap * x0 + k.dir * eps) + k.noise * zv.f;                  \
    }
    HEDIT_REV(x) HEDIT_REV(y) HEDIT_REV(z) HEDIT_REV(w)
#undef HEDIT_REV
    *reinterpret_cast<float4*>(o + i) = r;
  }
}

// Call-C combine (p2p_h_edit.py:658-667): corr = eps_tar - eps_src_edit
//   eps_src_edit = u + w_hat (c_src - u),  eps_tar = u + w_tar (c_tar - u)
// Optionally (MOS iterations k>0) emits per-image partial sums of corr^2 and of |sign(x_opt - x_base)| for rho.
struct CorrParams {
  const float* eps;       // pool
  const int* iu; const int* ics; const int* ict;   // [B] sample indices of u_tar, c_src, c_tar
  float w_src_edit, w_tar;
  float* corr;            // [B][n]
  const float* x_opt;     // [B] rows with stride x_stride, or null
  const float* x_base;    // [B] rows with stride xb_stride
  size_t x_stride, xb_stride;
  float2* partial;        // [B][gridDim.x] (sum corr^2, count nonzero sign) or null
  int n;
};

static __global__ void hstep_corr_kernel(const CorrParams p) {
  const int b = blockIdx.y;
  const float* u = p.eps + size_t(p.iu[b]) * p.n;
  const float* cs = p.eps + size_t(p.ics[b]) * p.n;
  const float* ct = p.eps + size_t(p.ict[b]) * p.n;
  float sq = 0.f, nz = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += gridDim.x * blockDim.x) {
    const float uu = u[i];
    const float e_src = uu + p.w_src_edit * (cs[i] - uu);
    const float e_tar = uu + p.w_tar * (ct[i] - uu);
    const float c = e_tar - e_src;
    p.corr[size_t(b) * p.n + i] = c;
    if (p.partial) {
      sq += c * c;
      nz += (p.x_opt[size_t(b) * p.x_stride + i] != p.x_base[size_t(b) * p.xb_stride + i]) ? 1.f : 0.f;
    }
  }
  if (p.partial) {
    __shared__ float s1[32], s2[32];
#pragma unroll
    for (int o = 16; o; o >>= 1) { sq += __shfl_xor_sync(0xffffffffu, sq, o); nz += __shfl_xor_sync(0xffffffffu, nz, o); }
    if ((threadIdx.x & 31) == 0) { s1[threadIdx.x >> 5] = sq; s2[threadIdx.x >> 5] = nz; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = 0.f, c = 0.f;
      for (int w = 0; w < (blockDim.x >> 5); ++w) { a += s1[w]; c += s2[w]; }
      p.partial[size_t(b) * gridDim.x + blockIdx.x] = make_float2(a, c);
    }
  }
}

// x_opt <- rec + coeff * corr,   rec = x_opt - rho * sign(x_opt - x_base)/n  for MOS iterations k > 0
// (p2p_h_edit.py:670-692; rho = rms(corr) / (rms(grad) + 1e-8) * w_rec, both RMS per image).
struct UpdateParams {
  float* x_opt; const float* x_base; size_t x_stride, xb_stride;
  const float* corr;
  const float2* partial; int nparts;     // null for k == 0
  float coeff, w_rec;
  int n;
};

static __global__ void hstep_update_kernel(const UpdateParams p) {
  const int b = blockIdx.y;
  float rho_g = 0.f;      // rho / n
  if (p.partial) {
    float a = 0.f, c = 0.f;
    for (int k = 0; k < p.nparts; ++k) { const float2 t = p.partial[size_t(b) * p.nparts + k]; a += t.x; c += t.y; }
    const float inv_n = 1.f / float(p.n);
    const float corr_norm = sqrtf(a * inv_n);
    const float grad_norm = sqrtf(c * inv_n * inv_n * inv_n);      // mean((sign/n)^2) = count / n^3
    rho_g = corr_norm / (grad_norm + 1e-8f) * p.w_rec * inv_n;
  }
  float* x = p.x_opt + size_t(b) * p.x_stride;
  const float* xb = p.x_base + size_t(b) * p.xb_stride;
  const float* c = p.corr + size_t(b) * p.n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += gridDim.x * blockDim.x) {
    float v = x[i];
    if (p.partial) {
      const float d = v - xb[i];
      const float sg = (d > 0.f) ? 1.f : (d < 0.f ? -1.f : 0.f);
      v = v - rho_g * sg;
    }
    x[i] = v + p.coeff * c[i];
  }
}

// ---- reward-guided (Langevin) move of the style path (text-guided-n-style/inversion/h_edit.py:150-172) ----------------------
// x0 = (x_opt - sqrt(1-abar_tt) eps_tar) / sqrt(abar_tt)   (reverse_step_pred_x0, inversion_utils.py:128-140), eps_tar = u + w_tar (c_tar - u)
struct X0PredParams {
  const float* eps; const int* iu; const int* ict;
  float w_tar, sqrt_1m_att, sqrt_att;
  const float* x_opt; size_t x_stride;
  float* x0;              // [B][n]
  int n;
};

static __global__ void hstep_x0pred_kernel(const X0PredParams p) {
  const int b = blockIdx.y;
  const float* u = p.eps + size_t(p.iu[b]) * p.n;
  const float* ct = p.eps + size_t(p.ict[b]) * p.n;
  const float* x = p.x_opt + size_t(b) * p.x_stride;
  float* o = p.x0 + size_t(b) * p.n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += gridDim.x * blockDim.x) {
    const float uu = u[i];
    const float e_tar = uu + p.w_tar * (ct[i] - uu);
    o[i] = (x[i] - p.sqrt_1m_att * e_tar) / p.sqrt_att;
  }
}

// per-image partial sums of corr^2 and of (dL/dx)^2, dL/dx = grad_x0 / sqrt(abar_tt): the two RMS values of rho (h_edit.py:166-167)
struct GuidNormParams {
  const float* corr; const float* grad_x0;   // [B][n]
  float inv_sqrt_att;
  float2* partial;        // [B][gridDim.x]
  int n;
};

static __global__ void hstep_guid_norm_kernel(const GuidNormParams p) {
  const int b = blockIdx.y;
  float sc = 0.f, sg = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += gridDim.x * blockDim.x) {
    const float c = p.corr[size_t(b) * p.n + i], g = p.grad_x0[size_t(b) * p.n + i] * p.inv_sqrt_att;
    sc = fmaf(c, c, sc); sg = fmaf(g, g, sg);
  }
  __shared__ float s1[32], s2[32];
#pragma unroll
  for (int o = 16; o; o >>= 1) { sc += __shfl_xor_sync(0xffffffffu, sc, o); sg += __shfl_xor_sync(0xffffffffu, sg, o); }
  if ((threadIdx.x & 31) == 0) { s1[threadIdx.x >> 5] = sc; s2[threadIdx.x >> 5] = sg; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) { a += s1[w]; c += s2[w]; }
    p.partial[size_t(b) * gridDim.x + blockIdx.x] = make_float2(a, c);
  }
}

// x_opt <- x_opt - rho * dL/dx,   rho = rms(corr) / rms(dL/dx) * weight  (per image; h_edit.py:166-169)
struct GuidUpdateParams {
  float* x_opt; size_t x_stride;
  const float* grad_x0; float inv_sqrt_att, weight;
  const float2* partial; int nparts;
  int n;
};

static __global__ void hstep_guid_update_kernel(const GuidUpdateParams p) {
  const int b = blockIdx.y;
  float a = 0.f, c = 0.f;
  for (int k = 0; k < p.nparts; ++k) { const float2 t = p.partial[size_t(b) * p.nparts + k]; a += t.x; c += t.y; }
  const float inv_n = 1.f / float(p.n);
  const float rho = sqrtf(a * inv_n) / sqrtf(c * inv_n) * p.weight;
  float* x = p.x_opt + size_t(b) * p.x_stride;
  const float* g = p.grad_x0 + size_t(b) * p.n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += gridDim.x * blockDim.x) x[i] -= rho * (g[i] * p.inv_sqrt_att);
}

// LocalBlend (ptp_classes.py:44-72): one CTA per image.  acc holds, per (src|tar, layer, head, pixel), the running sum
// over steps of sum_j blend_alpha_j * prob_j for the 16x16 cross-attention layers.
struct BlendParams {
  const float* acc;       // [B][rows][L][H][256]
  const int* has_blend;   // [B]
  int L, H;
  int rows;               // 2, or 4 with substruct_words (rows 2,3 = word maps of the words to EXCLUDE)
  float th, th_sub;       // LocalBlend.th[0] (pooled word mask), th[1] (un-pooled substruct mask)
  float* xt;              // [B][2][C][hh][ww]: row 1 (edit) is blended towards row 0 (orig)
  int C, hh, ww;
};

// LocalBlend.__call__ / get_mask (ptp_classes.py:44-72): word maps averaged over the 16x16 cross-attention layers and heads, 3x3 max-pool,
// normalised by the map maximum, thresholded, source | target; with substruct_words the same without pooling at th[1], negated and ANDed.
static __global__ void local_blend_kernel(const BlendParams p) {
  const int b = blockIdx.x;
  if (!p.has_blend[b]) return;
  __shared__ float m[2][256], pooled[2][256], red[2][256];
  __shared__ unsigned char mask16[256];
  const int t = threadIdx.x;     // 256 threads
  const int rows = p.rows > 2 ? 4 : 2;
  for (int pass = 0; pass < rows / 2; ++pass) {          // pass 0: blend words (pooled); pass 1: substruct words (not pooled)
    for (int which = 0; which < 2; ++which) {
      const float* a = p.acc + ((size_t(b) * rows + 2 * pass + which) * p.L * p.H) * 256 + t;
      float s = 0.f;
      for (int k = 0; k < p.L * p.H; ++k) s += a[size_t(k) * 256];
      m[which][t] = s / float(p.L * p.H);
    }
    __syncthreads();
    const int y = t >> 4, x = t & 15;
    for (int which = 0; which < 2; ++which) {
      float mx = m[which][t];
      if (pass == 0) {
        mx = -INFINITY;
        for (int dy = -1; dy <= 1; ++dy)
          for (int dx = -1; dx <= 1; ++dx) {
            const int yy = y + dy, xx = x + dx;
            if (yy >= 0 && yy < 16 && xx >= 0 && xx < 16) mx = fmaxf(mx, m[which][yy * 16 + xx]);
          }
      }
      pooled[which][t] = mx;
      red[which][t] = mx;
    }
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
      if (t < o) { red[0][t] = fmaxf(red[0][t], red[0][t + o]); red[1][t] = fmaxf(red[1][t], red[1][t + o]); }
      __syncthreads();
    }
    const float th = pass == 0 ? p.th : p.th_sub;
    const bool on = ((pooled[0][t] / red[0][0]) > th) || ((pooled[1][t] / red[1][0]) > th);
    if (pass == 0) mask16[t] = on; else mask16[t] = mask16[t] && !on;
    __syncthreads();
  }
  const int n = p.C * p.hh * p.ww;
  float* x0 = p.xt + size_t(b) * 2 * n;
  float* x1 = x0 + n;
  const int sy = p.hh / 16, sx = p.ww / 16;
  for (int i = t; i < n; i += blockDim.x) {
    const int xx = i % p.ww, yy = (i / p.ww) % p.hh;
    const float mk = mask16[(yy / sy) * 16 + (xx / sx)] ? 1.f : 0.f;
    const float o = x0[i];
    x1[i] = o + mk * (x1[i] - o);
  }
}

}  // namespace hedit
