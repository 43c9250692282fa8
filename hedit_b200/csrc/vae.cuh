// Element / reduction kernels of the VAE decoder forward and its input-gradient backward (the decode the style path differentiates
// through: text-guided-n-style/inversion/h_edit.py:158-164).  Convolutions and linear layers run on the tcgen05 GEMM (gemm.cuh);
// these kernels are the HBM-bound glue: GroupNorm backward, SiLU', nearest-upsample backward, the materialised single-head
// attention of the mid block (softmax forward / backward over 4096-wide rows, 16-bit transposes) and the 4x4 post_quant_conv.
#pragma once
#include "ptx.cuh"

namespace hedit {

HEDIT_DEVICE float vae_ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// ------------------------------------------------------------------------------------------------ weight layout conversion
// mode 0: conv fwd   [O][I][3][3] -> 16-bit [O][tap][I]
// mode 1: conv dgrad [O][I][3][3] -> 16-bit [I][8-tap][O]          (input-gradient = conv of the output gradient with flipped taps)
// mode 2: rows       [O][K]       -> 16-bit dst[o*ld + off + k]
// mode 3: rows^T     [O][I]       -> 16-bit dst[i*ld + off + o]
static __global__ void vae_cvt_weight_kernel(const float* __restrict__ src, op_t* __restrict__ dst, int O, int I, int mode, int ld, int off) {
  const size_t total = size_t(O) * I * (mode <= 1 ? 9 : 1);
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    if (mode <= 1) {
      const int tap = int(i % 9), ci = int((i / 9) % I), o = int(i / (size_t(I) * 9));
      const size_t d = mode == 0 ? (size_t(o) * 9 + tap) * I + ci : (size_t(ci) * 9 + (8 - tap)) * O + o;
      dst[d] = to_op(src[i]);
    } else {
      const int k = int(i % I), o = int(i / I);
      const size_t d = mode == 2 ? size_t(o) * ld + off + k : size_t(k) * ld + off + o;
      dst[d] = to_op(src[i]);
    }
  }
}
// conv_out.weight [3][C0][3][3] -> fp32 [C0][4][9] with flipped taps and a zero 4th input channel: the weights conv_in_kernel needs to
// compute conv_out's input gradient (3 -> C0 channels, K = 27: CUDA-core math)
static __global__ void vae_cvt_convout_dgrad_kernel(const float* __restrict__ src, float* __restrict__ dst, int O, int C0) {
  const int total = C0 * 4 * 9;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int tap = i % 9, o = (i / 9) % 4, c = i / 36;
    dst[i] = (o < O) ? src[(size_t(o) * C0 + c) * 9 + (8 - tap)] : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------ post_quant_conv (1x1, 4 -> 4), NCHW
// transpose = 0: y = W x + b ; transpose = 1: y = W^T x (input gradient)
static __global__ void vae_pointwise4_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                             float* __restrict__ y, int C, int HW, int transpose) {
  const int s = blockIdx.y;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    float in[8], out[8];
    for (int c = 0; c < C; ++c) in[c] = x[(size_t(s) * C + c) * HW + p];
    for (int o = 0; o < C; ++o) {
      float a = (b && !transpose) ? b[o] : 0.f;
      for (int c = 0; c < C; ++c) a = fmaf(transpose ? w[c * C + o] : w[o * C + c], in[c], a);
      out[o] = a;
    }
    for (int o = 0; o < C; ++o) y[(size_t(s) * C + o) * HW + p] = out[o];
  }
}

// NCHW fp32 [S][Cin][HW] -> NCHW fp32 [S][4][HW] zero-padded (the layout conv_in_kernel reads)
static __global__ void vae_pad4_kernel(const float* __restrict__ x, float* __restrict__ y, int Cin, int HW) {
  const int s = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 4 * HW; i += gridDim.x * blockDim.x) {
    const int c = i / HW, p = i - c * HW;
    y[size_t(s) * 4 * HW + i] = (c < Cin) ? x[(size_t(s) * Cin + c) * HW + p] : 0.f;
  }
}

// fp32 -> 16-bit operand copy with an optional per-sample scale 1/rms (keeps back-propagated gradients inside the fp16 range; the
// Langevin step only uses the gradient's direction and per-image RMS ratio, so the scale never has to be undone)
static __global__ void vae_cast_kernel(const float* __restrict__ x, op_t* __restrict__ y, size_t n4) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n4; i += size_t(gridDim.x) * blockDim.x) {
    const float4 t = reinterpret_cast<const float4*>(x)[i];
    reinterpret_cast<uint2*>(y)[i] = make_uint2(pack_op2(t.x, t.y), pack_op2(t.z, t.w));
  }
}

// (sum, sumsq) chunk partials of gn_stats_kernel -> (mean, rstd) per (sample, group); grid (groups, S), 128 threads
static __global__ void gn_partial_finalize_kernel(const float2* __restrict__ partial, float2* __restrict__ stats, int nchunks, int groups,
                                                  double inv_n, float eps) {
  __shared__ double su[128], sq[128];
  const int g = blockIdx.x, s = blockIdx.y;
  float a = 0.f, b = 0.f;
  for (int k = threadIdx.x; k < nchunks; k += blockDim.x) { const float2 t = partial[(size_t(s) * nchunks + k) * groups + g]; a += t.x; b += t.y; }
  su[threadIdx.x] = a; sq[threadIdx.x] = b;
  __syncthreads();
  for (int o = 64; o; o >>= 1) {
    if (threadIdx.x < o) { su[threadIdx.x] += su[threadIdx.x + o]; sq[threadIdx.x] += sq[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double mean = su[0] * inv_n;
    double var = sq[0] * inv_n - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[size_t(s) * groups + g] = make_float2(float(mean), float(1.0 / sqrt(var + double(eps))));
  }
}

// ------------------------------------------------------------------------------------------------ GroupNorm(+SiLU) backward
// y = xh * gamma + beta, xh = (x - mean) * rstd, out = silu(y) or y.  Given g = dL/dout:
//   dxh = g * silu'(y) * gamma ; per (sample, group): m1 = mean(dxh), m2 = mean(dxh * xh) ; dx = rstd * (dxh - m1 - xh * m2)
struct GNBwdParams {
  const float* g;         // [S][HW][C] fp32
  const float* x;         // [S][HW][C] fp32 (saved GroupNorm input)
  int C, HW, groups, chunk, nchunks;
  const float2* stats;    // [S][groups] (mean, rstd) saved by the forward
  const float* gamma; const float* beta;
  int silu;
  float2* partial;        // [S][nchunks][groups] (sum dxh, sum dxh*xh)
  const float2* red;      // [S][groups] (m1, m2) (apply pass)
  const float* add;       // [S][HW][C] fp32 gradient of a parallel branch added to dx, or null
  float* dx;              // fp32 out or null
  op_t* dx16;             // 16-bit out or null
};

HEDIT_DEVICE float silu_grad_f(float y) {
  const float s = __fdividef(1.0f, 1.0f + __expf(-y));
  return s * fmaf(y, 1.0f - s, 1.0f);
}

// grid (nchunks, S); blockDim = quads * nsub: thread <-> (channel quad, pixel sub-lane); fixed-order reductions (no atomics)
static __global__ void gn_bwd_stats_kernel(const GNBwdParams p) {
  __shared__ float2 csum[2048];            // [nsub][C]  (blockDim.x * 4 entries)
  __shared__ float sa[32], sb[32];
  const int quads = p.C >> 2, cpg = p.C / p.groups;
  const int s = blockIdx.y, ch = blockIdx.x;
  const int p0 = ch * p.chunk, p1 = min(p.HW, p0 + p.chunk);
  const int nsub = max(1, int(blockDim.x) / quads);
  if (threadIdx.x < p.groups) { const float2 t = p.stats[size_t(s) * p.groups + threadIdx.x]; sa[threadIdx.x] = t.x; sb[threadIdx.x] = t.y; }
  __syncthreads();
  const int v = threadIdx.x % quads, sub = threadIdx.x / quads;
  if (sub < nsub) {
    const int c = 4 * v;
    float mean[4], rstd[4], ga[4], be[4], a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 4; ++k) { const int g = (c + k) / cpg; mean[k] = sa[g]; rstd[k] = sb[g]; ga[k] = p.gamma[c + k]; be[k] = p.beta[c + k]; }
    const float* gp = p.g + size_t(s) * p.HW * p.C + c;
    const float* xp = p.x + size_t(s) * p.HW * p.C + c;
    for (int px = p0 + sub; px < p1; px += 2 * nsub) {
      const bool two = px + nsub < p1;
      const float4 g4 = *reinterpret_cast<const float4*>(gp + size_t(px) * p.C);
      const float4 x4 = *reinterpret_cast<const float4*>(xp + size_t(px) * p.C);
      float4 g5 = make_float4(0.f, 0.f, 0.f, 0.f), x5 = g5;
      if (two) { g5 = *reinterpret_cast<const float4*>(gp + size_t(px + nsub) * p.C); x5 = *reinterpret_cast<const float4*>(xp + size_t(px + nsub) * p.C); }
      const float gv[8] = {g4.x, g4.y, g4.z, g4.w, g5.x, g5.y, g5.z, g5.w}, xv[8] = {x4.x, x4.y, x4.z, x4.w, x5.x, x5.y, x5.z, x5.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (k >= 4 && !two) break;
        const int kk = k & 3;
        const float xh = (xv[k] - mean[kk]) * rstd[kk];
        float d = gv[k] * ga[kk];
        if (p.silu) d *= silu_grad_f(fmaf(xh, ga[kk], be[kk]));
        a[kk] += d; b[kk] = fmaf(d, xh, b[kk]);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) csum[sub * p.C + c + k] = make_float2(a[k], b[k]);
  }
  __syncthreads();
  if (threadIdx.x < p.groups) {
    float su = 0.f, sq = 0.f;
    for (int c = threadIdx.x * cpg; c < (threadIdx.x + 1) * cpg; ++c)
      for (int u = 0; u < nsub; ++u) { su += csum[u * p.C + c].x; sq += csum[u * p.C + c].y; }
    p.partial[(size_t(s) * p.nchunks + ch) * p.groups + threadIdx.x] = make_float2(su, sq);
  }
}

// grid (groups, S), 128 threads: (m1, m2) = chunk partial sums / (HW * cpg)
static __global__ void gn_bwd_reduce_kernel(const float2* __restrict__ partial, float2* __restrict__ red, int nchunks, int groups, float inv_n) {
  __shared__ double su[128], sq[128];
  const int g = blockIdx.x, s = blockIdx.y;
  float a = 0.f, b = 0.f;
  for (int k = threadIdx.x; k < nchunks; k += blockDim.x) { const float2 t = partial[(size_t(s) * nchunks + k) * groups + g]; a += t.x; b += t.y; }
  su[threadIdx.x] = a; sq[threadIdx.x] = b;
  __syncthreads();
  for (int o = 64; o; o >>= 1) {
    if (threadIdx.x < o) { su[threadIdx.x] += su[threadIdx.x + o]; sq[threadIdx.x] += sq[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) red[size_t(s) * groups + g] = make_float2(float(su[0] * inv_n), float(sq[0] * inv_n));
}

// grid (ceil(HW / chunk), S), blockDim = quads * nsub (like gn_apply_kernel)
static __global__ void gn_bwd_apply_kernel(const GNBwdParams p) {
  __shared__ float sa[32], sb[32], m1[32], m2[32];
  const int quads = p.C >> 2, cpg = p.C / p.groups;
  const int s = blockIdx.y;
  if (threadIdx.x < p.groups) {
    const float2 t = p.stats[size_t(s) * p.groups + threadIdx.x]; sa[threadIdx.x] = t.x; sb[threadIdx.x] = t.y;
    const float2 r = p.red[size_t(s) * p.groups + threadIdx.x]; m1[threadIdx.x] = r.x; m2[threadIdx.x] = r.y;
  }
  __syncthreads();
  const int p0 = blockIdx.x * p.chunk, p1 = min(p.HW, p0 + p.chunk);
  const int nsub = max(1, int(blockDim.x) / quads);
  const int v = threadIdx.x % quads, sub = threadIdx.x / quads;
  if (sub >= nsub) return;
  const int c = 4 * v;
  float mean[4], rstd[4], ga[4], be[4], r1[4], r2[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int g = (c + k) / cpg;
    mean[k] = sa[g]; rstd[k] = sb[g]; ga[k] = p.gamma[c + k]; be[k] = p.beta[c + k]; r1[k] = m1[g]; r2[k] = m2[g];
  }
  const size_t base = size_t(s) * p.HW * p.C + c;
  for (int px = p0 + sub; px < p1; px += nsub) {
    const size_t o = base + size_t(px) * p.C;
    const float4 g4 = *reinterpret_cast<const float4*>(p.g + o);
    const float4 x4 = *reinterpret_cast<const float4*>(p.x + o);
    float4 ad = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.add) ad = *reinterpret_cast<const float4*>(p.add + o);
    const float gv[4] = {g4.x, g4.y, g4.z, g4.w}, xv[4] = {x4.x, x4.y, x4.z, x4.w}, av[4] = {ad.x, ad.y, ad.z, ad.w};
    float r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float xh = (xv[k] - mean[k]) * rstd[k];
      float d = gv[k] * ga[k];
      if (p.silu) d *= silu_grad_f(fmaf(xh, ga[k], be[k]));
      r[k] = rstd[k] * (d - r1[k] - xh * r2[k]) + av[k];
    }
    if (p.dx) *reinterpret_cast<float4*>(p.dx + o) = make_float4(r[0], r[1], r[2], r[3]);
    if (p.dx16) *reinterpret_cast<uint2*>(p.dx16 + o) = make_uint2(pack_op2(r[0], r[1]), pack_op2(r[2], r[3]));
  }
}

// ------------------------------------------------------------------------------------------------ nearest-2x upsample backward
// g [S][2H][2W][C] fp32 -> dx [S][H][W][C] = sum of the 2x2 block (fp32 and/or 16-bit)
static __global__ void upsample2x_bwd_kernel(const float* __restrict__ g, float* __restrict__ dx, op_t* __restrict__ dx16, int S, int H, int W, int C) {
  const int quads = C >> 2;
  const size_t total = size_t(S) * H * W * quads;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int cq = int(i % quads);
    const int x = int((i / quads) % W), y = int((i / (size_t(quads) * W)) % H), s = int(i / (size_t(quads) * W * H));
    const float* b = g + ((size_t(s) * 2 * H + 2 * y) * 2 * W + 2 * x) * C + 4 * cq;
    const float4 a0 = *reinterpret_cast<const float4*>(b), a1 = *reinterpret_cast<const float4*>(b + C);
    const float4 a2 = *reinterpret_cast<const float4*>(b + size_t(2) * W * C), a3 = *reinterpret_cast<const float4*>(b + size_t(2) * W * C + C);
    const float4 r = make_float4((a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y), (a0.z + a1.z) + (a2.z + a3.z), (a0.w + a1.w) + (a2.w + a3.w));
    const size_t o = ((size_t(s) * H + y) * W + x) * C + 4 * cq;
    if (dx) *reinterpret_cast<float4*>(dx + o) = r;
    if (dx16) *reinterpret_cast<uint2*>(dx16 + o) = make_uint2(pack_op2(r.x, r.y), pack_op2(r.z, r.w));
  }
}

// ------------------------------------------------------------------------------------------------ materialised attention glue
// P[row] = softmax(scale * S[row]) over N columns; one CTA (256 threads) per row; grid (N, batch)
static __global__ void attn_softmax_rows_kernel(const float* __restrict__ S, op_t* __restrict__ P, int N, float scale_log2) {
  __shared__ float red[8];
  const size_t row = size_t(blockIdx.y) * N + blockIdx.x;
  const float* sr = S + row * N;
  op_t* pr = P + row * N;
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < N; j += blockDim.x) mx = fmaxf(mx, sr[j]);
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  const float m = mx * scale_log2;
  float l = 0.f;
  for (int j = threadIdx.x; j < N; j += blockDim.x) l += vae_ex2f(fmaf(sr[j], scale_log2, -m));
#pragma unroll
  for (int o = 16; o; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = l;
  __syncthreads();
  l = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) l += red[w];
  const float inv = 1.f / l;
  for (int j = threadIdx.x; j < N; j += blockDim.x) pr[j] = to_op(vae_ex2f(fmaf(sr[j], scale_log2, -m)) * inv);
}

// dS[row] = scale * P[row] * (dP[row] - sum_j P[row][j] dP[row][j]); one CTA per row; grid (N, batch)
static __global__ void attn_softmax_bwd_rows_kernel(const op_t* __restrict__ P, const float* __restrict__ dP, op_t* __restrict__ dS, int N, float scale) {
  __shared__ float red[8];
  const size_t row = size_t(blockIdx.y) * N + blockIdx.x;
  const op_t* pr = P + row * N;
  const float* dr = dP + row * N;
  float t = 0.f;
  for (int j = threadIdx.x; j < N; j += blockDim.x) t = fmaf(op_to_float(pr[j]), dr[j], t);
#pragma unroll
  for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  t = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
  for (int j = threadIdx.x; j < N; j += blockDim.x) dS[row * N + j] = to_op(scale * op_to_float(pr[j]) * (dr[j] - t));
}

// out[c][r] = in[r][c] for a batch of matrices (16-bit); grid (cols/32, rows/32, batch), block (32, 8)
static __global__ void transpose_h16_kernel(const op_t* __restrict__ in, size_t in_batch, int ld_in, op_t* __restrict__ out, size_t out_batch,
                                            int ld_out, int rows, int cols) {
  __shared__ op_t tile[32][34];
  const op_t* src = in + size_t(blockIdx.z) * in_batch;
  op_t* dst = out + size_t(blockIdx.z) * out_batch;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = src[size_t(r) * ld_in + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[size_t(c) * ld_out + r] = tile[threadIdx.x][j];
  }
}

}  // namespace hedit
