// Glue kernels of the two face-swapping reward networks (reward.cu): ArcFace IR-SE50 identity loss and LPIPS-VGG16, forward and
// input gradient (the reference differentiates both 300 times per image: face-swapping/inversion/h_edit_R.py:109-110,128-129).
// The 3x3 convolutions with >= 64 input channels and the 1x1 shortcut convs run on the tcgen05 implicit-GEMM kernel (gemm.cuh); what is
// here is HBM-bound: BatchNorm (inference affine) / PReLU / ReLU / max-pool / squeeze-excitation / crop + adaptive pooling / the
// 3-channel first layers / the 25088 -> 512 embedding layer / the LPIPS unit-normalised feature distance, each with its backward.
//
// Layouts: activations are NHWC fp32 `[B][P][P][C]` (+ a 16-bit operand copy for the next conv).  IR-SE50 works on 112/56/28/14/7-pixel
// maps; they are stored on P = 128/64/32/16/8 grids whose cells outside the valid V x V corner are ZERO in every tensor a convolution
// reads, so the implicit-GEMM conv's geometry rules (W | 128, level = half of the previous one) hold and the zero cells act as the
// convolution's zero padding.  Convolution OUTPUTS hold garbage outside the valid corner; every kernel below masks it.
#pragma once
#include "ptx.cuh"

namespace hedit {

HEDIT_DEVICE float rw_sigmoid(float x) { return 1.f / (1.f + __expf(-x)); }

// ------------------------------------------------------------------------------------------------ 3-channel first layer (CUDA cores)
// y[b][oy][ox][co] = bias[co] + sum_{ky,kx,ci} w[(ky*3+kx)*3+ci][co] * x[b][oy+ky-1][ox+kx-1][ci];  x [B][V][V][3] fp32 compact,
// y [B][P][P][CO] fp32 (zero outside V x V).  block = 256 threads = 16 pixels x 16 channel quads (CO = 64).
static __global__ void rw_conv_c3_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                             float* __restrict__ y, int V, int P) {
  __shared__ float4 sw[27][16];
  for (int i = threadIdx.x; i < 27 * 16; i += blockDim.x) sw[i / 16][i % 16] = reinterpret_cast<const float4*>(w)[i];
  __syncthreads();
  const int b = blockIdx.y, q = threadIdx.x & 15;
  const int pix = blockIdx.x * 16 + (threadIdx.x >> 4);
  if (pix >= P * P) return;
  const int oy = pix / P, ox = pix - oy * P;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (oy < V && ox < V) {
    a = bias ? reinterpret_cast<const float4*>(bias)[q] : a;
    const float* xb = x + size_t(b) * V * V * 3;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy + ky - 1;
      if (iy < 0 || iy >= V) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox + kx - 1;
        if (ix < 0 || ix >= V) continue;
        const float* xp = xb + (size_t(iy) * V + ix) * 3;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
          const float v = xp[ci];
          const float4 ww = sw[(ky * 3 + kx) * 3 + ci][q];
          a.x = fmaf(v, ww.x, a.x); a.y = fmaf(v, ww.y, a.y); a.z = fmaf(v, ww.z, a.z); a.w = fmaf(v, ww.w, a.w);
        }
      }
    }
  }
  reinterpret_cast<float4*>(y + (size_t(b) * P * P + pix) * 64)[q] = a;
}

// input gradient of the layer above: dx[b][iy][ix][ci] = sum_{ky,kx,co} w[(ky*3+kx)*3+ci][co] * dy[b][iy-ky+1][ix-kx+1][co];
// dy [B][P][P][64] fp32 (zero outside the valid corner), dx [B][V][V][3].  One warp per input pixel: lane <-> 2 output channels.
static __global__ void rw_conv_c3_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx, int V, int P) {
  __shared__ float sw[27][64];
  for (int i = threadIdx.x; i < 27 * 64; i += blockDim.x) sw[i / 64][i % 64] = w[i];
  __syncthreads();
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int pix = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pix >= V * V) return;
  const int iy = pix / V, ix = pix - iy * V;
  float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int oy = iy - ky + 1;
    if (oy < 0 || oy >= V) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ox = ix - kx + 1;
      if (ox < 0 || ox >= V) continue;
      const float2 g = reinterpret_cast<const float2*>(dy + (size_t(b) * P * P + size_t(oy) * P + ox) * 64)[lane];
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const float* wr = sw[(ky * 3 + kx) * 3 + ci];
        acc[ci] = fmaf(g.x, wr[2 * lane], fmaf(g.y, wr[2 * lane + 1], acc[ci]));
      }
    }
  }
#pragma unroll
  for (int ci = 0; ci < 3; ++ci) {
    float v = acc[ci];
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    acc[ci] = v;
  }
  if (lane < 3) dx[(size_t(b) * V * V + pix) * 3 + lane] = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : acc[2]);
}

// ------------------------------------------------------------------------------------------------ ArcFace input: crop + adaptive pool
// IDLoss.extract_feats (arcface/arcface_model.py:41-47): x[:, :, 35:223, 32:220] -> AdaptiveAvgPool2d(112).  Bin o of an adaptive
// pool over L inputs covers [floor(o L / 112), ceil((o + 1) L / 112)).  img NCHW [B][3][R][R] -> out [B][112][112][3].
HEDIT_DEVICE void rw_bin(int o, int L, int n, int& lo, int& hi) { lo = (o * L) / n; hi = ((o + 1) * L + n - 1) / n; }

static __global__ void rw_crop_pool_fwd_kernel(const float* __restrict__ img, float* __restrict__ out, int R, int y0, int x0, int L, int n) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * n * 3) return;
  const int c = i % 3, ox = (i / 3) % n, oy = i / (3 * n);
  int ylo, yhi, xlo, xhi;
  rw_bin(oy, L, n, ylo, yhi); rw_bin(ox, L, n, xlo, xhi);
  const float* p = img + (size_t(b) * 3 + c) * R * R;
  float s = 0.f;
  for (int y = ylo; y < yhi; ++y)
    for (int x = xlo; x < xhi; ++x) s += p[size_t(y0 + y) * R + x0 + x];
  out[size_t(b) * n * n * 3 + i] = s / float((yhi - ylo) * (xhi - xlo));
}

// backward: dimg NCHW [B][3][R][R] (zero outside the crop) from dout [B][n][n][3]; gather form (every crop pixel collects the bins covering it)
static __global__ void rw_crop_pool_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dimg, int R, int y0, int x0, int L, int n) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * R * R) return;
  const int x = i % R, y = (i / R) % R, c = i / (R * R);
  float s = 0.f;
  const int cy = y - y0, cx = x - x0;
  if (cy >= 0 && cy < L && cx >= 0 && cx < L) {
    const int oy0 = max(0, (cy * n) / L - 1), ox0 = max(0, (cx * n) / L - 1);
    for (int oy = oy0; oy < min(n, oy0 + 4); ++oy) {
      int ylo, yhi; rw_bin(oy, L, n, ylo, yhi);
      if (cy < ylo || cy >= yhi) continue;
      for (int ox = ox0; ox < min(n, ox0 + 4); ++ox) {
        int xlo, xhi; rw_bin(ox, L, n, xlo, xhi);
        if (cx < xlo || cx >= xhi) continue;
        s += dout[((size_t(b) * n + oy) * n + ox) * 3 + c] / float((yhi - ylo) * (xhi - xlo));
      }
    }
  }
  dimg[size_t(b) * 3 * R * R + i] = s;
}

// ------------------------------------------------------------------------------------------------ IR-SE50 pointwise kernels
// All: grid.x covers P*P*(C/4) (pixel, channel-quad) items of one sample, grid.y = sample.
struct RwActParams {
  const float* z;            // [B][P][P][C] conv output (garbage outside the valid corner)
  const float* slope;        // PReLU slope [C] or null (identity)
  const float* bn_s; const float* bn_b;   // following BatchNorm affine [C] or null
  float* x;                  // fp32 activation (after PReLU) or null
  op_t* o16;                 // 16-bit operand = bn(prelu(z)) or null
  int P, V, C;
};
static __global__ void rw_act_kernel(const RwActParams p) {
  const int b = blockIdx.y, quads = p.C >> 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.P * p.P * quads) return;
  const int q = i % quads, pix = i / quads, y = pix / p.P, x = pix - y * p.P;
  const size_t off = (size_t(b) * p.P * p.P + pix) * p.C + 4 * q;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f), o = v;
  if (y < p.V && x < p.V) {
    v = *reinterpret_cast<const float4*>(p.z + off);
    if (p.slope) {
      const float4 s = reinterpret_cast<const float4*>(p.slope)[q];
      v.x = v.x > 0.f ? v.x : v.x * s.x; v.y = v.y > 0.f ? v.y : v.y * s.y; v.z = v.z > 0.f ? v.z : v.z * s.z; v.w = v.w > 0.f ? v.w : v.w * s.w;
    }
    o = v;
    if (p.bn_s) {
      const float4 s = reinterpret_cast<const float4*>(p.bn_s)[q], t = reinterpret_cast<const float4*>(p.bn_b)[q];
      o.x = fmaf(v.x, s.x, t.x); o.y = fmaf(v.y, s.y, t.y); o.z = fmaf(v.z, s.z, t.z); o.w = fmaf(v.w, s.w, t.w);
    }
  }
  if (p.x) *reinterpret_cast<float4*>(p.x + off) = v;
  if (p.o16) *reinterpret_cast<uint2*>(p.o16 + off) = make_uint2(pack_op2(o.x, o.y), pack_op2(o.z, o.w));
}

// unit output: out = r * gate[b][c] + shortcut (+ the next unit's BatchNorm operand).  shortcut = sc[b][s*y][s*x][c] on a grid of side s*P.
struct RwCombineParams {
  const float* r; const float* gate;      // [B][P][P][C], [B][C]
  const float* sc; int sc_stride;         // shortcut tensor on a (sc_stride * P) grid
  const float* bn_s; const float* bn_b;   // next unit's first BatchNorm or null
  float* out; op_t* o16;
  int P, V, C;
};
static __global__ void rw_combine_kernel(const RwCombineParams p) {
  const int b = blockIdx.y, quads = p.C >> 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.P * p.P * quads) return;
  const int q = i % quads, pix = i / quads, y = pix / p.P, x = pix - y * p.P;
  const size_t off = (size_t(b) * p.P * p.P + pix) * p.C + 4 * q;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f), o = v;
  if (y < p.V && x < p.V) {
    const float4 r = *reinterpret_cast<const float4*>(p.r + off);
    const float4 g = reinterpret_cast<const float4*>(p.gate + size_t(b) * p.C)[q];
    const int SP = p.sc_stride * p.P;
    const float4 s = *reinterpret_cast<const float4*>(p.sc + ((size_t(b) * SP + size_t(p.sc_stride) * y) * SP + size_t(p.sc_stride) * x) * p.C + 4 * q);
    v = make_float4(fmaf(r.x, g.x, s.x), fmaf(r.y, g.y, s.y), fmaf(r.z, g.z, s.z), fmaf(r.w, g.w, s.w));
    o = v;
    if (p.bn_s) {
      const float4 a = reinterpret_cast<const float4*>(p.bn_s)[q], t = reinterpret_cast<const float4*>(p.bn_b)[q];
      o.x = fmaf(v.x, a.x, t.x); o.y = fmaf(v.y, a.y, t.y); o.z = fmaf(v.z, a.z, t.z); o.w = fmaf(v.w, a.w, t.w);
    }
  }
  *reinterpret_cast<float4*>(p.out + off) = v;
  if (p.o16) *reinterpret_cast<uint2*>(p.o16 + off) = make_uint2(pack_op2(o.x, o.y), pack_op2(o.z, o.w));
}

// strided 16-bit copy: o[b][y][x][c] = x[b][2y][2x][c] (operand of the stride-2 1x1 shortcut conv); x on a 2P grid
static __global__ void rw_gather2_kernel(const float* __restrict__ x, op_t* __restrict__ o, int P, int C) {
  const int b = blockIdx.y, quads = C >> 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * P * quads) return;
  const int q = i % quads, pix = i / quads, y = pix / P, xx = pix - y * P;
  const float4 v = *reinterpret_cast<const float4*>(x + ((size_t(b) * 2 * P + 2 * y) * 2 * P + 2 * xx) * C + 4 * q);
  *reinterpret_cast<uint2*>(o + (size_t(b) * P * P + pix) * C + 4 * q) = make_uint2(pack_op2(v.x, v.y), pack_op2(v.z, v.w));
}

// per-sample channel sums over the valid corner: partial[b][chunk][c] = sum_{pixels of chunk} a * (w ? w : 1); chunk = RW_POOL_ROWS rows.
// blockDim = 256 = (C/4 quads) x nsub pixel lanes; fixed summation order (no atomics).
constexpr int RW_POOL_ROWS = 2;
// sq (optional): per-channel sums of a^2 in the same layout (the backward's per-sample gradient magnitude, see rw_se_bwd_kernel).
static __global__ void rw_pool_partial_kernel(const float* __restrict__ a, const float* __restrict__ w, float* __restrict__ partial,
                                              float* __restrict__ sq, int P, int V, int C) {
  extern __shared__ float4 rw_sm[];       // [2][nsub][quads]
  const int b = blockIdx.y, ch = blockIdx.x, quads = C >> 2, nsub = blockDim.x / quads;
  const int q = threadIdx.x % quads, sub = threadIdx.x / quads;
  const int y0 = ch * RW_POOL_ROWS, y1 = min(V, y0 + RW_POOL_ROWS);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s;
  if (sub < nsub) {
    const int npx = (y1 - y0) * V;
    for (int k = sub; k < npx; k += nsub) {
      const int y = y0 + k / V, x = k % V;
      const size_t off = ((size_t(b) * P + y) * P + x) * C + 4 * q;
      float4 v = *reinterpret_cast<const float4*>(a + off);
      s2.x = fmaf(v.x, v.x, s2.x); s2.y = fmaf(v.y, v.y, s2.y); s2.z = fmaf(v.z, v.z, s2.z); s2.w = fmaf(v.w, v.w, s2.w);
      if (w) { const float4 u = *reinterpret_cast<const float4*>(w + off); v.x *= u.x; v.y *= u.y; v.z *= u.z; v.w *= u.w; }
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    rw_sm[sub * quads + q] = s;
    rw_sm[(nsub + sub) * quads + q] = s2;
  }
  __syncthreads();
  if (threadIdx.x < quads) {
    float4 t = rw_sm[q], t2 = rw_sm[nsub * quads + q];
    for (int k = 1; k < nsub; ++k) {
      const float4 u = rw_sm[k * quads + q], u2 = rw_sm[(nsub + k) * quads + q];
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
      t2.x += u2.x; t2.y += u2.y; t2.z += u2.z; t2.w += u2.w;
    }
    reinterpret_cast<float4*>(partial + (size_t(b) * gridDim.x + ch) * C)[q] = t;
    if (sq) reinterpret_cast<float4*>(sq + (size_t(b) * gridDim.x + ch) * C)[q] = t2;
  }
}

// squeeze-excitation gate (helpers.py SEModule): m = mean(r); h = relu(W1 m); gate = sigmoid(W2 h).  grid B, block 128.  C <= 512.
static __global__ void rw_se_fc_kernel(const float* __restrict__ partial, int nch, const float* __restrict__ w1, const float* __restrict__ w2,
                                       float* __restrict__ hbuf, float* __restrict__ gate, int C, float inv_n) {
  __shared__ float m[512], h[32], red[512];
  const int b = blockIdx.x, Cr = C / 16;
  {   // blockDim = 512 = C channels x (512 / C) chunk lanes; lanes are combined in a fixed order
    const int lanes = blockDim.x / C, c = threadIdx.x % C, kl = threadIdx.x / C;
    float s = 0.f;
    for (int k = kl; k < nch; k += lanes) s += partial[(size_t(b) * nch + k) * C + c];
    red[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < C) {
      float t = red[c];
      for (int l = 1; l < lanes; ++l) t += red[l * C + c];
      m[c] = t * inv_n;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = warp; j < Cr; j += blockDim.x >> 5) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(w1[size_t(j) * C + c], m[c], s);
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) { h[j] = fmaxf(s, 0.f); hbuf[size_t(b) * 32 + j] = h[j]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < Cr; ++j) s = fmaf(w2[size_t(c) * Cr + j], h[j], s);
    gate[size_t(b) * C + c] = rw_sigmoid(s);
  }
}

// backward of the gate: dgate = sum partial (= sum_pixels dout * r) -> dmean[b][c] = (W1^T ((W2^T (dgate g (1 - g))) * (h > 0))) / n.
// Also gscale[b] = 1 / rms(dout[b]) from the a^2 partials: inside the unit the back-propagated gradient is carried as 16-bit conv
// operands, so it is normalised per sample on entry (rw_dr_kernel) and restored on exit (rw_unit_in_bwd_kernel); the unit is linear in it.
static __global__ void rw_se_bwd_kernel(const float* __restrict__ partial, const float* __restrict__ sq, int nch, const float* __restrict__ w1,
                                        const float* __restrict__ w2, const float* __restrict__ hbuf, const float* __restrict__ gate,
                                        float* __restrict__ dmean, float* __restrict__ gscale, int C, float inv_n) {
  __shared__ float dz[512], dh[32], red[512], red2[512];
  const int b = blockIdx.x, Cr = C / 16;
  {   // blockDim = 512 = C channels x (512 / C) chunk lanes
    const int lanes = blockDim.x / C, c = threadIdx.x % C, kl = threadIdx.x / C;
    float s = 0.f, ss = 0.f;
    for (int k = kl; k < nch; k += lanes) { s += partial[(size_t(b) * nch + k) * C + c]; ss += sq[(size_t(b) * nch + k) * C + c]; }
    red[threadIdx.x] = s; red2[threadIdx.x] = ss;
    __syncthreads();
    if (threadIdx.x < C) {
      float t = red[c], t2 = red2[c];
      for (int l = 1; l < lanes; ++l) { t += red[l * C + c]; t2 += red2[l * C + c]; }
      const float g = gate[size_t(b) * C + c];
      dz[c] = t * g * (1.f - g);
      red2[c] = t2;
    }
    __syncthreads();
    if (threadIdx.x < 32) {      // sum of the per-channel squares in a fixed order: lane strides, then a shuffle tree
      float t = 0.f;
      for (int cc = threadIdx.x; cc < C; cc += 32) t += red2[cc];
#pragma unroll
      for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (threadIdx.x == 0) {
        const float ms = t * inv_n / float(C);
        gscale[b] = ms > 1e-36f ? rsqrtf(ms) : 1.f;
      }
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = warp; j < Cr; j += blockDim.x >> 5) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(w2[size_t(c) * Cr + j], dz[c], s);
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) dh[j] = hbuf[size_t(b) * 32 + j] > 0.f ? s : 0.f;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < Cr; ++j) s = fmaf(w1[size_t(j) * C + c], dh[j], s);
    dmean[size_t(b) * C + c] = s * inv_n;
  }
}

// gradient entering the unit's second conv: dr = dout * gate + dmean, as the 16-bit operand of the conv dgrad.  up = 1: written at the
// even cells of a (2P) grid with zeros elsewhere (input gradient of a stride-2 conv = stride-1 dgrad over the zero-inserted gradient).
static __global__ void rw_dr_kernel(const float* __restrict__ dout, const float* __restrict__ gate, const float* __restrict__ dmean,
                                    const float* __restrict__ gscale, op_t* __restrict__ o16, int P, int V, int C, int up) {
  const int b = blockIdx.y, quads = C >> 2, PO = up ? 2 * P : P;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= PO * PO * quads) return;
  const int q = i % quads, pix = i / quads, yo = pix / PO, xo = pix - yo * PO;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool on = up ? !((yo | xo) & 1) : true;
  const int y = up ? yo >> 1 : yo, x = up ? xo >> 1 : xo;
  if (on && y < V && x < V) {
    const float4 d = *reinterpret_cast<const float4*>(dout + ((size_t(b) * P + y) * P + x) * C + 4 * q);
    const float4 g = reinterpret_cast<const float4*>(gate + size_t(b) * C)[q], m = reinterpret_cast<const float4*>(dmean + size_t(b) * C)[q];
    const float gs = gscale[b];
    v = make_float4(fmaf(d.x, g.x, m.x) * gs, fmaf(d.y, g.y, m.y) * gs, fmaf(d.z, g.z, m.z) * gs, fmaf(d.w, g.w, m.w) * gs);
  }
  *reinterpret_cast<uint2*>(o16 + (size_t(b) * PO * PO + pix) * C + 4 * q) = make_uint2(pack_op2(v.x, v.y), pack_op2(v.z, v.w));
}

// PReLU backward: d = g * (z > 0 ? 1 : slope) masked -> 16-bit operand and / or fp32
static __global__ void rw_prelu_bwd_kernel(const float* __restrict__ g, const float* __restrict__ z, const float* __restrict__ slope,
                                           op_t* __restrict__ o16, float* __restrict__ o32, int P, int V, int C) {
  const int b = blockIdx.y, quads = C >> 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * P * quads) return;
  const int q = i % quads, pix = i / quads, y = pix / P, x = pix - y * P;
  const size_t off = (size_t(b) * P * P + pix) * C + 4 * q;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (y < V && x < V) {
    const float4 d = *reinterpret_cast<const float4*>(g + off), zz = *reinterpret_cast<const float4*>(z + off);
    const float4 s = reinterpret_cast<const float4*>(slope)[q];
    v = make_float4(zz.x > 0.f ? d.x : d.x * s.x, zz.y > 0.f ? d.y : d.y * s.y, zz.z > 0.f ? d.z : d.z * s.z, zz.w > 0.f ? d.w : d.w * s.w);
  }
  if (o16) *reinterpret_cast<uint2*>(o16 + off) = make_uint2(pack_op2(v.x, v.y), pack_op2(v.z, v.w));
  if (o32) *reinterpret_cast<float4*>(o32 + off) = v;
}

// gradient leaving the unit: dx = da * bn_s + (shortcut gradient sg[b][y/s][x/s][c] at cells with y % s == x % s == 0), masked;
// sg lives on a (P / s) grid.  Optional 16-bit copy (operand of the previous unit's shortcut-conv dgrad).
struct RwUnitInBwdParams {
  const float* da; const float* bn_s;
  const float* gscale; int sg_scaled;     // da (and sg when sg_scaled) carry the unit's per-sample normalisation gscale[b]
  const float* sg; int sg_stride;
  float* dx; op_t* dx16;
  int P, V, C;
};
static __global__ void rw_unit_in_bwd_kernel(const RwUnitInBwdParams p) {
  const int b = blockIdx.y, quads = p.C >> 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.P * p.P * quads) return;
  const int q = i % quads, pix = i / quads, y = pix / p.P, x = pix - y * p.P;
  const size_t off = (size_t(b) * p.P * p.P + pix) * p.C + 4 * q;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (y < p.V && x < p.V) {
    const float4 d = *reinterpret_cast<const float4*>(p.da + off);
    const float4 s = reinterpret_cast<const float4*>(p.bn_s)[q];
    const float inv = 1.f / p.gscale[b];
    v = make_float4(d.x * s.x * inv, d.y * s.y * inv, d.z * s.z * inv, d.w * s.w * inv);
    const int st = p.sg_stride;
    if (st == 1 || !((y | x) & 1)) {
      const int SP = p.P / st;
      const float4 g = *reinterpret_cast<const float4*>(p.sg + ((size_t(b) * SP + y / st) * SP + x / st) * p.C + 4 * q);
      const float m = p.sg_scaled ? inv : 1.f;
      v.x = fmaf(g.x, m, v.x); v.y = fmaf(g.y, m, v.y); v.z = fmaf(g.z, m, v.z); v.w = fmaf(g.w, m, v.w);
    }
  }
  *reinterpret_cast<float4*>(p.dx + off) = v;
  if (p.dx16) *reinterpret_cast<uint2*>(p.dx16 + off) = make_uint2(pack_op2(v.x, v.y), pack_op2(v.z, v.w));
}

// fp32 -> 16-bit copy (valid corner only, zero elsewhere), optionally times the per-sample gscale[b]
static __global__ void rw_cast_masked_kernel(const float* __restrict__ x, const float* __restrict__ gscale, op_t* __restrict__ o, int P, int V, int C) {
  const int b = blockIdx.y, quads = C >> 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * P * quads) return;
  const int q = i % quads, pix = i / quads, y = pix / P, xx = pix - y * P;
  const size_t off = (size_t(b) * P * P + pix) * C + 4 * q;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (y < V && xx < V) {
    v = *reinterpret_cast<const float4*>(x + off);
    const float gs = gscale ? gscale[b] : 1.f;
    v.x *= gs; v.y *= gs; v.z *= gs; v.w *= gs;
  }
  *reinterpret_cast<uint2*>(o + off) = make_uint2(pack_op2(v.x, v.y), pack_op2(v.z, v.w));
}

// ------------------------------------------------------------------------------------------------ IR-SE50 embedding layer
// output_layer (model_irse.py:23-27) with both BatchNorms folded: f[b][o] = bias[o] + sum_{pos < 49, c < 512} W[o][pos][c] * x[b][pos][c];
// x [B][8][8][512] fp32 (7 x 7 valid), W 16-bit [512][49][512].  B <= RW_HEAD_MAXB per launch.
constexpr int RW_HEAD_MAXB = 8;
// one block (256 threads) per 2 outputs: the 25088-long dot products are split over the 8 warps and combined in a fixed order
static __global__ void rw_head_fwd_kernel(const float* __restrict__ x, const op_t* __restrict__ W, const float* __restrict__ bias,
                                          float* __restrict__ f, int B) {
  __shared__ float part[8][2][RW_HEAD_MAXB];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o0 = blockIdx.x * 2;
  float acc[2][RW_HEAD_MAXB];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int b = 0; b < RW_HEAD_MAXB; ++b) acc[j][b] = 0.f;
  constexpr int KW = 49 * 512 / 8;       // 3136 inputs per warp
#pragma unroll 2
  for (int k = warp * KW + lane * 4; k < (warp + 1) * KW; k += 128) {
    const int pos = k >> 9, c = k & 511, cell = (pos / 7) * 8 + (pos % 7);
    float wv[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint2 u = *reinterpret_cast<const uint2*>(W + size_t(o0 + j) * 49 * 512 + k);
      const float2 a = op2_to_float2(u.x), bq = op2_to_float2(u.y);
      wv[j][0] = a.x; wv[j][1] = a.y; wv[j][2] = bq.x; wv[j][3] = bq.y;
    }
#pragma unroll
    for (int b = 0; b < RW_HEAD_MAXB; ++b) {
      if (b < B) {
        const float4 xv = *reinterpret_cast<const float4*>(x + (size_t(b) * 64 + cell) * 512 + c);
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[j][b] = fmaf(wv[j][0], xv.x, fmaf(wv[j][1], xv.y, fmaf(wv[j][2], xv.z, fmaf(wv[j][3], xv.w, acc[j][b]))));
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int b = 0; b < RW_HEAD_MAXB; ++b) {
      float v = acc[j][b];
#pragma unroll
      for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) part[warp][j][b] = v;
    }
  __syncthreads();
  if (threadIdx.x < 2 * RW_HEAD_MAXB) {
    const int j = threadIdx.x / RW_HEAD_MAXB, b = threadIdx.x % RW_HEAD_MAXB;
    if (b < B) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += part[w][j][b];
      f[size_t(b) * 512 + o0 + j] = v + bias[o0 + j];
    }
  }
}

// dx[b][cell][c] = sum_o df[b][o] W[o][pos][c] (zero at the padded cells).  grid (64 cells, ceil(B / MAXB)), block 256 = channel pairs.
static __global__ void rw_head_bwd_kernel(const float* __restrict__ df, const op_t* __restrict__ W, float* __restrict__ dx, int B) {
  __shared__ float sdf[RW_HEAD_MAXB][512];
  const int cell = blockIdx.x, cy = cell >> 3, cx = cell & 7, b0 = blockIdx.y * RW_HEAD_MAXB, nb = min(RW_HEAD_MAXB, B - b0);
  const int c = 2 * threadIdx.x;
  if (cy >= 7 || cx >= 7) {
    for (int b = 0; b < nb; ++b) *reinterpret_cast<float2*>(dx + (size_t(b0 + b) * 64 + cell) * 512 + c) = make_float2(0.f, 0.f);
    return;
  }
  for (int i = threadIdx.x; i < nb * 512; i += blockDim.x) sdf[i >> 9][i & 511] = df[size_t(b0) * 512 + i];
  __syncthreads();
  const int pos = cy * 7 + cx;
  float2 acc[RW_HEAD_MAXB];
#pragma unroll
  for (int b = 0; b < RW_HEAD_MAXB; ++b) acc[b] = make_float2(0.f, 0.f);
#pragma unroll 16
  for (int o = 0; o < 512; ++o) {
    const float2 w = op2_to_float2(*reinterpret_cast<const uint32_t*>(W + (size_t(o) * 49 + pos) * 512 + c));
#pragma unroll
    for (int b = 0; b < RW_HEAD_MAXB; ++b)
      if (b < nb) { const float d = sdf[b][o]; acc[b].x = fmaf(d, w.x, acc[b].x); acc[b].y = fmaf(d, w.y, acc[b].y); }
  }
  for (int b = 0; b < nb; ++b) *reinterpret_cast<float2*>(dx + (size_t(b0 + b) * 64 + cell) * 512 + c) = acc[b];
}

// identity loss (arcface_model.py:49-70): 1 - cos(ref, f); every normalisation of the reference chain (l2_norm, F.normalize,
// cosine_similarity) is scale-invariant, so loss = 1 - <rhat, f> / |f| and dloss/df = -(rhat - cos * f / |f|) / |f|.
// ref_hat [512] unit vector.  grid B, block 128.  mode 0: loss + gradient; mode 1: write the unit feature to `out_unit` (reference set-up).
static __global__ void rw_cos_loss_kernel(const float* __restrict__ f, const float* __restrict__ ref_hat, float* __restrict__ loss,
                                          float* __restrict__ df, float* __restrict__ out_unit, int mode) {
  __shared__ float red[2][4];
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float v[4], nn = 0.f, dt = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v[k] = f[size_t(b) * 512 + threadIdx.x + 128 * k];
    nn = fmaf(v[k], v[k], nn);
    if (mode == 0) dt = fmaf(v[k], ref_hat[threadIdx.x + 128 * k], dt);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) { nn += __shfl_xor_sync(0xffffffffu, nn, o); dt += __shfl_xor_sync(0xffffffffu, dt, o); }
  if (lane == 0) { red[0][warp] = nn; red[1][warp] = dt; }
  __syncthreads();
  nn = red[0][0] + red[0][1] + red[0][2] + red[0][3];
  dt = red[1][0] + red[1][1] + red[1][2] + red[1][3];
  const float inv = rsqrtf(fmaxf(nn, 1e-24f));
  if (mode == 1) {
#pragma unroll
    for (int k = 0; k < 4; ++k) out_unit[size_t(b) * 512 + threadIdx.x + 128 * k] = v[k] * inv;
    return;
  }
  const float cs = dt * inv;
  if (threadIdx.x == 0 && loss) loss[b] = 1.f - cs;
#pragma unroll
  for (int k = 0; k < 4; ++k) df[size_t(b) * 512 + threadIdx.x + 128 * k] = -(ref_hat[threadIdx.x + 128 * k] - cs * v[k] * inv) * inv;
}

// ------------------------------------------------------------------------------------------------ LPIPS-VGG16
// ScalingLayer of the lpips package: (x - shift) / scale per channel; img NCHW [B][3][R][R] -> [B][R][R][3]
static __global__ void rw_vgg_prep_kernel(const float* __restrict__ img, float* __restrict__ out, int R, float3 shift, float3 inv_scale) {
  const int b = blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= R * R) return;
  const float* p = img + size_t(b) * 3 * R * R + pix;
  float* o = out + (size_t(b) * R * R + pix) * 3;
  o[0] = (p[0] - shift.x) * inv_scale.x; o[1] = (p[size_t(R) * R] - shift.y) * inv_scale.y; o[2] = (p[size_t(2) * R * R] - shift.z) * inv_scale.z;
}
// and its backward: dimg NCHW = d[B][R][R][3] * inv_scale / (gs[b] * tfin)   (gs, tfin: the backward's gradient normalisation)
static __global__ void rw_vgg_unprep_kernel(const float* __restrict__ d, float* __restrict__ dimg, int R, float3 inv_scale_in,
                                            const float* __restrict__ gs, float tfin) {
  const int b = blockIdx.y;
  const float un = 1.f / (gs[b] * tfin);
  const float3 inv_scale = make_float3(inv_scale_in.x * un, inv_scale_in.y * un, inv_scale_in.z * un);
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= R * R) return;
  const float* p = d + (size_t(b) * R * R + pix) * 3;
  float* o = dimg + size_t(b) * 3 * R * R + pix;
  o[0] = p[0] * inv_scale.x; o[size_t(R) * R] = p[1] * inv_scale.y; o[size_t(2) * R * R] = p[2] * inv_scale.z;
}

// a16 = relu(z), optionally followed by a 2x2 max-pool: z [B][H][W][C] fp32 -> o [B][H/pool][W/pool][C] 16-bit
static __global__ void rw_vgg_act_kernel(const float* __restrict__ z, op_t* __restrict__ o, int H, int W, int C, int pool) {
  const int b = blockIdx.y, quads = C >> 2, Ho = H / pool, Wo = W / pool;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Ho * Wo * quads) return;
  const int q = i % quads, pix = i / quads, y = pix / Wo, x = pix - y * Wo;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);       // relu floor
  for (int dy = 0; dy < pool; ++dy)
    for (int dx = 0; dx < pool; ++dx) {
      const float4 t = *reinterpret_cast<const float4*>(z + ((size_t(b) * H + pool * y + dy) * W + pool * x + dx) * C + 4 * q);
      v.x = fmaxf(v.x, t.x); v.y = fmaxf(v.y, t.y); v.z = fmaxf(v.z, t.z); v.w = fmaxf(v.w, t.w);
    }
  *reinterpret_cast<uint2*>(o + (size_t(b) * Ho * Wo + pix) * C + 4 * q) = make_uint2(pack_op2(v.x, v.y), pack_op2(v.z, v.w));
}

// LPIPS tap (lpips: normalize_tensor eps 1e-10, squared difference, non-negative 1x1 `lin` weights, spatial mean): one warp per pixel.
//   f = relu(z); n = f / (|f| + eps); loss_b += sum_c w_c (n_c - ref_c)^2 / HW
//   mode 1: write n to `nref` (source-image set-up).  mode 0: block loss partials + gtap = d loss / d f (before the ReLU mask) + block
//   partials of sum gtap^2 (the per-sample gradient magnitude, see rw_lpips_scale_kernel).
// ref [Bref][HW][C] with Bref = 1 (broadcast) or B.  block = 256 threads = 8 pixels; partial[b][blockIdx.x].
template <int C>
static __global__ void rw_lpips_tap_kernel(const float* __restrict__ z, float* __restrict__ nref, const float* __restrict__ ref, int ref_bstride,
                                           const float* __restrict__ lin, float* __restrict__ gtap, float* __restrict__ partial,
                                           float* __restrict__ gsq, int HW, int mode) {
  constexpr int NV = C / 64;             // float2 per lane
  __shared__ float sl[8], sg2[8];
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pix = blockIdx.x * 8 + warp;
  float lsum = 0.f, g2 = 0.f;
  if (pix < HW) {
    const size_t off = (size_t(b) * HW + pix) * C;
    float2 f[NV];
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      f[k] = *reinterpret_cast<const float2*>(z + off + 2 * (lane + 32 * k));
      f[k].x = fmaxf(f[k].x, 0.f); f[k].y = fmaxf(f[k].y, 0.f);
      ss = fmaf(f[k].x, f[k].x, fmaf(f[k].y, f[k].y, ss));
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float r = sqrtf(ss), inv = 1.f / (r + 1e-10f);
    if (mode == 1) {
#pragma unroll
      for (int k = 0; k < NV; ++k) *reinterpret_cast<float2*>(nref + off + 2 * (lane + 32 * k)) = make_float2(f[k].x * inv, f[k].y * inv);
    } else {
      const float* rp = ref + size_t(b) * ref_bstride + size_t(pix) * C;
      const float ihw = 1.f / float(HW);
      float2 dn[NV];
      float dot = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = 2 * (lane + 32 * k);
        const float2 rr = *reinterpret_cast<const float2*>(rp + c), w = *reinterpret_cast<const float2*>(lin + c);
        const float dx = f[k].x * inv - rr.x, dy = f[k].y * inv - rr.y;
        lsum = fmaf(w.x * dx, dx, fmaf(w.y * dy, dy, lsum));
        dn[k] = make_float2(2.f * w.x * dx * ihw, 2.f * w.y * dy * ihw);
        dot = fmaf(dn[k].x, f[k].x, fmaf(dn[k].y, f[k].y, dot));
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) { lsum += __shfl_xor_sync(0xffffffffu, lsum, o); dot += __shfl_xor_sync(0xffffffffu, dot, o); }
      lsum *= ihw;
      // d n_i / d f_j = delta_ij / (r + eps) - f_i f_j / (r (r + eps)^2)
      const float k2 = r > 0.f ? dot * inv * inv / r : 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const float2 gv = make_float2(dn[k].x * inv - f[k].x * k2, dn[k].y * inv - f[k].y * k2);
        g2 = fmaf(gv.x, gv.x, fmaf(gv.y, gv.y, g2));
        *reinterpret_cast<float2*>(gtap + off + 2 * (lane + 32 * k)) = gv;
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) g2 += __shfl_xor_sync(0xffffffffu, g2, o);
    }
  }
  if (mode == 0) {
    if (lane == 0) { sl[warp] = lsum; sg2[warp] = g2; }
    __syncthreads();
    if (threadIdx.x == 0) {
      partial[size_t(b) * gridDim.x + blockIdx.x] = ((sl[0] + sl[1]) + (sl[2] + sl[3])) + ((sl[4] + sl[5]) + (sl[6] + sl[7]));
      gsq[size_t(b) * gridDim.x + blockIdx.x] = ((sg2[0] + sg2[1]) + (sg2[2] + sg2[3])) + ((sg2[4] + sg2[5]) + (sg2[6] + sg2[7]));
    }
  }
}

// Per sample: loss[b] = sum over the five taps of their block partials, and the gradient normalisation
//   gs[b] = 1 / max_k (rms(gtap_k[b]) * tstat[k])
// The VGG backward carries gradients as 16-bit conv operands: every tensor is stored as (true gradient) x gs[b] x T_i with STATIC
// per-layer factors T_i (estimated dgrad gains, LpipsNet::finalize) and this one dynamic per-sample factor, so that each tap's
// contribution enters with rms <= 1.  grid B, block 256.  partial / gsq: tap k owns [B][nblk[k]] at offset boff[k] * B.
struct RwLpipsScaleParams { int nblk[5], boff[5]; float tstat[5], inv_count[5]; };
static __global__ void rw_lpips_scale_kernel(const float* __restrict__ partial, const float* __restrict__ gsq, RwLpipsScaleParams p, int B,
                                             float* __restrict__ loss, float* __restrict__ gs) {
  __shared__ double r1[256], r2[256];
  const int b = blockIdx.x;
  double lsum = 0.0;
  float worst = 0.f;
  for (int k = 0; k < 5; ++k) {
    double a = 0.0, q = 0.0;
    const size_t base = size_t(p.boff[k]) * B + size_t(b) * p.nblk[k];
    for (int i = threadIdx.x; i < p.nblk[k]; i += blockDim.x) { a += double(partial[base + i]); q += double(gsq[base + i]); }
    r1[threadIdx.x] = a; r2[threadIdx.x] = q;
    __syncthreads();
    for (int o = blockDim.x >> 1; o; o >>= 1) {
      if (threadIdx.x < o) { r1[threadIdx.x] += r1[threadIdx.x + o]; r2[threadIdx.x] += r2[threadIdx.x + o]; }
      __syncthreads();
    }
    lsum += r1[0];
    worst = fmaxf(worst, sqrtf(float(r2[0]) * p.inv_count[k]) * p.tstat[k]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (loss) loss[b] = float(lsum);
    gs[b] = worst > 1e-30f ? 1.f / worst : 1.f;
  }
}

// gradient w.r.t. a conv output z (pre-ReLU): dz = (da routed through the optional 2x2 max-pool + gtap) * (z > 0) -> 16-bit (+ fp32).
//   da [B][H/pool][W/pool][C] fp32 gradient w.r.t. the next conv's operand (or null for the last layer), times the static factor cda;
//   gtap [B][H][W][C] or null, times gs[b] * ttap (rw_lpips_scale_kernel).
// thread <-> (coarse pixel, channel quad); the pool routes to the FIRST maximum of relu(z) in window order, like torch.
template <int pool>
static __global__ void rw_vgg_bwd_kernel(const float* __restrict__ da, const float* __restrict__ z, const float* __restrict__ gtap,
                                         op_t* __restrict__ o16, float* __restrict__ o32, int H, int W, int C, float cda,
                                         const float* __restrict__ gs, float ttap) {
  const int b = blockIdx.y, quads = C >> 2, Ho = H / pool, Wo = W / pool;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Ho * Wo * quads) return;
  const int q = i % quads, pix = i / quads, y = pix / Wo, x = pix - y * Wo;
  float g[4] = {0.f, 0.f, 0.f, 0.f};
  if (da) {
    const float4 t = *reinterpret_cast<const float4*>(da + (size_t(b) * Ho * Wo + pix) * C + 4 * q);
    g[0] = t.x * cda; g[1] = t.y * cda; g[2] = t.z * cda; g[3] = t.w * cda;
  }
  const float tsc = gtap ? gs[b] * ttap : 0.f;
  constexpr int n = pool * pool;
  float zz[n][4];
#pragma unroll
  for (int k = 0; k < n; ++k) {
    const float4 t = *reinterpret_cast<const float4*>(z + ((size_t(b) * H + pool * y + k / pool) * W + pool * x + k % pool) * C + 4 * q);
    zz[k][0] = t.x; zz[k][1] = t.y; zz[k][2] = t.z; zz[k][3] = t.w;
  }
  int arg[4] = {0, 0, 0, 0};
  if (pool == 2) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float best = fmaxf(zz[0][c], 0.f);
#pragma unroll
      for (int k = 1; k < 4; ++k) { const float v = fmaxf(zz[k][c], 0.f); if (v > best) { best = v; arg[c] = k; } }
    }
  }
#pragma unroll
  for (int k = 0; k < n; ++k) {
    const size_t off = ((size_t(b) * H + pool * y + k / pool) * W + pool * x + k % pool) * C + 4 * q;
    float4 t = gtap ? *reinterpret_cast<const float4*>(gtap + off) : make_float4(0.f, 0.f, 0.f, 0.f);
    float v[4] = {t.x * tsc, t.y * tsc, t.z * tsc, t.w * tsc};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (arg[c] == k) v[c] += g[c];
      if (!(zz[k][c] > 0.f)) v[c] = 0.f;
    }
    if (o16) *reinterpret_cast<uint2*>(o16 + off) = make_uint2(pack_op2(v[0], v[1]), pack_op2(v[2], v[3]));
    if (o32) *reinterpret_cast<float4*>(o32 + off) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// ------------------------------------------------------------------------------------------------ weight preparation
// BatchNorm (inference) -> affine: s = gamma / sqrt(var + eps), t = beta - mean * s
static __global__ void rw_bn_affine_kernel(const float* __restrict__ g, const float* __restrict__ b, const float* __restrict__ mean,
                                           const float* __restrict__ var, float eps, float* __restrict__ s, float* __restrict__ t, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = (g ? g[c] : 1.f) * rsqrtf(var[c] + eps);
  s[c] = sc; t[c] = (b ? b[c] : 0.f) - mean[c] * sc;
}
// w[o][...] *= s[o] (folds a following BatchNorm into a conv / linear weight); n_per_o elements per output
static __global__ void rw_scale_rows_kernel(float* __restrict__ w, const float* __restrict__ s, size_t n_per_o, size_t total) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) w[i] *= s[i / n_per_o];
}
// out[0] = sum w^2 (one block)
static __global__ void rw_sumsq_kernel(const float* __restrict__ w, size_t n, float* __restrict__ out) {
  __shared__ double red[1024];
  double a = 0.0;
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) a += double(w[i]) * double(w[i]);
  red[threadIdx.x] = a;
  __syncthreads();
  for (int s = blockDim.x >> 1; s; s >>= 1) { if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s]; __syncthreads(); }
  if (threadIdx.x == 0) out[0] = float(red[0]);
}
// first-layer weights [CO][3][3][3] (O, I, ky, kx) -> fp32 [27][CO] with k = (ky*3+kx)*3+ci
static __global__ void rw_cvt_c3_kernel(const float* __restrict__ src, float* __restrict__ dst, int CO) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= CO * 27) return;
  const int tap = i % 9, ci = (i / 9) % 3, o = i / 27;
  dst[(tap * 3 + ci) * CO + o] = src[i];
}
// embedding layer: Linear weight [512][512*49] (input index c*49 + pos, the NCHW flatten) with the preceding BatchNorm2d (s2, t2 per c) and
// the following BatchNorm1d (s1, t1 per o) folded -> 16-bit [512][49][512] + fp32 bias[o] = s1 (b + sum W t2) + t1
static __global__ void rw_cvt_head_kernel(const float* __restrict__ Wl, const float* __restrict__ bl, const float* __restrict__ s2,
                                          const float* __restrict__ t2, const float* __restrict__ s1, const float* __restrict__ t1,
                                          op_t* __restrict__ Wo, float* __restrict__ bo) {
  __shared__ double red[256];
  const int o = blockIdx.x;
  double acc = 0.0;
  for (int k = threadIdx.x; k < 512 * 49; k += blockDim.x) {
    const int c = k / 49, pos = k % 49;
    const float w = Wl[size_t(o) * 512 * 49 + k];
    acc += double(w) * double(t2[c]);
    Wo[(size_t(o) * 49 + pos) * 512 + c] = to_op(w * s2[c] * s1[o]);
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x >> 1; s; s >>= 1) { if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s]; __syncthreads(); }
  if (threadIdx.x == 0) bo[o] = s1[o] * (bl[o] + float(red[0])) + t1[o];
}

}  // namespace hedit
