// Kernels of the CLIP-Gram style reward (forward + input gradient): the image branch the reference's style path differentiates
// (text-guided-n-style/clip_guidance/base_clip.py:55-66 get_gram_matrix_residual on clip/model.py:339-359 encode_image_with_features).
// Linear layers / patch embedding run on the tcgen05 GEMM (gemm.cuh); these are the small fp32 pieces around them.  Every reduction
// has a fixed order (no atomics).
#pragma once
#include "ptx.cuh"

namespace hedit {

// ------------------------------------------------------------------------------------------------ bicubic resize (separable, CSR taps)
// One pass along x or y: out[c][..] = sum_k w[k] * in[c][.. idx[k] ..] over the taps rowptr[o] .. rowptr[o+1] of output index o.
// The host builds the taps of torch's bicubic (A = -0.75, align_corners = False, border clamp) and their transpose (backward).
// along_x = 1: in [C][H][Win] -> out [C][H][Wout] ; along_x = 0: in [C][Hin][W] -> out [C][Hout][W].
// post: out = (val - mean[c]) * inv_std[c] (forward, last pass) ; pre_scale: val *= inv_std[c] (backward, first pass)
static __global__ void sparse_resize_kernel(const float* __restrict__ in, float* __restrict__ out, const int* __restrict__ rowptr,
                                            const int* __restrict__ idx, const float* __restrict__ w, int C, int n_other, int n_in, int n_out,
                                            int along_x, const float* __restrict__ mean, const float* __restrict__ inv_std, int pre_scale) {
  const size_t total = size_t(C) * n_other * n_out;
  const int b = blockIdx.y;
  in += size_t(b) * C * n_other * n_in;
  out += size_t(b) * total;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    int c, o, r;          // channel, output index along the resized axis, index along the other axis
    if (along_x) { o = int(i % n_out); r = int((i / n_out) % n_other); c = int(i / (size_t(n_out) * n_other)); }
    else { r = int(i % n_other); o = int((i / n_other) % n_out); c = int(i / (size_t(n_out) * n_other)); }
    float acc = 0.f;
    for (int k = rowptr[o]; k < rowptr[o + 1]; ++k) {
      const size_t src = along_x ? (size_t(c) * n_other + r) * n_in + idx[k] : (size_t(c) * n_in + idx[k]) * n_other + r;
      acc = fmaf(w[k], in[src], acc);
    }
    if (mean) acc = (acc - mean[c]) * inv_std[c];
    else if (pre_scale) acc *= inv_std[c];
    out[i] = acc;
  }
}

// ------------------------------------------------------------------------------------------------ patch embedding glue
// img [B][3][R][R] fp32 -> patches 16-bit [B*G*G][3*P*P] (k = c*P*P + py*P + px, the layout of conv1.weight.reshape(W, -1))
static __global__ void patchify_kernel(const float* __restrict__ img, op_t* __restrict__ out, int B, int R, int P) {
  const int G = R / P, K = 3 * P * P;
  const size_t total = size_t(B) * G * G * K;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int k = int(i % K); const size_t t = i / K;
    const int gx = int(t % G), gy = int((t / G) % G), b = int(t / (size_t(G) * G));
    const int px = k % P, py = (k / P) % P, c = k / (P * P);
    out[i] = to_op(img[((size_t(b) * 3 + c) * R + gy * P + py) * R + gx * P + px]);
  }
}
// gradient wrt patches fp32 [B*G*G][3*P*P] -> image gradient [B][3][R][R]
static __global__ void unpatchify_kernel(const float* __restrict__ gp, float* __restrict__ gimg, int B, int R, int P, float post) {
  const int G = R / P, K = 3 * P * P;
  const size_t total = size_t(B) * 3 * R * R;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int x = int(i % R), y = int((i / R) % R), c = int((i / (size_t(R) * R)) % 3), b = int(i / (size_t(3) * R * R));
    const int gx = x / P, px = x % P, gy = y / P, py = y % P;
    gimg[i] = post * gp[((size_t(b) * G + gy) * G + gx) * K + (c * P + py) * P + px];
  }
}
// token + position embedding: x[b][t][:] = tok[ids[b][t]][:] + pos[t][:]   (CLIP text tower)
static __global__ void embed_tokens_kernel(const int* __restrict__ ids, const float* __restrict__ tok, const float* __restrict__ pos,
                                           float* __restrict__ x, int T, int W, int vocab) {
  const int r = blockIdx.x, t = r % T;
  int id = ids[r];
  id = min(max(id, 0), vocab - 1);
  for (int c = threadIdx.x; c < W; c += blockDim.x) x[size_t(r) * W + c] = tok[size_t(id) * W + c] + pos[size_t(t) * W + c];
}
// class-token rows: x[b][0][:] = class_embedding + positional_embedding[0]
static __global__ void class_token_kernel(float* __restrict__ x, const float* __restrict__ cls, const float* __restrict__ pos, int T, int W) {
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < W; c += blockDim.x) x[size_t(b) * T * W + c] = cls[c] + pos[c];
}
// copy rows [B][T][W] -> patch rows [B][T-1][W] (drop the class token), fp32 -> 16-bit and/or fp32
static __global__ void drop_class_rows_kernel(const float* __restrict__ x, float* __restrict__ y32, op_t* __restrict__ y16, int B, int T, int W) {
  const size_t total = size_t(B) * (T - 1) * W;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int c = int(i % W); const size_t r = i / W;
    const int t = int(r % (T - 1)), b = int(r / (T - 1));
    const float v = x[(size_t(b) * T + t + 1) * W + c];
    if (y32) y32[i] = v;
    if (y16) y16[i] = to_op(v);
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm forward / backward
// one warp per row; W % 32 == 0, W <= 4096.  stats[row] = (mean, rstd).
static __global__ void ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                     float* __restrict__ y32, op_t* __restrict__ y16, float2* __restrict__ stats, int rows, int W, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + size_t(row) * W;
  float s = 0.f;
  for (int c = lane; c < W; c += 32) s += xr[c];
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / W;
  float q = 0.f;
  for (int c = lane; c < W; c += 32) { const float d = xr[c] - mean; q = fmaf(d, d, q); }
#pragma unroll
  for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / W + eps);
  if (lane == 0 && stats) stats[row] = make_float2(mean, rstd);
  for (int c = lane; c < W; c += 32) {
    const float v = (xr[c] - mean) * rstd * gamma[c] + beta[c];
    if (y32) y32[size_t(row) * W + c] = v;
    if (y16) y16[size_t(row) * W + c] = to_op(v);
  }
}
// dx = rstd * (g*gamma - mean(g*gamma) - xh * mean(g*gamma*xh)) (+ add)
static __global__ void ln_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x, const float2* __restrict__ stats,
                                     const float* __restrict__ gamma, const float* __restrict__ add, float* __restrict__ dx32,
                                     op_t* __restrict__ dx16, int rows, int W) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float2 st = stats[row];
  const float* xr = x + size_t(row) * W; const float* gr = g + size_t(row) * W;
  float a = 0.f, b = 0.f;
  for (int c = lane; c < W; c += 32) { const float d = gr[c] * gamma[c], xh = (xr[c] - st.x) * st.y; a += d; b = fmaf(d, xh, b); }
#pragma unroll
  for (int o = 16; o; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  a /= W; b /= W;
  for (int c = lane; c < W; c += 32) {
    const float d = gr[c] * gamma[c], xh = (xr[c] - st.x) * st.y;
    float v = st.y * (d - a - xh * b);
    if (add) v += add[size_t(row) * W + c];
    if (dx32) dx32[size_t(row) * W + c] = v;
    if (dx16) dx16[size_t(row) * W + c] = to_op(v);
  }
}

// ------------------------------------------------------------------------------------------------ QuickGELU (x * sigmoid(1.702 x))
static __global__ void quickgelu_fwd_kernel(const float* __restrict__ h, op_t* __restrict__ a, size_t n) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const float v = h[i];
    a[i] = to_op(v / (1.0f + __expf(-1.702f * v)));
  }
}
static __global__ void quickgelu_bwd_kernel(const float* __restrict__ da, const float* __restrict__ h, op_t* __restrict__ dh, size_t n) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const float v = h[i], s = 1.0f / (1.0f + __expf(-1.702f * v));
    dh[i] = to_op(da[i] * (s + 1.702f * v * s * (1.0f - s)));
  }
}
static __global__ void cast_rows_kernel(const float* __restrict__ x, op_t* __restrict__ y, size_t n) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) y[i] = to_op(x[i]);
}

// ------------------------------------------------------------------------------------------------ small multi-head attention (N <= 256, d = 64)
// qkv 16-bit [B][N][3*H*64] (q | k | v, head h at columns h*64 of each third).  grid (ceil(N/32), H, B), 256 threads: warp <-> 4 query rows.
// Forward keeps P fp32 [B][H][N][N] for the backward.
constexpr int kAttD = 64, kAttMaxN = 256, kAttLd = 66;     // padded smem row (halves): conflict-free when lanes read different rows

HEDIT_DEVICE void att_load_rows(op_t* dst, const op_t* src, int N, int ld_src) {      // [N][64] -> smem [N][66]
  for (int i = threadIdx.x; i < N * (kAttD / 2); i += blockDim.x) {
    const int r = i / (kAttD / 2), c2 = i % (kAttD / 2);
    *reinterpret_cast<uint32_t*>(dst + r * kAttLd + 2 * c2) = *reinterpret_cast<const uint32_t*>(src + size_t(r) * ld_src + 2 * c2);
  }
}

static __global__ void att_small_fwd_kernel(const op_t* __restrict__ qkv, float* __restrict__ P, op_t* __restrict__ out, int N, int H, float scale,
                                            int causal = 0) {
  extern __shared__ __align__(16) uint8_t att_sm[];
  op_t* sK = reinterpret_cast<op_t*>(att_sm);
  op_t* sV = sK + kAttMaxN * kAttLd;
  float* sP = reinterpret_cast<float*>(sV + kAttMaxN * kAttLd);      // [8 warps][256]
  float* sQ = sP + 8 * kAttMaxN;                                      // [8 warps][64]
  const int h = blockIdx.y, b = blockIdx.z, ld = 3 * H * kAttD, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const op_t* base = qkv + size_t(b) * N * ld + h * kAttD;
  att_load_rows(sK, base + H * kAttD, N, ld);
  att_load_rows(sV, base + 2 * H * kAttD, N, ld);
  __syncthreads();
  for (int rr = 0; rr < 4; ++rr) {
    const int row = blockIdx.x * 32 + warp * 4 + rr;
    if (row >= N) break;                                              // warp-uniform
    sQ[warp * 64 + lane] = op_to_float(base[size_t(row) * ld + lane]);
    sQ[warp * 64 + 32 + lane] = op_to_float(base[size_t(row) * ld + 32 + lane]);
    __syncwarp();
    float sc[kAttMaxN / 32], mx = -INFINITY;
#pragma unroll
    for (int u = 0; u < kAttMaxN / 32; ++u) {
      const int j = lane + 32 * u;
      float a = -INFINITY;
      if (j < N && !(causal && j > row)) {
        a = 0.f;
        for (int c = 0; c < kAttD; c += 2) {
          const float2 kk = op2_to_float2(*reinterpret_cast<const uint32_t*>(sK + j * kAttLd + c));
          a = fmaf(sQ[warp * 64 + c], kk.x, a); a = fmaf(sQ[warp * 64 + c + 1], kk.y, a);
        }
        a *= scale;
      }
      sc[u] = a; mx = fmaxf(mx, a);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float l = 0.f;
#pragma unroll
    for (int u = 0; u < kAttMaxN / 32; ++u) { sc[u] = (sc[u] > -INFINITY) ? __expf(sc[u] - mx) : 0.f; l += sc[u]; }
#pragma unroll
    for (int o = 16; o; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    const float inv = 1.f / l;
    float* prow = P + ((size_t(b) * H + h) * N + row) * N;
#pragma unroll
    for (int u = 0; u < kAttMaxN / 32; ++u) {
      const int j = lane + 32 * u;
      if (j < N) { const float p = sc[u] * inv; if (P) prow[j] = p; sP[warp * kAttMaxN + j] = p; }
    }
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < N; ++j) {
      const float p = sP[warp * kAttMaxN + j];
      o0 = fmaf(p, op_to_float(sV[j * kAttLd + lane]), o0); o1 = fmaf(p, op_to_float(sV[j * kAttLd + 32 + lane]), o1);
    }
    op_t* orow = out + (size_t(b) * N + row) * (H * kAttD) + h * kAttD;
    orow[lane] = to_op(o0); orow[32 + lane] = to_op(o1);
    __syncwarp();
  }
}

// dP = dO V^T ; dS = scale * P * (dP - sum_j P dP) (kept fp32 [B][H][N][N]) ; dQ = dS K -> dqkv q-slice.  Same grid as the forward.
static __global__ void att_small_bwd_dq_kernel(const op_t* __restrict__ qkv, const float* __restrict__ P, const op_t* __restrict__ dO,
                                               float* __restrict__ dS, op_t* __restrict__ dqkv, int N, int H, float scale) {
  extern __shared__ __align__(16) uint8_t att_sm[];
  op_t* sK = reinterpret_cast<op_t*>(att_sm);
  op_t* sV = sK + kAttMaxN * kAttLd;
  float* sP = reinterpret_cast<float*>(sV + kAttMaxN * kAttLd);
  float* sQ = sP + 8 * kAttMaxN;                                      // holds the dO row here
  const int h = blockIdx.y, b = blockIdx.z, ld = 3 * H * kAttD, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const op_t* base = qkv + size_t(b) * N * ld + h * kAttD;
  att_load_rows(sK, base + H * kAttD, N, ld);
  att_load_rows(sV, base + 2 * H * kAttD, N, ld);
  __syncthreads();
  for (int rr = 0; rr < 4; ++rr) {
    const int row = blockIdx.x * 32 + warp * 4 + rr;
    if (row >= N) break;
    const op_t* dor = dO + (size_t(b) * N + row) * (H * kAttD) + h * kAttD;
    sQ[warp * 64 + lane] = op_to_float(dor[lane]); sQ[warp * 64 + 32 + lane] = op_to_float(dor[32 + lane]);
    __syncwarp();
    const float* prow = P + ((size_t(b) * H + h) * N + row) * N;
    float dp[kAttMaxN / 32], pv[kAttMaxN / 32], t = 0.f;
#pragma unroll
    for (int u = 0; u < kAttMaxN / 32; ++u) {
      const int j = lane + 32 * u;
      float a = 0.f, p = 0.f;
      if (j < N) {
        for (int c = 0; c < kAttD; c += 2) {
          const float2 vv = op2_to_float2(*reinterpret_cast<const uint32_t*>(sV + j * kAttLd + c));
          a = fmaf(sQ[warp * 64 + c], vv.x, a); a = fmaf(sQ[warp * 64 + c + 1], vv.y, a);
        }
        p = prow[j];
      }
      dp[u] = a; pv[u] = p; t = fmaf(p, a, t);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    float* dsrow = dS + ((size_t(b) * H + h) * N + row) * N;
#pragma unroll
    for (int u = 0; u < kAttMaxN / 32; ++u) {
      const int j = lane + 32 * u;
      if (j < N) { const float d = scale * pv[u] * (dp[u] - t); dsrow[j] = d; sP[warp * kAttMaxN + j] = d; }
    }
    __syncwarp();
    float q0 = 0.f, q1 = 0.f;
    for (int j = 0; j < N; ++j) {
      const float d = sP[warp * kAttMaxN + j];
      q0 = fmaf(d, op_to_float(sK[j * kAttLd + lane]), q0); q1 = fmaf(d, op_to_float(sK[j * kAttLd + 32 + lane]), q1);
    }
    op_t* qrow = dqkv + (size_t(b) * N + row) * ld + h * kAttD;
    qrow[lane] = to_op(q0); qrow[32 + lane] = to_op(q1);
    __syncwarp();
  }
}

// dV[j] = sum_i P[i][j] dO[i] ; dK[j] = sum_i dS[i][j] Q[i].  grid (ceil(N/32), H, B): warp <-> 4 key rows, lanes <-> 2 columns each.
static __global__ void att_small_bwd_dkv_kernel(const op_t* __restrict__ qkv, const float* __restrict__ P, const float* __restrict__ dS,
                                                const op_t* __restrict__ dO, op_t* __restrict__ dqkv, int N, int H) {
  extern __shared__ __align__(16) uint8_t att_sm[];
  op_t* sQ = reinterpret_cast<op_t*>(att_sm);
  op_t* sO = sQ + kAttMaxN * kAttLd;
  const int h = blockIdx.y, b = blockIdx.z, ld = 3 * H * kAttD, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  att_load_rows(sQ, qkv + size_t(b) * N * ld + h * kAttD, N, ld);
  att_load_rows(sO, dO + size_t(b) * N * (H * kAttD) + h * kAttD, N, H * kAttD);
  __syncthreads();
  const float* Pb = P + (size_t(b) * H + h) * N * N;
  const float* Sb = dS + (size_t(b) * H + h) * N * N;
  for (int rr = 0; rr < 4; ++rr) {
    const int j = blockIdx.x * 32 + warp * 4 + rr;
    if (j >= N) break;
    float v0 = 0.f, v1 = 0.f, k0 = 0.f, k1 = 0.f;
    for (int i = 0; i < N; ++i) {
      const float p = Pb[size_t(i) * N + j], d = Sb[size_t(i) * N + j];
      v0 = fmaf(p, op_to_float(sO[i * kAttLd + lane]), v0); v1 = fmaf(p, op_to_float(sO[i * kAttLd + 32 + lane]), v1);
      k0 = fmaf(d, op_to_float(sQ[i * kAttLd + lane]), k0); k1 = fmaf(d, op_to_float(sQ[i * kAttLd + 32 + lane]), k1);
    }
    op_t* row = dqkv + (size_t(b) * N + j) * ld + h * kAttD;
    row[H * kAttD + lane] = to_op(k0); row[H * kAttD + 32 + lane] = to_op(k1);
    row[2 * H * kAttD + lane] = to_op(v0); row[2 * H * kAttD + 32 + lane] = to_op(v1);
  }
}

// ------------------------------------------------------------------------------------------------ Gram residual, its norm and gradient (fp32)
// G[b] = F[b]^T F[b] - Gref, F [B][T][W] (T = patch tokens).  grid (W/16, W/16, B), block (16,16)
static __global__ void gram_residual_kernel(const float* __restrict__ F, const float* __restrict__ Gref, float* __restrict__ G, int T, int W) {
  __shared__ float sa[16][17], sb[16][17];
  const int b = blockIdx.z, i0 = blockIdx.y * 16, j0 = blockIdx.x * 16;
  const float* Fb = F + size_t(b) * T * W;
  float acc = 0.f;
  for (int t0 = 0; t0 < T; t0 += 16) {
    const int t = t0 + threadIdx.y;
    sa[threadIdx.y][threadIdx.x] = (t < T) ? Fb[size_t(t) * W + i0 + threadIdx.x] : 0.f;
    sb[threadIdx.y][threadIdx.x] = (t < T) ? Fb[size_t(t) * W + j0 + threadIdx.x] : 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc = fmaf(sa[k][threadIdx.y], sb[k][threadIdx.x], acc);
    __syncthreads();
  }
  const size_t o = size_t(i0 + threadIdx.y) * W + j0 + threadIdx.x;
  G[size_t(b) * W * W + o] = acc - (Gref ? Gref[o] : 0.f);
}
// loss[b] = ||G[b]||_F ; one CTA per image, fixed-order tree
static __global__ void frob_norm_kernel(const float* __restrict__ G, float* __restrict__ loss, size_t n) {
  __shared__ double red[256];
  const float* g = G + size_t(blockIdx.x) * n;
  double s = 0.0;
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) s += double(g[i]) * g[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) loss[blockIdx.x] = float(sqrt(red[0]));
}
// dF = (2 / loss) * F G   (G symmetric): [T][W] x [W][W].  Written into rows 1.. of a [B][T+1][W] buffer (row 0 = class token: zero)
// times `gscale` (the caller's power-of-two-free scale that keeps 16-bit gradient operands in range).  grid (W/16, ceil(T/16), B)
static __global__ void gram_grad_kernel(const float* __restrict__ F, const float* __restrict__ G, const float* __restrict__ loss,
                                        float* __restrict__ dX, int T, int W, float gscale) {
  __shared__ float sa[16][17], sb[16][17];
  const int b = blockIdx.z, t0 = blockIdx.y * 16, j0 = blockIdx.x * 16;
  const float* Fb = F + size_t(b) * T * W; const float* Gb = G + size_t(b) * W * W;
  float acc = 0.f;
  for (int k0 = 0; k0 < W; k0 += 16) {
    const int t = t0 + threadIdx.y;
    sa[threadIdx.y][threadIdx.x] = (t < T) ? Fb[size_t(t) * W + k0 + threadIdx.x] : 0.f;
    sb[threadIdx.y][threadIdx.x] = Gb[size_t(k0 + threadIdx.y) * W + j0 + threadIdx.x];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc = fmaf(sa[threadIdx.y][k], sb[k][threadIdx.x], acc);
    __syncthreads();
  }
  const int t = t0 + threadIdx.y;
  if (t < T) dX[(size_t(b) * (T + 1) + t + 1) * W + j0 + threadIdx.x] = acc * (2.f * gscale / fmaxf(loss[b], 1e-30f));
  if (t0 == 0 && threadIdx.y == 0) dX[size_t(b) * (T + 1) * W + j0 + threadIdx.x] = 0.f;
}

}  // namespace hedit
