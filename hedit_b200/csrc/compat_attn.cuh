// Compatibility attention: probabilities MATERIALISED in fp32 so that an arbitrary user controller can edit them in place between the
// softmax and P.V, exactly where the reference's processor calls it (p2p/ptp_utils.py:96-107: get_attention_scores -> controller(probs,
// is_cross, place, save_attn) -> bmm(probs, value)).  Plain CUDA-core kernels: this path exists for API exactness (any controller
// object), not speed; the stock controllers are compiled into the fused tcgen05 attention kernels instead (attention.cuh).
// Layout of probs: [(sample * H + head)][Nq][Nkv], the reference's head_to_batch_dim order.
#pragma once
#include <cuda_runtime.h>

#include "ptx.cuh"

namespace hedit {

struct CompatAttnParams {
  const op_t* q; int ldq; size_t q_sample;     // q rows: q + s * q_sample + i * ldq + h * d
  const op_t* k; const op_t* v; int ldkv; size_t kv_sample;
  const int* kv_idx;                           // per sample: which K/V block (text context) to read, or null = own sample
  float* probs;
  op_t* out; int ldo;                          // [S * Nq][ldo], head h at columns h * d
  int H, d, Nq, Nkv;
  float scale;
};

// scores[b][i][j] = scale * <q_i, k_j>; 64 x 64 tile per CTA, 4 x 4 per thread, head dim in chunks of 8
static __global__ void __launch_bounds__(256) compat_scores_kernel(const CompatAttnParams p) {
  __shared__ float qs[8][64 + 4], ks[8][64 + 4];
  const int b = blockIdx.z, s = b / p.H, h = b % p.H;
  const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  const int kvs = p.kv_idx ? p.kv_idx[s] : s;
  const op_t* qb = p.q + size_t(s) * p.q_sample + size_t(h) * p.d;
  const op_t* kb = p.k + size_t(kvs) * p.kv_sample + size_t(h) * p.d;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
  for (int d0 = 0; d0 < p.d; d0 += 8) {
    // 64 rows x 8 dims of q and of k: 512 elements each, 2 per thread
    for (int e = threadIdx.x; e < 512; e += 256) {
      const int r = e >> 3, c = e & 7;
      qs[c][r] = (i0 + r < p.Nq && d0 + c < p.d) ? op_to_float(qb[size_t(i0 + r) * p.ldq + d0 + c]) : 0.f;
      ks[c][r] = (j0 + r < p.Nkv && d0 + c < p.d) ? op_to_float(kb[size_t(j0 + r) * p.ldkv + d0 + c]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float qa[4], ka[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) { qa[a] = qs[c][ty * 4 + a]; ka[a] = ks[c][tx * 4 + a]; }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[a][e] = fmaf(qa[a], ka[e], acc[a][e]);
    }
    __syncthreads();
  }
  float* pb = p.probs + size_t(b) * p.Nq * p.Nkv;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = i0 + ty * 4 + a;
    if (i >= p.Nq) continue;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = j0 + tx * 4 + e;
      if (j < p.Nkv) pb[size_t(i) * p.Nkv + j] = acc[a][e] * p.scale;
    }
  }
}

// in-place row softmax; one warp per row, 8 rows per CTA
static __global__ void __launch_bounds__(256) compat_softmax_kernel(float* __restrict__ probs, size_t rows, int N) {
  const size_t row = size_t(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float* r = probs + row * N;
  const int lane = threadIdx.x & 31;
  float mx = -INFINITY;
  for (int j = lane; j < N; j += 32) mx = fmaxf(mx, r[j]);
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float l = 0.f;
  for (int j = lane; j < N; j += 32) { const float e = expf(r[j] - mx); r[j] = e; l += e; }
#pragma unroll
  for (int o = 16; o; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  const float inv = 1.f / l;
  for (int j = lane; j < N; j += 32) r[j] *= inv;
}

// out[s][i][h*d + c] = sum_j probs[b][i][j] * v[j][h*d + c]; 32 query rows per CTA, keys in chunks of 32, thread = (row, column lane of 8)
template <int DMAX>
static __global__ void __launch_bounds__(256) compat_pv_kernel(const CompatAttnParams p) {
  __shared__ float ps[32][32 + 1];
  __shared__ float vs[32][DMAX];
  const int b = blockIdx.y, s = b / p.H, h = b % p.H;
  const int i0 = blockIdx.x * 32;
  const int kvs = p.kv_idx ? p.kv_idx[s] : s;
  const op_t* vb = p.v + size_t(kvs) * p.kv_sample + size_t(h) * p.d;
  const float* pb = p.probs + size_t(b) * p.Nq * p.Nkv;
  const int r = threadIdx.x >> 3, cl = threadIdx.x & 7;
  constexpr int NC = DMAX / 8;
  float acc[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) acc[c] = 0.f;
  for (int j0 = 0; j0 < p.Nkv; j0 += 32) {
    for (int e = threadIdx.x; e < 32 * 32; e += 256) {
      const int rr = e >> 5, jj = e & 31;
      ps[rr][jj] = (i0 + rr < p.Nq && j0 + jj < p.Nkv) ? pb[size_t(i0 + rr) * p.Nkv + j0 + jj] : 0.f;
    }
    for (int e = threadIdx.x; e < 32 * p.d; e += 256) {
      const int jj = e / p.d, c = e % p.d;
      vs[jj][c] = (j0 + jj < p.Nkv) ? op_to_float(vb[size_t(j0 + jj) * p.ldkv + c]) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int jj = 0; jj < 32; ++jj) {
      const float pv = ps[r][jj];
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (cl + 8 * c < p.d) acc[c] = fmaf(pv, vs[jj][cl + 8 * c], acc[c]);
    }
    __syncthreads();
  }
  if (i0 + r < p.Nq) {
    op_t* o = p.out + (size_t(s) * p.Nq + i0 + r) * p.ldo + size_t(h) * p.d;
#pragma unroll
    for (int c = 0; c < NC; ++c)
      if (cl + 8 * c < p.d) o[cl + 8 * c] = to_op(acc[c]);
  }
}

// ---- editor protocol (masactrl/masactrl_utils.py:40-89): the hook receives q, k, v as fp32 [(S*H)][N][d], the scaled scores and the
// probabilities, and returns the layer output [S][N][H*d]
// 16-bit [samples][N][ld], head h at columns h*d  ->  fp32 [(s*H + h)][N][d]; grid (ceil(N*d / 256), S*H)
static __global__ void compat_split_heads_kernel(const op_t* __restrict__ src, int ld, size_t sample_stride, const int* __restrict__ idx,
                                                 float* __restrict__ dst, int H, int d, int N) {
  const int b = blockIdx.y, s = b / H, h = b % H;
  const int ss = idx ? idx[s] : s;
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= N * d) return;
  const int n = e / d, c = e % d;
  dst[(size_t(b) * N + n) * d + c] = op_to_float(src[size_t(ss) * sample_stride + size_t(n) * ld + h * d + c]);
}
// fp32 [rows][C] -> 16-bit [rows][ldo]
static __global__ void compat_store_out_kernel(const float* __restrict__ src, op_t* __restrict__ dst, int ldo, int C, size_t rows) {
  const size_t e = size_t(blockIdx.x) * 256 + threadIdx.x;
  if (e >= rows * C) return;
  const size_t r = e / C; const int c = int(e % C);
  dst[r * ldo + c] = to_op(src[e]);
}

}  // namespace hedit
