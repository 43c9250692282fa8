// Persistent, warp-specialised tcgen05 GEMM / implicit-GEMM-conv kernel for sm_100a.
//
//   D[M,N] = A[M,K] * W[N,K]^T  (+ bias[N]) (+ rowvec[row / rows_per_group][N]) (+ residual[M,N])
//
// A and W are bf16, K-major, fetched by TMA (128-byte swizzle) into a multi-stage shared-memory ring; the fp32
// accumulator lives in TMEM (two buffers, so the epilogue of tile i overlaps the main loop of tile i+1).
// Warp roles: 0 = TMA producer, 1 = MMA issuer (single elected thread) + TMEM owner, 2..5 = epilogue.
//
// A-operand addressing modes (all NHWC bf16 activations, no im2col buffer is ever materialised):
//   A_LINEAR   : 2-D map [M][K]                                   (linear layers, 1x1 convs)
//   A_CONV3X3  : 4-D map (C, W, H, S); tap (ky,kx) = box shifted by (kx-1, ky-1); TMA zero-fills the padding
//   A_CONV3X3S2: 5-D map (2C, W/2, 2, H/2, S) over the same memory = space-to-depth view for stride 2, pad 1
#pragma once
#include "ptx.cuh"

namespace hedit {

enum { A_LINEAR = 0, A_CONV3X3 = 1, A_CONV3X3S2 = 2 };

struct GemmEpilogue {
  const float* bias;        // [N] or null
  const float* rowvec;      // [groups][ldrv] or null; group = row / rows_per_group (time-embedding add)
  const float* residual;    // [M][ldr] fp32 or null
  float* out_f32;           // [M][ldo] or null
  op_t* out_bf16;  // [M][ldob] or null
  int rows_per_group, ldrv, ldr, ldo, ldob;
  int geglu;                // 1: every 32-col chunk = 16 value | 16 gate -> bf16 out has N/2 columns
};

struct GemmParams {
  CUtensorMap tmA, tmB;
  int M, N, num_kb;
  int a_mode, conv_W, conv_H, conv_cin, cin_blocks;
  GemmEpilogue ep;
};

template <int BN>
struct GemmCfg {
  static constexpr int BM = 128, BK = 64;
  static constexpr int STAGES = (BN > 160) ? 4 : 6;
  static constexpr uint32_t A_BYTES = BM * BK * 2;
  static constexpr uint32_t B_BYTES = BN * BK * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 256 /*barriers*/;
  static constexpr int THREADS = 192;
};

template <int BN>
__global__ void __launch_bounds__(192, 1) gemm_bf16_tcgen05_kernel(const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();     // 128B-swizzle tiles need 1024-byte alignment
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;       // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (p.M + 127) >> 7;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) << 7;
        const int n0 = (tile % n_tiles) * BN;
        int s0 = 0, y0 = 0;
        if (p.a_mode != A_LINEAR) {
          const int hw = p.conv_H * p.conv_W;
          s0 = m0 / hw;
          y0 = (m0 % hw) / p.conv_W;
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          if (p.a_mode == A_LINEAR) {
            tma_load_2d(sa, &p.tmA, &full_bar[stage], kb * 64, m0);
          } else {
            const int tap = kb / p.cin_blocks, cb = kb - tap * p.cin_blocks;
            const int ky = tap / 3, kx = tap - ky * 3;
            if (p.a_mode == A_CONV3X3) {
              tma_load_4d(sa, &p.tmA, &full_bar[stage], cb * 64, kx - 1, y0 + ky - 1, s0);
            } else {
              const int px = (kx == 1) ? 0 : 1, dx = (kx == 0) ? -1 : 0;
              const int py = (ky == 1) ? 0 : 1, dy = (ky == 0) ? -1 : 0;
              tma_load_5d(sa, &p.tmA, &full_bar[stage], px * p.conv_cin + cb * 64, dx, py, y0 + dy, s0);
            }
          }
          tma_load_2d(sb, &p.tmB, &full_bar[stage], kb * 64, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
      int stage = 0; uint32_t phase = 0; int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t da = umma_desc_kmajor_sw128(sa);
          const uint64_t db = umma_desc_kmajor_sw128(sa + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)   // 4 x (K=16) per 64-wide stage; +32 B per step inside the swizzle atom
            umma_f16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (4 warps, one TMEM lane quarter each)
    const int quarter = warp & 3;
    const GemmEpilogue& e = p.ep;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m0 = (tile / n_tiles) << 7;
      const int n0 = (tile % n_tiles) * BN;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row = m0 + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t t_row = tmem_base + acc * 256 + (uint32_t(quarter * 32) << 16);
      const float* rv = (e.rowvec && row_ok) ? e.rowvec + size_t(row / e.rows_per_group) * e.ldrv : nullptr;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        const int col = n0 + c0;
        if (col >= p.N) break;                      // warp-uniform
        uint32_t raw[32];
        tmem_ld32(t_row + c0, raw);
        tmem_ld_wait();
        if (!row_ok) continue;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
        const bool full = (col + 32 <= p.N);
        if (full) {
          if (e.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = *reinterpret_cast<const float4*>(e.bias + col + j);
              v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
          }
          if (rv) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = *reinterpret_cast<const float4*>(rv + col + j);
              v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
          }
          if (e.geglu) {
            // chunk = 16 value columns followed by their 16 gate columns -> 16 outputs
            op_t* o = e.out_bf16 + size_t(row) * e.ldob + (col >> 1);
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 16; j += 2)
              pk[j >> 1] = pack_op2(v[j] * gelu_erf_f(v[16 + j]), v[j + 1] * gelu_erf_f(v[17 + j]));
            *reinterpret_cast<uint4*>(o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(o + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            continue;
          }
          if (e.residual) {
            const float* r = e.residual + size_t(row) * e.ldr + col;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = *reinterpret_cast<const float4*>(r + j);
              v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
          }
          if (e.out_f32) {
            float* o = e.out_f32 + size_t(row) * e.ldo + col;
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
          if (e.out_bf16) {
            op_t* o = e.out_bf16 + size_t(row) * e.ldob + col;
#pragma unroll
            for (int j = 0; j < 32; j += 8)
              *reinterpret_cast<uint4*>(o + j) = make_uint4(pack_op2(v[j], v[j + 1]), pack_op2(v[j + 2], v[j + 3]),
                                                            pack_op2(v[j + 4], v[j + 5]), pack_op2(v[j + 6], v[j + 7]));
          }
        } else {
          // ragged tail (N not a multiple of 32): scalar, no geglu
          for (int j = 0; j < 32 && col + j < p.N; ++j) {
            float x = v[j];
            if (e.bias) x += e.bias[col + j];
            if (rv) x += rv[col + j];
            if (e.residual) x += e.residual[size_t(row) * e.ldr + col + j];
            if (e.out_f32) e.out_f32[size_t(row) * e.ldo + col + j] = x;
            if (e.out_bf16) e.out_bf16[size_t(row) * e.ldob + col + j] = to_op(x);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace hedit
