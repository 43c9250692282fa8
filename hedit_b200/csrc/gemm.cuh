// Persistent, warp-specialised tcgen05 GEMM / implicit-GEMM-conv kernel for sm_100a.
//
//   D[M,N] = A[M,K] * W[N,K]^T  (+ bias[N]) (+ rowvec[row / rows_per_group][N]) (+ residual[M,N])
//
// A and W are bf16, K-major, fetched by TMA (128-byte swizzle) into a multi-stage shared-memory ring; the fp32
// accumulator lives in TMEM (two buffers, so the epilogue of tile i overlaps the main loop of tile i+1).
// Warp roles: 0 = TMA producer, 1 = MMA issuer (single elected thread) + TMEM owner, 2..9 = epilogue.
//
// A-operand addressing modes (all NHWC bf16 activations, no im2col buffer is ever materialised):
//   A_LINEAR   : 2-D map [M][K]                                   (linear layers, 1x1 convs)
//   A_CONV3X3  : 4-D map (C, W, H, S); tap (ky,kx) = box shifted by (kx-1, ky-1); TMA zero-fills the padding
//   A_CONV3X3S2: 5-D map (2C, W/2, 2, H/2, S) over the same memory = space-to-depth view for stride 2, pad 1
//   A_CONV2X2  : 4-D map like A_CONV3X3, 4 taps (a,b) = box shifted by (b + conv_ox, a + conv_oy): one output phase of
//                "nearest 2x upsample -> 3x3 conv" evaluated on the COARSE input with pre-summed weights (2.25x fewer flops);
//                the epilogue scatters row (s,y,x) to the fine pixel (2y+py, 2x+px) (GemmEpilogue::up_*)
#pragma once
#include <type_traits>

#include "ptx.cuh"

namespace hedit {

enum { A_LINEAR = 0, A_CONV3X3 = 1, A_CONV3X3S2 = 2, A_CONV2X2 = 3 };

struct GemmEpilogue {
  const float* bias;        // [N] or null
  const float* rowvec;      // [groups][ldrv] or null; group = row / rows_per_group (time-embedding add)
  const float* residual;    // [M][ldr] fp32 or null
  float* out_f32;           // [M][ldo] or null
  op_t* out_bf16;  // [M][ldob] or null
  int rows_per_group, ldrv, ldr, ldo, ldob;
  int diag_skip;            // diagnostics only (op_bench): the write-back is skipped, accumulators are just released
  int geglu;                // 1: every 32-col chunk = 16 value | 16 gate -> bf16 out has N/2 columns
  int nchw_hw;              // >0: write out_f32 as [row / hw][N][row % hw] (NCHW latent layout; small-N generic path only)
  int up_W, up_H, up_py, up_px;   // up_W > 0: GEMM row (s,y,x) of a coarse up_H x up_W grid is written to fine row (s, 2y+up_py, 2x+up_px)
                            // (out_f32 + bias only); its colstats block goes to [s][phase][block] so that the 4 phases tile the sample
  float2* colstats;         // [ceil(M/32)][N] or null: (sum, sum of squares) of the fp32 output over each block of 32 rows, per
                            // column -- the GroupNorm statistics of the NEXT layer, produced while the tile is still in registers
};

struct GemmParams {
  CUtensorMap tmA, tmB;
  CUtensorMap tmD;          // 16-bit output [M][ldob] as boxes of 64 columns x 32 rows (128B swizzle) when use_tmd (bulk tensor stores)
  int use_tmd;
  // split-K (launches with far fewer tiles than SMs and a long K loop: the deep 8x8 / 16x16 levels at 1-5 samples).  splits > 1: work
  // item = (tile, split); split s accumulates K blocks [s * kb_per_split, ...) and writes its fp32 partial tile to
  // ep.out_f32 + s * split_stride (ep then carries no bias / residual / statistics); splitk_reduce_kernel adds the partials in a fixed
  // order and applies the real epilogue
  int splits, kb_per_split;
  size_t split_stride;
  int M, N, num_kb;
  int a_mode, conv_W, conv_H, conv_cin, cin_blocks;
  int conv_ox, conv_oy;     // A_CONV2X2 only: first tap offset per axis (-1 for output phase 0, 0 for phase 1)
  int b_full_box;           // 1: tmB's box covers all BN rows (single-CTA launches: one W load per K block instead of two)
  int conv_pad01;           // A_CONV3X3S2 only: 0 = padding 1 on every side (SD downsampler); 1 = padding (0,1,0,1) (DDPM downsampler)
  GemmEpilogue ep;
};

// Coalesced write-back of one 32x32 chunk from the per-warp staging tile: lane = (row-in-group-of-4, quad); 8 iterations
// cover the 32 rows.  All feature switches are compile-time so the loop body is ~12 instructions per float4.
// Column statistics of one 32x32 chunk: every lane holds partial sums of its 4 columns over 8 rows; the 4 lanes that share a
// column quad (lane bits 3,4) are combined in a fixed order and lane rr == 0 writes (sum, sumsq) x 4 columns.
HEDIT_DEVICE void epi_store_colstats(float4 su, float4 sq, int rr, float2* dst) {
#pragma unroll
  for (int o = 8; o <= 16; o <<= 1) {
    su.x += __shfl_xor_sync(0xffffffffu, su.x, o); su.y += __shfl_xor_sync(0xffffffffu, su.y, o);
    su.z += __shfl_xor_sync(0xffffffffu, su.z, o); su.w += __shfl_xor_sync(0xffffffffu, su.w, o);
    sq.x += __shfl_xor_sync(0xffffffffu, sq.x, o); sq.y += __shfl_xor_sync(0xffffffffu, sq.y, o);
    sq.z += __shfl_xor_sync(0xffffffffu, sq.z, o); sq.w += __shfl_xor_sync(0xffffffffu, sq.w, o);
  }
  if (rr == 0) {
    reinterpret_cast<float4*>(dst)[0] = make_float4(su.x, sq.x, su.y, sq.y);
    reinterpret_cast<float4*>(dst)[1] = make_float4(su.z, sq.z, su.w, sq.w);
  }
}

template <bool BIAS, bool RV, bool RES, bool F32, bool H16>
HEDIT_DEVICE void epi_store_rows(const float4* stg, int rq, int rr, const float4& bb, const float4& rvv, const float4 (&rs)[8],
                                 float* o32, size_t ldo4, op_t* o16, size_t ldob4, float2* cst) {
  float4 su = make_float4(0.f, 0.f, 0.f, 0.f), sq = su;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = 4 * i + rr;
    float4 a = stg[r * 8 + (rq ^ (r & 7))];
    if (BIAS) { a.x += bb.x; a.y += bb.y; a.z += bb.z; a.w += bb.w; }
    if (RV) { a.x += rvv.x; a.y += rvv.y; a.z += rvv.z; a.w += rvv.w; }
    if (RES) { a.x += rs[i].x; a.y += rs[i].y; a.z += rs[i].z; a.w += rs[i].w; }
    if (F32) *reinterpret_cast<float4*>(o32 + i * ldo4) = a;
    if (H16) *reinterpret_cast<uint2*>(o16 + i * ldob4) = make_uint2(pack_op2(a.x, a.y), pack_op2(a.z, a.w));
    if (F32) {   // (only fp32 outputs feed a GroupNorm)
      su.x += a.x; su.y += a.y; su.z += a.z; su.w += a.w;
      sq.x = fmaf(a.x, a.x, sq.x); sq.y = fmaf(a.y, a.y, sq.y); sq.z = fmaf(a.z, a.z, sq.z); sq.w = fmaf(a.w, a.w, sq.w);
    }
  }
  if (F32 && cst) epi_store_colstats(su, sq, rr, cst);      // cst is warp-uniform
}

// EPI: 0 = generic write-back (8 epilogue warps), 1 = GEGLU (16 warps, 1 KB staging each), 2 = bias + fp32 residual -> fp32 of the small-K
// projections (16 warps with 4 KB staging each, paid for with one pipeline stage)
template <int BN, bool PAIR, int EPI = 0>
struct GemmCfg {
  static constexpr int BM = 128, BK = 64;
  static constexpr uint32_t A_BYTES = BM * BK * 2;
  static constexpr uint32_t B_BYTES = (PAIR ? BN / 2 : BN) * BK * 2;     // PAIR: each CTA stages half of the W tile
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (PAIR ? (BN > 160 ? 6 : 7) : (BN > 160 ? 4 : (BN > 128 ? 5 : (BN > 64 ? 6 : 8)))) - (EPI == 2 ? 1 : 0);
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 256 /*barriers*/ + (EPI == 2 ? 16 : 8) * 4096 /*epilogue staging*/;
  static constexpr int THREADS = EPI ? 576 : 320;
};

// PAIR = true: cta_group::2.  Two CTAs of one cluster (an SM pair) own two consecutive M tiles of the same N tile and issue
// ONE 256 x BN x 16 tcgen05.mma per K step from the leader: each SM reads its own A tile and only HALF of the W tile from
// its shared memory (the tensor cores exchange the halves), which halves the W shared-memory traffic that bounds the
// single-CTA kernel (SS-mode operand reads + TMA writes ~ 2 x 115 B/clk/SM at 128x160 tiles vs 128 B/clk/SM available).
// Protocol: both CTAs' TMA loads complete on the LEADER's full barrier; the leader's MMA commits multicast to both CTAs'
// empty / accumulator-full barriers; both CTAs' epilogue warps arrive on the leader's accumulator-empty barrier.
// GEGLU = the fused GEGLU write-back as its own instantiation: the generic write-back's double-buffered bias / residual registers (64+)
// are not allocated, which lets the compiler keep all 8 packed GELU evaluations of a chunk in flight (the K = 320 feed-forward
// projection is bound by the latency of that epilogue, not by the tensor pipe).
template <int BN, bool CLUSTER, int EPI = 0>
__global__ void __launch_bounds__(EPI ? 576 : 320, 1) gemm_bf16_tcgen05_kernel(const __grid_constant__ GemmParams p) {
  constexpr bool GEGLU = (EPI == 1);
  // epilogue warps: the GEGLU write-back is instruction-bound and the fp32-residual write-back of the K <= 640 projections is bound by
  // the bytes it keeps in flight (HBM latency), so both get 4 warps per scheduler; EPI == 2 then loads its bias / residual for the
  // CURRENT chunk (thread-level parallelism instead of the generic path's register double-buffering: 112 registers per thread)
  constexpr int EW = EPI ? 16 : 8;
  constexpr int CSTEP = 32 * (EW / 4);    // column distance between two chunks of one warp
  using Cfg = GemmCfg<BN, CLUSTER, EPI>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();     // 128B-swizzle tiles need 1024-byte alignment
  constexpr uint32_t STG_BYTES = (EPI == 2 ? 16 : 8) * 4096;      // epilogue staging: right after the ring, so every 4 KB slice is 1024-byte aligned (TMA store source)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + STG_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;       // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (p.M + 127) >> 7;
  const int n_tiles = (p.N + BN - 1) / BN;
  // work items: single tiles, or (CLUSTER) pairs of M tiles; `rank` selects this CTA's M tile inside a pair
  const uint32_t rank = CLUSTER ? cluster_ctarank() : 0u;
  const int num_tiles = CLUSTER ? ((m_tiles + 1) >> 1) * n_tiles : m_tiles * n_tiles;
  const int tile0 = CLUSTER ? int(blockIdx.x >> 1) : int(blockIdx.x);
  const int tstep = CLUSTER ? int(gridDim.x >> 1) : int(gridDim.x);
  auto tile_m0 = [&](int tile) { return CLUSTER ? (((tile / n_tiles) * 2 + int(rank)) << 7) : ((tile / n_tiles) << 7); };
  // split-K: the loops below iterate WORK items w = split * num_tiles + tile (splits == 1: w == tile)
  const int nsplit = (!CLUSTER && p.splits > 1) ? p.splits : 1;
  const int num_work = num_tiles * nsplit;
  auto work_kb0 = [&](int w) { return (w / num_tiles) * p.kb_per_split; };
  auto work_kb1 = [&](int w) { return nsplit > 1 ? min(p.num_kb, (w / num_tiles + 1) * p.kb_per_split) : p.num_kb; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], CLUSTER ? 2 * EW : EW); }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (CLUSTER) { tmem_alloc_2sm(tmem_slot, 512); tmem_relinquish_2sm(); }
    else { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (CLUSTER) cluster_sync_all();        // peer barriers are initialised before any multicast / remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch: everything above (barriers, tensor memory, descriptor prefetch) ran while the previous kernel was
  // still finishing; from here on the kernel reads what that kernel wrote
  pdl_wait();
  pdl_launch();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (whole warp runs the loop so the
    // addressing stays in uniform registers; one elected lane issues the copies)
    int stage = 0; uint32_t phase = 0;
    for (int work = tile0; work < num_work; work += tstep) {
      const int tile = work % num_tiles;
      const int kb0 = nsplit > 1 ? work_kb0(work) : 0, kb1 = work_kb1(work);
      const int m0 = tile_m0(tile);
      const int n0 = (tile % n_tiles) * BN;
      int s0 = 0, y0 = 0, x0 = 0;
      if (p.a_mode != A_LINEAR) {
        const int hw = p.conv_H * p.conv_W;
        s0 = m0 / hw;
        y0 = (m0 % hw) / p.conv_W;
        x0 = (m0 % hw) % p.conv_W;       // non-zero only for images wider than one 128-row tile (VAE decoder, W = 256 / 512)
      }
      // The producer's issue rate is on the critical path of the conv GEMMs, so the K loop is specialised per addressing mode at compile
      // time (constant tap arithmetic, no mode tests inside the loop).
      auto k_loop = [&](auto mode_c) {
        constexpr int MODE = decltype(mode_c)::value;
        constexpr int KW = (MODE == A_CONV2X2) ? 2 : 3;
        int tap = (MODE == A_LINEAR) ? 0 : kb0 / p.cin_blocks, cb = (MODE == A_LINEAR) ? 0 : kb0 % p.cin_blocks;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
            uint8_t* sb = sa + Cfg::A_BYTES;
            const int ky = tap / KW, kx = tap - ky * KW;
            int c1 = 0, c2 = 0, c3 = 0, c4 = 0;          // TMA coordinates of the A box beyond the channel coordinate
            if (MODE == A_CONV3X3) { c1 = x0 + kx - 1; c2 = y0 + ky - 1; c3 = s0; }
            else if (MODE == A_CONV2X2) { c1 = x0 + kx + p.conv_ox; c2 = y0 + ky + p.conv_oy; c3 = s0; }
            else if (MODE == A_CONV3X3S2) {
              // stride 2: input x = 2X + kx - pad_left -> (parity, coarse offset) in the space-to-depth view
              const int px = p.conv_pad01 ? (kx == 1 ? 1 : 0) : (kx == 1 ? 0 : 1), dx = p.conv_pad01 ? (kx == 2 ? 1 : 0) : (kx == 0 ? -1 : 0);
              const int py = p.conv_pad01 ? (ky == 1 ? 1 : 0) : (ky == 1 ? 0 : 1), dy = p.conv_pad01 ? (ky == 2 ? 1 : 0) : (ky == 0 ? -1 : 0);
              c1 = x0 + dx; c2 = py; c3 = y0 + dy; c4 = px * p.conv_cin;
            }
            if (CLUSTER) {
              // both CTAs' bytes are reported to the leader's barrier (the leader's MMA consumes both halves)
              const uint32_t lfull = mapa_u32(smem_u32(&full_bar[stage]), 0);
              if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
              if (MODE == A_LINEAR) tma_load_2d_2sm(sa, &p.tmA, lfull, kb * 64, m0);
              else if (MODE == A_CONV3X3S2) tma_load_5d_2sm(sa, &p.tmA, lfull, c4 + cb * 64, c1, c2, c3, s0);
              else tma_load_4d_2sm(sa, &p.tmA, lfull, cb * 64, c1, c2, c3);
              tma_load_2d_2sm(sb, &p.tmB, lfull, kb * 64, n0 + int(rank) * (BN / 2));
            } else {
              mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
              if (MODE == A_LINEAR) tma_load_2d(sa, &p.tmA, &full_bar[stage], kb * 64, m0);
              else if (MODE == A_CONV3X3S2) tma_load_5d(sa, &p.tmA, &full_bar[stage], c4 + cb * 64, c1, c2, c3, s0);
              else tma_load_4d(sa, &p.tmA, &full_bar[stage], cb * 64, c1, c2, c3);
              if (p.b_full_box) {
                tma_load_2d(sb, &p.tmB, &full_bar[stage], kb * 64, n0);
              } else {       // the W map's box is BN/2 rows (shared with the pair variant): two loads
                tma_load_2d(sb, &p.tmB, &full_bar[stage], kb * 64, n0);
                tma_load_2d(sb + (BN / 2) * 128, &p.tmB, &full_bar[stage], kb * 64, n0 + BN / 2);
              }
            }
          }
          __syncwarp();
          if (MODE != A_LINEAR && ++cb == p.cin_blocks) { cb = 0; ++tap; }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      };
      if (p.a_mode == A_CONV3X3) k_loop(std::integral_constant<int, A_CONV3X3>{});
      else if (p.a_mode == A_LINEAR) k_loop(std::integral_constant<int, A_LINEAR>{});
      else if (p.a_mode == A_CONV2X2) k_loop(std::integral_constant<int, A_CONV2X2>{});
      else k_loop(std::integral_constant<int, A_CONV3X3S2>{});
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform loop, elected lane issues);
    // in PAIR mode only the leader CTA issues (cta_group::2 instructions drive both SMs' tensor cores)
    if (!CLUSTER || rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(CLUSTER ? 256 : 128, BN, 0, 0);
      const uint32_t a_lo0 = umma_desc_lo_kmajor(smem_u32(smem));
      const uint32_t b_lo0 = umma_desc_lo_kmajor(smem_u32(smem) + Cfg::A_BYTES);
      int stage = 0; uint32_t phase = 0; int it = 0;
      // The issue thread is the critical resource at 128x160 tiles (4 MMAs = 320 tensor cycles per K block): probe the NEXT
      // stage's barrier (non-blocking) before issuing the current stage's MMAs so its latency hides behind them.
      bool ready = mbar_test(&full_bar[0], 0);
      for (int work = tile0; work < num_work; work += tstep, ++it) {
        const int kb0 = nsplit > 1 ? work_kb0(work) : 0, kb1 = work_kb1(work);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        const bool last_tile = (work + tstep >= num_work);
        for (int kb = kb0; kb < kb1; ++kb) {
          if (!ready) mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          int nstage = stage + 1; uint32_t nphase = phase;
          if (nstage == STAGES) { nstage = 0; nphase ^= 1; }
          const bool more = !(last_tile && kb + 1 == kb1);
          ready = more ? mbar_test(&full_bar[nstage], nphase) : false;
          if (elect_one()) {
            const uint32_t la = a_lo0 + stage * (Cfg::STAGE_BYTES >> 4);
            const uint32_t lb = b_lo0 + stage * (Cfg::STAGE_BYTES >> 4);
            if (CLUSTER) {
              umma_f16_ss_2sm(d_tmem, umma_desc_make(la), umma_desc_make(lb), idesc, kb != kb0);
              umma_f16_ss_2sm(d_tmem, umma_desc_make(la + 2), umma_desc_make(lb + 2), idesc, 1);
              umma_f16_ss_2sm(d_tmem, umma_desc_make(la + 4), umma_desc_make(lb + 4), idesc, 1);
              umma_f16_ss_2sm(d_tmem, umma_desc_make(la + 6), umma_desc_make(lb + 6), idesc, 1);
              umma_commit_2sm(&empty_bar[stage], 3);
            } else {
              umma_f16_ss(d_tmem, umma_desc_make(la), umma_desc_make(lb), idesc, kb != kb0);
              umma_f16_ss(d_tmem, umma_desc_make(la + 2), umma_desc_make(lb + 2), idesc, 1);
              umma_f16_ss(d_tmem, umma_desc_make(la + 4), umma_desc_make(lb + 4), idesc, 1);
              umma_f16_ss(d_tmem, umma_desc_make(la + 6), umma_desc_make(lb + 6), idesc, 1);
              umma_commit(&empty_bar[stage]);
            }
          }
          __syncwarp();
          stage = nstage; phase = nphase;
        }
        if (elect_one()) { if (CLUSTER) umma_commit_2sm(&tfull_bar[acc], 3); else umma_commit(&tfull_bar[acc]); }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: 8 warps; warp w owns TMEM lane quarter
    // (w & 3) and every second 32-column chunk.  Accumulators are read row-per-thread (TMEM lane == row), transposed
    // through a per-warp XOR-swizzled 32x32 fp32 staging tile, and written back with 8 lanes per row so that every
    // global load/store instruction touches 4 fully used 128-byte lines (the row-per-thread pattern touches 32).
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const GemmEpilogue& e = p.ep;
    float4* stg = reinterpret_cast<float4*>(smem + STAGES * Cfg::STAGE_BYTES) + (warp - 2) * (GEGLU ? 64 : 256);   // [32 rows][8 quads] (GEGLU: [32 rows][32 B])
    const int rq = lane & 7, rr = lane >> 3;           // read-back role (non-GEGLU): quad, row-in-group-of-4
    // feature combination of this launch -> specialised write-back loop (0 = generic path)
    const bool has_b = e.bias != nullptr, has_rv = e.rowvec != nullptr, has_res = e.residual != nullptr;
    const bool has32 = e.out_f32 != nullptr, has16 = e.out_bf16 != nullptr;
    int mode = 0;
    if (!GEGLU) {
      if (has_b && !has_rv && !has_res && has32 && !has16) mode = 1;
      else if (has_b && !has_rv && has_res && has32 && !has16) mode = 2;
      else if (has_b && has_rv && !has_res && has32 && !has16 && (e.rows_per_group & 31) == 0) mode = 3;
      else if (!has_b && !has_rv && !has_res && !has32 && has16) mode = 4;
      else if (has_b && !has_rv && has_res && !has32 && has16) mode = 5;
      else if (!has_b && !has_rv && !has_res && has32 && !has16) mode = 6;
      else if (has_b && !has_rv && !has_res && !has32 && has16) mode = 7;
      else if (!has_b && !has_rv && has_res && has32 && !has16) mode = 8;
    }
    int it = 0;
    for (int work = tile0; work < num_work; work += tstep, ++it) {
      const int tile = work % num_tiles;
      float* const out_f32 = e.out_f32 ? e.out_f32 + (nsplit > 1 ? size_t(work / num_tiles) * p.split_stride : size_t(0)) : nullptr;
      const int m0 = tile_m0(tile);
      const int n0 = (tile % n_tiles) * BN;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int rbase = m0 + quarter * 32;
      const bool rows_full = (rbase + 32 <= p.M);
      const uint32_t t_row = tmem_base + acc * 256 + (uint32_t(quarter * 32) << 16);
      const size_t row0 = size_t(rbase + rr);
      // Operands of the write-back that do not depend on the accumulator (bias, time-embedding row, residual) are double-buffered: the
      // set of chunk c+1 is requested BEFORE the write-back of chunk c, so its global-load latency hides behind that write-back and the
      // next chunk's TMEM read (the epilogue of the small-K layers is latency-bound, not bandwidth-bound).
      float4 rs[8], rsN[8];
      float4 bb = make_float4(0.f, 0.f, 0.f, 0.f), rvv = bb, bbN = bb, rvvN = bb;
      auto prefetch = [&](int col, float4& b_, float4& rv_, float4 (&r_)[8]) {
        if (mode != 0 && rows_full && col + 32 <= p.N) {
          const int cq = col + 4 * rq;
          if (has_b) b_ = *reinterpret_cast<const float4*>(e.bias + cq);
          if (mode == 3) rv_ = *reinterpret_cast<const float4*>(e.rowvec + size_t(rbase / e.rows_per_group) * e.ldrv + cq);
          if (has_res) {
            const float* rp = e.residual + row0 * e.ldr + cq;
#pragma unroll
            for (int i = 0; i < 8; ++i) r_[i] = *reinterpret_cast<const float4*>(rp + size_t(4 * i) * e.ldr);
          }
        }
      };
      if (EPI != 2) prefetch(n0 + half * 32, bb, rvv, rs);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      // 16-bit outputs without bias / residual (q|k|v, q, text k|v projections): 64 columns per iteration, converted to 16 bits BEFORE the
      // staging transpose (so the 4 KB tile holds 32 rows x 64 columns) and written back as full 128-byte lines -- half the latency-bound
      // iterations and half the store instructions of the 32-column path.
      if (e.diag_skip == 1 || e.diag_skip == 2) {
        if (e.diag_skip == 2) {     // diagnostics: read the accumulator like a write-back would, discard it (TMEM-port contention probe)
          uint32_t acc_x = 0;
#pragma unroll 1
          for (int c0 = half * 32; c0 < BN; c0 += 64) {
            uint32_t raw[32];
            tmem_ld32(t_row + c0, raw);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc_x ^= raw[j];
          }
          if (acc_x == 0x7fc12345u && e.out_bf16) e.out_bf16[0] = to_op(1.f);
        }
      } else if (GEGLU) {
        // chunk = 16 value columns followed by their 16 gate columns -> 16 outputs (bias applied before the gate), packed to 16 bits
        // BEFORE the staging transpose ([32 rows][32 B], 16-byte slots XOR-swizzled) and written back as 16 rows x 32 B per instruction
        uint4* stg16 = reinterpret_cast<uint4*>(stg);
        const int hs = lane & 1, hr = lane >> 1;
        // warp (quarter, half = column group 0..3) owns every fourth 32-column chunk
#pragma unroll 1
        for (int c0 = half * 32; c0 < BN; c0 += CSTEP) {
          const int col = n0 + c0;
          if (col >= p.N) break;                      // warp-uniform
          uint32_t raw[32];
          tmem_ld32(t_row + c0, raw);
          const float4* bp = reinterpret_cast<const float4*>(e.bias + col);     // (GEGLU launches always carry a bias; N % 32 == 0)
          float4 gbias[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) gbias[j] = __ldg(bp + j);
          tmem_ld_wait();
          uint32_t pk[8];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 bv = gbias[q], bg = gbias[4 + q];
            const float2 g0 = gelu_fast_f2(__fadd2_rn(make_float2(__uint_as_float(raw[16 + 4 * q]), __uint_as_float(raw[17 + 4 * q])), make_float2(bg.x, bg.y)));
            const float2 g1 = gelu_fast_f2(__fadd2_rn(make_float2(__uint_as_float(raw[18 + 4 * q]), __uint_as_float(raw[19 + 4 * q])), make_float2(bg.z, bg.w)));
            const float2 o0 = __fmul2_rn(__fadd2_rn(make_float2(__uint_as_float(raw[4 * q]), __uint_as_float(raw[4 * q + 1])), make_float2(bv.x, bv.y)), g0);
            const float2 o1 = __fmul2_rn(__fadd2_rn(make_float2(__uint_as_float(raw[4 * q + 2]), __uint_as_float(raw[4 * q + 3])), make_float2(bv.z, bv.w)), g1);
            pk[2 * q] = pack_op2(o0.x, o0.y); pk[2 * q + 1] = pack_op2(o1.x, o1.y);
          }
          const int sw = (lane >> 2) & 1;
          stg16[lane * 2 + (0 ^ sw)] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          stg16[lane * 2 + (1 ^ sw)] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          __syncwarp();
          op_t* go = e.out_bf16 + size_t(rbase + hr) * e.ldob + (col >> 1) + 8 * hs;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int r = 16 * i + hr;
            const uint4 a = stg16[r * 2 + (hs ^ ((r >> 2) & 1))];
            if ((rows_full || rbase + r < p.M) && e.diag_skip != 3) *reinterpret_cast<uint4*>(go + size_t(16 * i) * e.ldob) = a;   // (3: probe without the stores)
          }
          __syncwarp();
        }
      } else {
      const bool wide16 = (BN == 256) && mode == 4 && rows_full && ((p.N - n0) >= BN || ((p.N - n0) & 63) == 0);
      if (wide16) {
        op_t* stg16 = reinterpret_cast<op_t*>(stg);
        const int wu = lane & 7, wr = lane >> 3;
#pragma unroll 1
        for (int c0 = half * 64; c0 < BN; c0 += 128) {
          const int col = n0 + c0;
          if (col + 64 > p.N) break;                  // warp-uniform
          uint32_t raw[64];
          tmem_ld32(t_row + c0, *reinterpret_cast<uint32_t(*)[32]>(&raw[0]));
          tmem_ld32(t_row + c0 + 32, *reinterpret_cast<uint32_t(*)[32]>(&raw[32]));
          tmem_ld_wait();
          if (p.use_tmd) { if (lane == 0) tma_store_wait_read0(); __syncwarp(); }      // the previous bulk store has read the staging tile
#pragma unroll
          for (int u = 0; u < 8; ++u)
            *reinterpret_cast<uint4*>(stg16 + lane * 64 + ((u ^ (lane & 7)) << 3)) =
                make_uint4(pack_op2(__uint_as_float(raw[8 * u]), __uint_as_float(raw[8 * u + 1])), pack_op2(__uint_as_float(raw[8 * u + 2]), __uint_as_float(raw[8 * u + 3])),
                           pack_op2(__uint_as_float(raw[8 * u + 4]), __uint_as_float(raw[8 * u + 5])), pack_op2(__uint_as_float(raw[8 * u + 6]), __uint_as_float(raw[8 * u + 7])));
          if (p.use_tmd) {
            // the staging tile IS the 128B-swizzled box layout (16-byte slot ^ (row & 7)): one bulk tensor store per 32 x 64 tile instead
            // of 8 STG.128 per lane; the source is reused only after the TMA engine has read it (wait at the top of the next iteration,
            // so the store overlaps the next chunk's TMEM read)
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) { tma_store_2d(&p.tmD, stg16, col, rbase); tma_store_commit(); }
            __syncwarp();
          } else {
            __syncwarp();
            op_t* o16 = e.out_bf16 + size_t(rbase + wr) * e.ldob + col + 8 * wu;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = 4 * i + wr;
              *reinterpret_cast<uint4*>(o16 + size_t(4 * i) * e.ldob) = *reinterpret_cast<const uint4*>(stg16 + r * 64 + ((wu ^ (r & 7)) << 3));
            }
            __syncwarp();
          }
        }
        if (p.use_tmd && lane == 0) tma_store_wait0();      // global writes of this tile are complete before the kernel can end
      } else
#pragma unroll 1
      for (int c0 = half * 32; c0 < BN; c0 += CSTEP) {
        const int col = n0 + c0;
        if (col >= p.N) break;                      // warp-uniform
        uint32_t raw[32];
        tmem_ld32(t_row + c0, raw);
        if (EPI == 2) prefetch(col, bb, rvv, rs);   // this chunk's bias / residual, in flight during the TMEM read
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q)
          stg[lane * 8 + (q ^ (lane & 7))] = make_float4(__uint_as_float(raw[4 * q]), __uint_as_float(raw[4 * q + 1]),
                                                         __uint_as_float(raw[4 * q + 2]), __uint_as_float(raw[4 * q + 3]));
        __syncwarp();
        if (e.up_W > 0) {
          // fused-upsample phase: bias + fp32 store to the fine-grid row of every coarse pixel, column statistics per (sample, phase)
          const int cq = col + 4 * rq, hw = e.up_W * e.up_H;
          float4 su = make_float4(0.f, 0.f, 0.f, 0.f), sq = su;
          const bool cols_ok = (cq + 4 <= p.N);
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (e.bias && cols_ok) b4 = *reinterpret_cast<const float4*>(e.bias + cq);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = 4 * i + rr, m = rbase + r;
            float4 a = stg[r * 8 + (rq ^ (r & 7))];
            a.x += b4.x; a.y += b4.y; a.z += b4.z; a.w += b4.w;
            if (m < p.M && cols_ok) {
              const int sm_ = m / hw, rem = m - sm_ * hw, y = rem / e.up_W, x = rem - y * e.up_W;
              const size_t frow = (size_t(sm_) * 2 * e.up_H + 2 * y + e.up_py) * (2 * e.up_W) + 2 * x + e.up_px;
              *reinterpret_cast<float4*>(out_f32 + frow * e.ldo + cq) = a;
              su.x += a.x; su.y += a.y; su.z += a.z; su.w += a.w;
              sq.x = fmaf(a.x, a.x, sq.x); sq.y = fmaf(a.y, a.y, sq.y); sq.z = fmaf(a.z, a.z, sq.z); sq.w = fmaf(a.w, a.w, sq.w);
            }
          }
          if (e.colstats && rbase < p.M && col + 32 <= p.N) {
            const int nbc = hw >> 5, blk = rbase >> 5, sb = blk / nbc;
            epi_store_colstats(su, sq, rr, e.colstats + size_t(blk + sb * 3 * nbc + (e.up_py * 2 + e.up_px) * nbc) * p.N + cq);
          }
          __syncwarp();
          continue;
        }
        if (EPI != 2 && c0 + CSTEP < BN) prefetch(col + CSTEP, bbN, rvvN, rsN);      // next chunk of this warp (warp-uniform condition)
        const int cq = col + 4 * rq;
        if (mode != 0 && rows_full && col + 32 <= p.N) {
          float* o32 = has32 ? out_f32 + row0 * e.ldo + cq : nullptr;
          op_t* o16 = has16 ? e.out_bf16 + row0 * e.ldob + cq : nullptr;
          const size_t l32 = size_t(4) * e.ldo, l16 = size_t(4) * e.ldob;
          float2* cst = e.colstats ? e.colstats + size_t(rbase >> 5) * p.N + cq : nullptr;
          switch (mode) {
            case 1: epi_store_rows<true, false, false, true, false>(stg, rq, rr, bb, rvv, rs, o32, l32, o16, l16, cst); break;
            case 2: epi_store_rows<true, false, true, true, false>(stg, rq, rr, bb, rvv, rs, o32, l32, o16, l16, cst); break;
            case 3: epi_store_rows<true, true, false, true, false>(stg, rq, rr, bb, rvv, rs, o32, l32, o16, l16, cst); break;
            case 4: epi_store_rows<false, false, false, false, true>(stg, rq, rr, bb, rvv, rs, o32, l32, o16, l16, cst); break;
            case 5: epi_store_rows<true, false, true, false, true>(stg, rq, rr, bb, rvv, rs, o32, l32, o16, l16, cst); break;
            case 6: epi_store_rows<false, false, false, true, false>(stg, rq, rr, bb, rvv, rs, o32, l32, o16, l16, cst); break;
            case 7: epi_store_rows<true, false, false, false, true>(stg, rq, rr, bb, rvv, rs, o32, l32, o16, l16, cst); break;
            default: epi_store_rows<false, false, true, true, false>(stg, rq, rr, bb, rvv, rs, o32, l32, o16, l16, cst); break;
          }
        } else {
          // generic path: partial tiles, ragged N, unusual feature combinations
          float gsu[4] = {0.f, 0.f, 0.f, 0.f}, gsq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
          for (int i = 0; i < 8; ++i) {
            const int r = 4 * i + rr, row = rbase + r;
            const float4 a4 = stg[r * 8 + (rq ^ (r & 7))];
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int cc = cq + k;
              if (row < p.M && cc < p.N) {
                float x = av[k];
                if (e.bias) x += e.bias[cc];
                if (e.rowvec) x += e.rowvec[size_t(row / e.rows_per_group) * e.ldrv + cc];
                if (e.residual) x += e.residual[size_t(row) * e.ldr + cc];
                gsu[k] += x; gsq[k] = fmaf(x, x, gsq[k]);
                if (out_f32) {
                  if (e.nchw_hw) out_f32[(size_t(row / e.nchw_hw) * p.N + cc) * e.nchw_hw + (row % e.nchw_hw)] = x;
                  else out_f32[size_t(row) * e.ldo + cc] = x;
                }
                if (e.out_bf16) e.out_bf16[size_t(row) * e.ldob + cc] = to_op(x);
              }
            }
          }
          if (e.colstats && rbase < p.M && col + 32 <= p.N)      // (colstats is only requested for N % 32 == 0)
            epi_store_colstats(make_float4(gsu[0], gsu[1], gsu[2], gsu[3]), make_float4(gsq[0], gsq[1], gsq[2], gsq[3]), rr,
                               e.colstats + size_t(rbase >> 5) * p.N + cq);
        }
        __syncwarp();
        if (EPI != 2) {
          bb = bbN; rvv = rvvN;
#pragma unroll
          for (int i = 0; i < 8; ++i) rs[i] = rsN[i];
        }
      }
      }   // !GEGLU
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CLUSTER) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));     // the leader's MMA waits for both CTAs
        else mbar_arrive(&tempty_bar[acc]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CLUSTER) cluster_sync_all();        // no CTA exits while its peer may still multicast into it / arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    if (CLUSTER) tmem_dealloc_2sm(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// Split-K second pass: out = sum_s ws[s] (fixed order) + bias + time-embedding row + residual -> fp32 and / or 16-bit, plus the column
// statistics of the fp32 output per 32-row block (same layout as the GEMM epilogue's).  block = 256 threads = 32 rows x 8 column quads;
// grid (ceil(N / 32), ceil(M / 32)).
struct SplitKReduceParams {
  const float* ws; size_t split_stride; int splits;
  int M, N;
  GemmEpilogue ep;
};
static __global__ void splitk_reduce_kernel(const SplitKReduceParams p) {
  pdl_wait(); pdl_launch();
  __shared__ float4 ssu[32][8], ssq[32][8];
  const int q = threadIdx.x & 7, r = threadIdx.x >> 3;
  const int row = blockIdx.y * 32 + r, col = blockIdx.x * 32 + 4 * q;
  const GemmEpilogue& e = p.ep;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool ok = row < p.M && col + 4 <= p.N;
  if (ok) {
    const float* w = p.ws + size_t(row) * p.N + col;
    a = *reinterpret_cast<const float4*>(w);
    for (int s = 1; s < p.splits; ++s) {
      const float4 t = *reinterpret_cast<const float4*>(w + size_t(s) * p.split_stride);
      a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
    }
    if (e.bias) { const float4 b = *reinterpret_cast<const float4*>(e.bias + col); a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
    if (e.rowvec) { const float4 b = *reinterpret_cast<const float4*>(e.rowvec + size_t(row / e.rows_per_group) * e.ldrv + col); a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
    if (e.residual) { const float4 b = *reinterpret_cast<const float4*>(e.residual + size_t(row) * e.ldr + col); a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
    if (e.out_f32) *reinterpret_cast<float4*>(e.out_f32 + size_t(row) * e.ldo + col) = a;
    if (e.out_bf16) *reinterpret_cast<uint2*>(e.out_bf16 + size_t(row) * e.ldob + col) = make_uint2(pack_op2(a.x, a.y), pack_op2(a.z, a.w));
  }
  if (e.colstats) {
    ssu[r][q] = ok ? a : make_float4(0.f, 0.f, 0.f, 0.f);
    ssq[r][q] = ok ? make_float4(a.x * a.x, a.y * a.y, a.z * a.z, a.w * a.w) : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    if (r == 0 && col + 4 <= p.N && blockIdx.y * 32 < p.M) {
      float4 su = ssu[0][q], sq = ssq[0][q];
      for (int k = 1; k < 32; ++k) {
        const float4 u = ssu[k][q], v = ssq[k][q];
        su.x += u.x; su.y += u.y; su.z += u.z; su.w += u.w;
        sq.x += v.x; sq.y += v.y; sq.z += v.z; sq.w += v.w;
      }
      float2* dst = e.colstats + size_t(blockIdx.y) * p.N + col;
      reinterpret_cast<float4*>(dst)[0] = make_float4(su.x, sq.x, su.y, sq.y);
      reinterpret_cast<float4*>(dst)[1] = make_float4(su.z, sq.z, su.w, sq.w);
    }
  }
}

}  // namespace hedit
