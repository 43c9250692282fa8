// Thin inline-PTX wrappers for the sm_100a features the kernels use: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld / fences) and the UMMA descriptors.  No CUTLASS/CuTe dependency.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace hedit {

// 16-bit tensor-core operand type.  Default fp16 (11-bit significand): same tcgen05 kind::f16 throughput as bf16
// with 8x smaller operand rounding error, which is what bounds parity against the reference's fp32 UNet.  Range is
// not an issue for SD-1.x (normalised activations, fp32 residual stream and accumulators).  -DHEDIT_OPERAND_BF16
// switches every kernel to bf16 operands.
#ifdef HEDIT_OPERAND_BF16
typedef __nv_bfloat16 op_t;
#define HEDIT_UMMA_FMT 1u
#define HEDIT_TMAP_DTYPE CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
#define HEDIT_OPERAND_NAME "bf16"
#else
typedef __half op_t;
#define HEDIT_UMMA_FMT 0u
#define HEDIT_TMAP_DTYPE CU_TENSOR_MAP_DATA_TYPE_FLOAT16
#define HEDIT_OPERAND_NAME "fp16"
#endif

#define HEDIT_DEVICE __device__ __forceinline__

HEDIT_DEVICE uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

HEDIT_DEVICE uint32_t lane_id() { return threadIdx.x & 31u; }

HEDIT_DEVICE bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
HEDIT_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
HEDIT_DEVICE void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
HEDIT_DEVICE void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
HEDIT_DEVICE void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
HEDIT_DEVICE bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe (for event-driven issue loops)
HEDIT_DEVICE bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug traps (the launch reports an error) instead of hanging the GPU box.
HEDIT_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("hedit: mbarrier wait timeout (block %d thread %d bar %p parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// generic-proxy smem writes -> visible to the async proxy (TMA / tcgen05.mma reading smem)
// programmatic dependent launch (launch.h): block until the preceding grid in the stream has completed and flushed its memory; then let
// the following grid become resident
HEDIT_DEVICE void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
HEDIT_DEVICE void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
HEDIT_DEVICE void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------- TMA
HEDIT_DEVICE void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// bulk tensor STORE shared -> global (bulk-group completion): the staged tile must be visible to the async proxy first
// (fence_proxy_async_smem after the generic-proxy writes); wait_group.read frees the shared-memory source for reuse
HEDIT_DEVICE void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
HEDIT_DEVICE void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
HEDIT_DEVICE void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
HEDIT_DEVICE void tma_store_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
HEDIT_DEVICE void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// multicast variant: the box lands at the same CTA-relative smem offset of every CTA in cta_mask and completes tx bytes on
// the mbarrier at the same offset in each of them
HEDIT_DEVICE void tma_load_2d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
HEDIT_DEVICE void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
HEDIT_DEVICE void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
HEDIT_DEVICE void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
HEDIT_DEVICE void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {   // whole warp, .sync.aligned
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
HEDIT_DEVICE void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
HEDIT_DEVICE void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
HEDIT_DEVICE void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
HEDIT_DEVICE void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], bf16/fp16 inputs (kind::f16), fp32 accumulate.  One thread issues.
HEDIT_DEVICE void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: the A operand is read from tensor memory (128 lanes = rows, 16-bit K elements packed two per 32-bit
// column, i.e. 8 columns per K = 16 step) -- used for P.V with the probabilities written to TMEM by tcgen05.st
HEDIT_DEVICE void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
// (implicitly performs tcgen05.fence::before_thread_sync)
HEDIT_DEVICE void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// same, arriving on the barrier at this offset in every CTA of cta_mask (cluster-shared pipeline stages)
HEDIT_DEVICE void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
// ---- cta_group::2 (CTA pair) primitives -------------------------------------------------------------------------------
HEDIT_DEVICE uint32_t mapa_u32(uint32_t saddr, uint32_t cta_rank) {     // shared::cluster address of `saddr` in CTA `cta_rank`
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(cta_rank));
  return r;
}
HEDIT_DEVICE void mbar_arrive_cluster(uint32_t cluster_addr) {           // arrive on a (possibly remote) barrier
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
HEDIT_DEVICE void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
HEDIT_DEVICE void tmem_relinquish_2sm() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
HEDIT_DEVICE void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem, both CTAs] (+)= A (128 rows per CTA) * B (N/2 rows per CTA); issued by the leader CTA only
HEDIT_DEVICE void umma_f16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
HEDIT_DEVICE void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
// TMA loads whose completion is reported to a barrier given as a shared::cluster address (the pair leader's)
HEDIT_DEVICE void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
HEDIT_DEVICE void tma_load_4d_2sm(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
HEDIT_DEVICE void tma_load_5d_2sm(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
HEDIT_DEVICE uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
HEDIT_DEVICE void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns starting at taddr.
HEDIT_DEVICE void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
HEDIT_DEVICE void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
HEDIT_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// registers -> TMEM: 32 lanes x 16 consecutive 32-bit columns
HEDIT_DEVICE void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
// registers -> TMEM: 32 lanes x 32 consecutive 32-bit columns
HEDIT_DEVICE void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
HEDIT_DEVICE void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- UMMA descriptors
// Shared-memory matrix descriptor, 128-byte swizzle (the layout TMA writes with CU_TENSOR_MAP_SWIZZLE_128B).
//   bits [0,14)  start address >> 4        bits [16,30) leading-dim byte offset >> 4
//   bits [32,46) stride-dim byte offset>>4 bits [46,48) version = 1 (sm_100)
//   bits [61,64) layout type: 2 = SWIZZLE_128B
// K-major operand  : rows of 64 bf16 (128 B); 8-row groups are SBO = 1024 B apart; LBO unused (=1).
// MN-major operand : 64 contiguous MN elements per 128-B row; next 64-MN chunk at LBO; 8-k groups at SBO.
HEDIT_DEVICE uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
HEDIT_DEVICE uint64_t umma_desc_kmajor_sw128(uint32_t saddr) { return umma_smem_desc(saddr, 16, 1024); }
// Split form for issue loops: the high word is a compile-time constant, the low word is (addr >> 4) | LBO field, and
// stepping K by 16 elements inside a 128-byte swizzle atom adds 2 to the low word.
constexpr uint32_t kDescHiSw128 = (1024u >> 4) | (1u << 14) | (2u << 29);          // SBO=1024, version=1, SWIZZLE_128B
HEDIT_DEVICE uint32_t umma_desc_lo_kmajor(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
HEDIT_DEVICE uint32_t umma_desc_lo(uint32_t saddr, uint32_t lbo_bytes) { return ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16); }
HEDIT_DEVICE uint64_t umma_desc_make(uint32_t lo) { return (static_cast<uint64_t>(kDescHiSw128) << 32) | lo; }

// Instruction descriptor (kind::f16): fp32 accumulate, 16-bit A/B.
//   [4,6) c_format=1(F32)  [7,10) a_format (0=F16, 1=BF16)  [10,13) b_format
//   [15] a_major (0=K,1=MN) [16] b_major            [17,23) N>>3          [24,29) M>>4
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (HEDIT_UMMA_FMT << 7) | (HEDIT_UMMA_FMT << 10) | (uint32_t(a_mn_major) << 15) | (uint32_t(b_mn_major) << 16) |
         (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

// ---------------------------------------------------------------- misc math / packing
HEDIT_DEVICE uint32_t pack_op2(float lo, float hi) {
#ifdef HEDIT_OPERAND_BF16
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
#else
  __half2 v = __floats2half2_rn(lo, hi);
#endif
  return *reinterpret_cast<uint32_t*>(&v);
}
HEDIT_DEVICE op_t to_op(float x) {
#ifdef HEDIT_OPERAND_BF16
  return __float2bfloat16(x);
#else
  return __float2half_rn(x);
#endif
}
HEDIT_DEVICE float op_to_float(op_t x) {
#ifdef HEDIT_OPERAND_BF16
  return __bfloat162float(x);
#else
  return __half2float(x);
#endif
}
HEDIT_DEVICE float2 op2_to_float2(uint32_t u) {
#ifdef HEDIT_OPERAND_BF16
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
#else
  return __half22float2(*reinterpret_cast<__half2*>(&u));
#endif
}
// x * sigmoid(x) with the fast reciprocal (2 ulp): the result is rounded to a 16-bit operand or feeds fp32 sums of thousands of terms
HEDIT_DEVICE float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
HEDIT_DEVICE float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
// exact-GELU to |error| < 2e-7 (Abramowitz-Stegun 7.1.26 for erf; gelu(x) = 0.5 x (1 + erf(x / sqrt2))) on two values with the packed
// fp32x2 instructions of sm_100: 11 FMA-pipe instructions + 4 MUFU per pair.  With z = |x| sqrt(log2(e) / 2):
//   t = 1 / (1 + p' z)  (p' = 0.3275911 / sqrt(log2 e), bare rcp: the argument is >= 1),   exp(-x^2 / 2) = 2^(-z z),
//   erf(|x| / sqrt2) = 1 - poly(t) 2^(-z z),   gelu = hx + |hx| erf(|x| / sqrt2)  with hx = x / 2  (the sign of x folded into |hx|).
// The fused GEGLU epilogue of the K = 320 / 640 feed-forward layers is bound by the issue rate of its two warps per scheduler.
HEDIT_DEVICE float2 gelu_fast_f2(float2 x) {
  const float2 z = __fmul2_rn(make_float2(fabsf(x.x), fabsf(x.y)), make_float2(0.8493218002880191f, 0.8493218002880191f));
  const float2 d = __ffma2_rn(make_float2(0.2727374808792225f, 0.2727374808792225f), z, make_float2(1.0f, 1.0f));
  float2 t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(d.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(d.y));
  float2 poly = __ffma2_rn(make_float2(1.061405429f, 1.061405429f), t, make_float2(-1.453152027f, -1.453152027f));
  poly = __ffma2_rn(poly, t, make_float2(1.421413741f, 1.421413741f));
  poly = __ffma2_rn(poly, t, make_float2(-0.284496736f, -0.284496736f));
  poly = __ffma2_rn(poly, t, make_float2(0.254829592f, 0.254829592f));
  poly = __fmul2_rn(poly, t);
  const float2 a = __fmul2_rn(make_float2(-z.x, -z.y), z);
  float ex0, ex1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex0) : "f"(a.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex1) : "f"(a.y));
  const float2 ea = __ffma2_rn(make_float2(-poly.x, -poly.y), make_float2(ex0, ex1), make_float2(1.0f, 1.0f));
  const float2 hx = __fmul2_rn(make_float2(0.5f, 0.5f), x);
  return __ffma2_rn(make_float2(fabsf(hx.x), fabsf(hx.y)), ea, hx);
}

}  // namespace hedit
