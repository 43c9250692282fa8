// Micro-benchmarks that settle two sm_100a questions the attention kernel's design depends on (B200, run under gpurun):
//   1. TMEM read bandwidth per SM as a function of the number of reading warps (tcgen05.ld 32x32b.x32 in a loop);
//   2. MUFU throughput of ex2.approx.ftz.f32 vs ex2.approx.f16x2 (the latter compiles to TWO MUFU.EX2.F16, one per half), and of a
//      packed-FMA polynomial exp2 (FFMA2) running next to the MUFU stream.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu ; prints bytes/clk/SM and ops/clk/SM.
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__global__ void tmem_read_kernel(int iters, unsigned long long* clk_out, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + (uint32_t((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  const unsigned long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t v[32];
    const uint32_t col = uint32_t((i * 32 + (warp >> 2) * 64) & 511) & ~31u;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(base + col)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    acc += __uint_as_float(v[0] ^ v[7] ^ v[13] ^ v[22] ^ v[31]);      // static indices: the values stay in registers
  }
  __syncthreads();
  const unsigned long long t1 = clock64();
  if (threadIdx.x == 0) clk_out[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

// mode 0: ex2.approx.ftz.f32 ; 1: ex2.approx.f16x2 ; 2: f32 MUFU + an equal number of polynomial exp2 pairs on the FMA pipe
template <int MODE>
__global__ void mufu_kernel(int iters, unsigned long long* clk_out, float* sink, float seed) {
  float x[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = seed * (k + 1) - 3.f - threadIdx.x * 1e-3f;
  float2 q[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) q[k] = make_float2(x[k], x[k + 4]);
  __syncthreads();
  const unsigned long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0 || MODE == 2) {
#pragma unroll
      for (int k = 0; k < 8; ++k) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[k])); x[k] = y - 1.5f; }
    }
    if (MODE == 1) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        __half2 h = __floats2half2_rn(x[2 * k], x[2 * k + 1]);
        uint32_t u = *reinterpret_cast<uint32_t*>(&h), y;
        asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(u));
        const float2 f = __half22float2(*reinterpret_cast<__half2*>(&y));
        x[2 * k] = f.x - 1.5f; x[2 * k + 1] = f.y - 1.5f;
      }
    }
    if (MODE == 2) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {          // 4 pairs = 8 polynomial exp2 per 8 MUFU exp2
        const float2 xc = q[k];
        const float2 t = __fadd2_rn(xc, make_float2(12582912.f, 12582912.f));
        const float2 j = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
        const float2 f = __ffma2_rn(j, make_float2(-1.f, -1.f), xc);
        float2 p = __ffma2_rn(f, make_float2(0.0551716648f, 0.0551716648f), make_float2(0.2426111251f, 0.2426111251f));
        p = __ffma2_rn(p, f, make_float2(0.6932609677f, 0.6932609677f));
        p = __ffma2_rn(p, f, make_float2(0.9999280572f, 0.9999280572f));
        q[k] = make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23)) - 1.5f,
                           __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23)) - 1.5f);
      }
    }
  }
  __syncthreads();
  const unsigned long long t1 = clock64();
  if (threadIdx.x == 0) clk_out[blockIdx.x] = t1 - t0;
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += x[k];
#pragma unroll
  for (int k = 0; k < 4; ++k) s += q[k].x + q[k].y;
  if (s == 123.456f) sink[0] = s;
}

static double median_clk(unsigned long long* d, int n) {
  unsigned long long h[1024];
  cudaMemcpy(h, d, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  for (int i = 0; i < n; ++i) for (int j = i + 1; j < n; ++j) if (h[j] < h[i]) { unsigned long long t = h[i]; h[i] = h[j]; h[j] = t; }
  return double(h[n / 2]);
}

int main() {
  unsigned long long* clk; float* sink;
  cudaMalloc(&clk, 1024 * sizeof(unsigned long long)); cudaMalloc(&sink, 16);
  const int nsm = 148, iters = 4096;
  for (int warps : {4, 8, 16}) {
    tmem_read_kernel<<<nsm, warps * 32>>>(iters, clk, sink);
    tmem_read_kernel<<<nsm, warps * 32>>>(iters, clk, sink);
    cudaError_t e = cudaDeviceSynchronize();
    const double c = median_clk(clk, nsm);
    printf("tmem_read warps=%2d: %.1f bytes/clk/SM (%.0f clk per 4 KB warp-load)  [%s]\n", warps, double(warps) * iters * 4096.0 / c, c / iters,
           cudaGetErrorString(e));
  }
  for (int warps : {4, 8, 16}) {
    mufu_kernel<0><<<nsm, warps * 32>>>(iters, clk, sink, 0.37f); cudaDeviceSynchronize();
    const double c0 = median_clk(clk, nsm);
    mufu_kernel<1><<<nsm, warps * 32>>>(iters, clk, sink, 0.37f); cudaDeviceSynchronize();
    const double c1 = median_clk(clk, nsm);
    mufu_kernel<2><<<nsm, warps * 32>>>(iters, clk, sink, 0.37f); cudaDeviceSynchronize();
    const double c2 = median_clk(clk, nsm);
    const double n = double(warps) * 32 * iters * 8;
    printf("exp2 warps=%2d: f32 MUFU %.2f/clk/SM | f16x2 %.2f/clk/SM | f32 MUFU + as many FFMA2-polynomial: %.2f/clk/SM total (%.0f vs %.0f clk)  [%s]\n", warps,
           n / c0, n / c1, 2 * n / c2, c2, c0, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
