// Microbenchmark: tcgen05.ld / tcgen05.st throughput per SM as a function of the number of warps issuing them.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

template <int MODE>  // 0: ld, 1: st, 2: ld+add+st
__global__ void __launch_bounds__(512, 1) probe(int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = slot + (((uint32_t)(warp & 3) * 32) << 16);
  uint32_t r[32];
  for (int j = 0; j < 32; ++j) r[j] = threadIdx.x + j;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    const uint32_t col = (uint32_t)(((i + (warp >> 2)) & 15) * 32);
    if (MODE == 0 || MODE == 2) {
      ld32(base + col, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= r[j];
    }
    if (MODE == 2) {
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] += 1;
    }
    if (MODE == 1 || MODE == 2) {
      st32(base + col, r);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot));
}

int main() {
  long long* out;
  uint32_t* sink;
  cudaMalloc(&out, 8 * 148);
  cudaMalloc(&sink, 4);
  const int iters = 2048;
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {1, 4, 8, 16}) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) probe<0><<<1, warps * 32>>>(iters, out, sink);
        if (mode == 1) probe<1><<<1, warps * 32>>>(iters, out, sink);
        if (mode == 2) probe<2><<<1, warps * 32>>>(iters, out, sink);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
      }
      const double bytes = (double)iters * warps * 4096.0 * (mode == 2 ? 2 : 1);
      printf("mode %d (%s) warps %2d: %lld clks, %.1f clk per instr per warp, %.1f B/clk/SM  [%s]\n", mode,
             mode == 0 ? "ld" : mode == 1 ? "st" : "ld+st", warps, h, (double)h / iters, bytes / h,
             cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
