// PROBE (not part of the library): issue rate of tcgen05.mma.kind::tf32 on one SM -- cycles per MMA for
// M = 128, N in {64, 128, 256}, K = 8, with the A operand from shared memory (SS) or from tensor memory (TS).
// Operands are whatever bits sit in shared memory / TMEM (timing only).  One thread issues `reps` MMAs back to back
// (descriptors advance through a 64 KB window), commits, and waits for the commit: cycles = (t_done - t_start) / reps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/probes/mma_rate tools/probes/mma_rate.cu
//   tools/probes/mma_rate            # prints a small table, for grid = 1 and grid = 148
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "../../multimodal_eeg_fmri_b200/csrc/xm_common.cuh"
#include "../../multimodal_eeg_fmri_b200/csrc/xm_ptx.cuh"

namespace xm {
int g_last_cuda_error = 0;
}
using namespace xm;

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
      "r"(a), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) rate_kernel(int n, int ts, int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f + (i & 7);
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    ptx::tmem_alloc(&slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::make_idesc_tf32(128, n, 0, 0);
    const uint64_t da0 = ptx::make_smem_desc(ptx::smem_u32(smem), 16, 1024, 2);
    const uint64_t db0 = ptx::make_smem_desc(ptx::smem_u32(smem + 64 * 1024), 16, 1024, 2);
    for (int round = 0; round < 2; ++round) {  // round 0 warms up
      const long long t0 = clock64();
      for (int i = 0; i < reps; ++i) {
        const uint64_t off = (uint64_t)(((i >> 2) & 3) * 1024 + (i & 3) * 2);  // 4 k-blocks x 4 K=8 slabs
        if (ts) mma_ts(tmem + 256, tmem + (uint32_t)((i & 15) * 8), db0 + off, idesc, i ? 1u : 0u);
        else ptx::mma_tf32_ss(tmem + 256, da0 + off, db0 + off, idesc, i ? 1u : 0u);
      }
      const long long t1 = clock64();
      ptx::mma_commit(&bar);
      ptx::mbar_wait(&bar, (uint32_t)round & 1u);
      const long long t2 = clock64();
      if (round == 1 && blockIdx.x == 0) {
        out[0] = t1 - t0;
        out[1] = t2 - t0;
      }
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem, 512);
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  const int smem = 161 * 1024 + 1024;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 2048;
  printf("%6s %4s %5s %14s %14s %12s\n", "grid", "N", "A", "issue clk/MMA", "total clk/MMA", "TFLOP/s@148");
  for (int grid : {1, 148})
    for (int ts = 0; ts < 2; ++ts)
      for (int n : {64, 128, 256}) {
        if (ts && n > 256) continue;
        rate_kernel<<<grid, 128, smem>>>(n, ts, reps, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) return printf("error %s\n", cudaGetErrorString(e)), 1;
        long long h[2];
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        const double per = (double)h[1] / reps;
        printf("%6d %4d %5s %14.1f %14.1f %12.1f\n", grid, n, ts ? "TMEM" : "smem", (double)h[0] / reps, per,
               2.0 * 128 * n * 8 / per * 148 * 1.9e9 / 1e12);
      }
  return 0;
}
