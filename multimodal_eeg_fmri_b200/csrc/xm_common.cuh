// Shared helpers for the xmodal-b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/xmodal_b200.h"

#ifndef XM_DEVICE
#define XM_DEVICE __device__ __forceinline__
#endif

namespace xm {

// Last CUDA runtime error seen by a launcher in this process (diagnostic only).
extern int g_last_cuda_error;

inline int check_launch() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    g_last_cuda_error = (int)e;
    return XM_ERR_LAUNCH;
  }
  return XM_OK;
}

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

constexpr int kNumSMs = 148;

XM_DEVICE float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
XM_DEVICE double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
XM_DEVICE float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Round-to-nearest fp32 -> tf32 (kept in an fp32 container). Tensor cores read the top 19
// bits only; rounding at the producer removes the truncation bias from the next contraction.
XM_DEVICE float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// erf-GELU (nn.GELU default) and its derivative from ONE exponential: the standard normal pdf phi(x) is needed by
// the derivative anyway, and Phi(x) = 1 - phi(x) * (b1 t + ... + b5 t^5), t = 1 / (1 + 0.2316419 |x|)
// (Abramowitz & Stegun 26.2.17, |error| < 7.5e-8; measured 2.9e-7 in fp32 with the approximate ex2 / rcp units,
// 8e-8 relative L2 on GELU and 9e-8 on GELU' over N(0, 1.5) inputs -- the same level as erff()).  About 16
// instructions instead of ~35 for erff + expf: the streaming kernels around GELU are instruction-issue bound.
XM_DEVICE float approx_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
XM_DEVICE float approx_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
XM_DEVICE void normal_cdf_pdf(float x, float& cdf, float& pdf) {
  pdf = 0.3989422804f * approx_ex2(-0.72134752f * x * x);  // exp(-x^2 / 2) / sqrt(2 pi)
  const float t = approx_rcp(fmaf(0.2316419f, fabsf(x), 1.0f));
  const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.330274429f, -1.821255978f), 1.781477937f), -0.356563782f), 0.319381530f);
  const float q = pdf * poly;  // 1 - Phi(|x|)
  cdf = x >= 0.0f ? 1.0f - q : q;
}
XM_DEVICE float gelu_erf(float x) {
  float cdf, pdf;
  normal_cdf_pdf(x, cdf, pdf);
  return x * cdf;
}
XM_DEVICE float gelu_erf_grad(float x) {
  float cdf, pdf;
  normal_cdf_pdf(x, cdf, pdf);
  return fmaf(x, pdf, cdf);
}

// Counter-based dropout mask: one 32-bit hash per element (three multiply / xor-shift rounds over a
// seed-keyed index; ~10 integer instructions -- the streaming kernels that apply dropout next to an erf
// GELU are instruction-bound, a 64-bit mixer doubled their integer work).
// keep  <=>  hash >= threshold, threshold = p * 2^32.
//
// Seed epoch (CUDA-graph replay): a captured launch carries its 64-bit seed as a frozen kernel argument, so every hash
// folds in a per-translation-unit constant that xm_seed_epoch_advance / xm_seed_epoch_set refresh from ONE device
// counter (a kernel node + 8-byte copy nodes inside the graph): replay k draws the masks of (seed, epoch k), forward
// and backward alike.  The epoch is 0 unless those entry points are used, which leaves every seed as passed.
static __constant__ unsigned long long xm_seed_epoch_c = 0ull;
XM_DEVICE uint64_t epoch_seed(uint64_t seed) { return seed + xm_seed_epoch_c * 0xD1B54A32D192ED03ull; }
// Defines this translation unit's accessor for the library-wide refresh (csrc/optimizer.cu).
#define XM_DEFINE_SEED_EPOCH_SLOT(tu)                                                     \
  extern "C" __attribute__((visibility("hidden"))) void* xm_seed_epoch_slot_##tu() {     \
    void* p = nullptr;                                                                    \
    return cudaGetSymbolAddress(&p, xm::xm_seed_epoch_c) == cudaSuccess ? p : nullptr;    \
  }

XM_DEVICE uint32_t hash_u32(uint64_t idx, uint64_t seed) {
  seed = epoch_seed(seed);
  uint32_t x = (uint32_t)idx ^ ((uint32_t)(idx >> 32) * 0x9E3779B1u) ^ (uint32_t)seed;
  x *= 0x85EBCA6Bu;
  x ^= x >> 13;
  x += (uint32_t)(seed >> 32);
  x *= 0xC2B2AE35u;
  x ^= x >> 16;
  x *= 0x27D4EB2Fu;
  x ^= x >> 15;
  return x;
}
XM_DEVICE bool dropout_keep(uint64_t idx, uint64_t seed, uint32_t threshold) {
  return hash_u32(idx, seed) >= threshold;
}

// Column sums over the warp's 32 rows of a 32-column block held one row per lane: butterfly reduce-scatter,
// 31 shuffles; lane j returns the sum of column j.
XM_DEVICE float warp_column_sums(const uint32_t (&r)[32], int lane) {
  float v[16];
  {
    const bool up = lane & 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float keep = __uint_as_float(up ? r[16 + i] : r[i]);
      const float send = __uint_as_float(up ? r[i] : r[16 + i]);
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
#pragma unroll
  for (int s = 8; s >= 1; s >>= 1) {
    const bool up = lane & s;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float keep = up ? v[s + i] : v[i];
      const float send = up ? v[i] : v[s + i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];  // column index = lane (bit b of the lane selected the upper half at step b)
}

// act codes shared with the host (include/xmodal_b200.h)
XM_DEVICE float apply_act(float x, int act) {
  if (act == XM_ACT_RELU) return fmaxf(x, 0.0f);
  if (act == XM_ACT_GELU) return gelu_erf(x);
  if (act == XM_ACT_TANH) return tanhf(x);
  if (act == XM_ACT_SIGMOID) return 1.0f / (1.0f + expf(-x));
  return x;
}
XM_DEVICE float act_grad(float x, int act) {
  if (act == XM_ACT_RELU) return x > 0.0f ? 1.0f : 0.0f;
  if (act == XM_ACT_GELU) return gelu_erf_grad(x);
  if (act == XM_ACT_TANH) {
    const float t = tanhf(x);
    return 1.0f - t * t;
  }
  if (act == XM_ACT_SIGMOID) {
    const float s = 1.0f / (1.0f + expf(-x));
    return s * (1.0f - s);
  }
  return 1.0f;
}

}  // namespace xm
