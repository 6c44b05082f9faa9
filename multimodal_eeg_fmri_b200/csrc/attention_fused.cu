// Fused multi-head self-attention core for sm_100a (nn.MultiheadAttention inside TemporalTransformerBlock,
// EEG_CODE/enhanced_models_v4.py:71-73,98): the probability matrix never leaves the SM.
//
//   forward   (queries on the TMEM lanes):  S = q k^T (tcgen05, one 128 x 256 accumulator holds complete rows)
//             -> row softmax + dropout by 16 epilogue warps -> P~ written BACK INTO TMEM (tcgen05.st) over S
//             -> O = P~ v with P~ as the TMEM-resident A operand of the second MMA -> TMA store.
//   backward, dq (queries on the lanes), per 128-key half:  S = q k^T and dP~ = dO v^T side by side in TMEM
//             -> dS = scale * P o (mask/keep * dP~ - delta) over S in TMEM -> dq += dS k (A from TMEM).
//   backward, dk / dv (KEYS on the lanes), per 128-query half:  S^T = k q^T and dP~^T = v dO^T
//             -> P~^T and dS^T in place -> dv += P~^T dO, dk += dS^T q (A from TMEM).
//   delta_i = dO_i . O_i (the row sum of P~ o dP~) is formed by the dq kernel and handed to the dk/dv kernel.
//
// HBM traffic per (sample, head): q, k, v, O, dO read a few times (L2) and dq, dk, dv written once -- the
// L x L matrices P~ and dS (4.2 GB each per layer at batch 4096) are never stored.  Supported: dh == 32, L <= 256.
//
// Dropout mask (a pure function of seed, slab, query, key, identical in all three kernels and in
// xm_attn_fused_mask_u8): per (query row r = slab*L + query, 32-key chunk c) a stream seed
// h0 = mix(rs(r) + (c + 1) * 0x9E3779B1), rs(r) = high word of hash_u64(r, seed), mix = two multiply / xorshift
// rounds; key c*32 + j takes LCG state h_{j+1} (h' = h * 747796405 + 2891336453) and is kept iff h_{j+1} >= p * 2^32.
// Along a query row that is one multiply-add per element; the kernels with keys on the lanes jump straight to
// state j + 1 with per-lane constants (A^(j+1), C * (A^(j+1) - 1) / (A - 1)).
//
// Operands written back to TMEM are NOT rounded to tf32 (the tensor core truncates them): their common scale
// factor carries (1 + 0.7213 * 2^-11), the mean relative truncation loss over a binade, so the truncation is
// zero-mean like round-to-nearest -- the two integer instructions per element of an explicit rounding were a
// quarter of the epilogue's ALU-pipe work, and the epilogues are instruction-issue bound.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/xmodal_b200.h"
#include "gemm_engine.cuh"

namespace xm {
namespace fa {

constexpr int kEpiWarps = 16;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr uint32_t kGold = 0x9E3779B1u;
constexpr float kLog2e = 1.4426950408889634f;

struct FaParams {
  int L, H, d;      // sequence length, heads, d = H*32
  int T;            // 128-row tiles per slab = ceil(L / 128)
  int items;        // B*H*T
  float alpha;      // softmax scale
  uint32_t thr;     // keep <=> x >= thr (0: no dropout)
  float dscale;     // 1 / (1 - p)
  unsigned long long seed;
  int round_out;
  float* lse;        // (B*H, L) natural-log logsumexp of the scaled scores
  float* delta;      // (B*H, L)
  const float* out;  // (B, L, d) forward output (backward: delta = dO . O)
  const float* dout; // (B, L, d)
};

XM_DEVICE uint32_t row_seed(unsigned long long row_id, unsigned long long seed) {
  return (uint32_t)(hash_u64(row_id, seed) >> 32);
}
XM_DEVICE uint32_t mask_mix(uint32_t x) {
  x *= 0x7FEB352Du;
  x ^= x >> 15;
  x *= 0x846CA68Bu;
  return x ^ (x >> 16);
}
XM_DEVICE uint32_t chunk_seed(uint32_t rs, int chunk) { return mask_mix(rs + (uint32_t)(chunk + 1) * kGold); }
constexpr uint32_t kLcgA = 747796405u, kLcgC = 2891336453u;
constexpr float kTruncComp = 1.0f + 0.7213f / 2048.0f;

// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (128 lanes x K columns, tf32 bit patterns) is read from TMEM.
XM_DEVICE void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

XM_DEVICE void epi_barrier_all() { asm volatile("bar.sync 5, 512;" ::: "memory"); }

// 32 x 32 fp32 box of this warp -> swizzled staging buffer -> TMA store (rows / columns outside the tensor clipped)
XM_DEVICE void store_box(const CUtensorMap* tm, uint8_t* sb, int lane, const uint32_t (&r)[32], bool round, int col, int row0,
                         int z) {
  if (lane == 0) ptx::bulk_wait_read<0>();  // the previous store has finished reading this buffer
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                           __uint_as_float(r[4 * j + 3]));
    if (round) v = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
    *reinterpret_cast<float4*>(sb + lane * 128 + ((j ^ (lane & 7)) << 4)) = v;
  }
  ptx::fence_proxy_async();
  __syncwarp();
  if (lane == 0) {
    ptx::tma_store_3d(tm, sb, col, row0, z);
    ptx::bulk_commit();
  }
}

struct Bars {
  uint64_t full[2], empty[2];      // per-item operand buffers (fwd: q,k,v; bwd: the lane-side tiles)
  uint64_t hfull[2], hempty[2];    // per-half operand ring (bwd)
  uint64_t s_full, p_ready;        // score accumulators complete / epilogue wrote the TMEM A operands
  uint64_t o_full[2], o_empty[2];  // output accumulators
};

XM_DEVICE void init_common(Bars& bar, uint32_t* tmem_slot, int warp, uint32_t cols) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bar.full[i], 1);
      ptx::mbar_init(&bar.empty[i], 1);
      ptx::mbar_init(&bar.hfull[i], 1);
      ptx::mbar_init(&bar.hempty[i], 1);
      ptx::mbar_init(&bar.o_full[i], 1);
      ptx::mbar_init(&bar.o_empty[i], 4);  // the four part-0 warps drain an output accumulator
    }
    ptx::mbar_init(&bar.s_full, 1);
    ptx::mbar_init(&bar.p_ready, kEpiWarps);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
}

// ============================================================================ forward
// smem: 2 stages of { q tile 16 KB | k 32 KB | v (MN-major) 32 KB }, then 4 staging boxes.
constexpr int kFwdStage = 16384 + 32768 + 32768;
constexpr int kFwdSmem = 2 * kFwdStage + 4 * 4096 + 1024;

__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmVmn, const __grid_constant__ CUtensorMap tmO, const FaParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ Bars bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float red[4][4][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* staging = smem + 2 * kFwdStage;
  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmQ);
    ptx::prefetch_tensormap(&tmK);
    ptx::prefetch_tensormap(&tmVmn);
    ptx::prefetch_tensormap(&tmO);
  }
  init_common(bar, &tmem_slot, warp, 512);
  const uint32_t tmem = tmem_slot;
  const uint32_t tS = tmem;            // columns [0, 256): S, then P~
  const uint32_t tO = tmem + 256;      // two output accumulators of 32 columns
  const int ksteps = (p.L + 7) >> 3;   // K = 8 slabs of keys with any valid key

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int w = blockIdx.x; w < p.items; w += gridDim.x, ++it) {
        const int s = it & 1;
        ptx::mbar_wait(&bar.empty[s], (((uint32_t)it >> 1) & 1u) ^ 1u);
        const int z = w / p.T, qt = w - z * p.T, b = z / p.H, h = z - b * p.H;
        uint8_t* st = smem + s * kFwdStage;
        ptx::mbar_arrive_expect_tx(&bar.full[s], kFwdStage);
        ptx::tma_load_3d(&tmQ, &bar.full[s], st, h * 32, qt * 128, b);
        ptx::tma_load_3d(&tmK, &bar.full[s], st + 16384, p.d + h * 32, 0, b);
        ptx::tma_load_3d(&tmVmn, &bar.full[s], st + 49152, 2 * p.d + h * 32, 0, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = ptx::make_idesc_tf32(128, 256, 0, 0);
      const uint32_t idesc_o = ptx::make_idesc_tf32(128, 32, 0, 1);
      int it = 0;
      for (int w = blockIdx.x; w < p.items; w += gridDim.x, ++it) {
        const int s = it & 1;
        const uint32_t ph = ((uint32_t)it >> 1) & 1u;
        ptx::mbar_wait(&bar.full[s], ph);
        ptx::tc_fence_after_sync();
        const uint32_t sa = ptx::smem_u32(smem + s * kFwdStage);
        const uint64_t dq = ptx::make_smem_desc(sa, 16, 1024, 2);
        const uint64_t dk = ptx::make_smem_desc(sa + 16384, 16, 1024, 2);
        const uint64_t dv = ptx::make_smem_desc(sa + 49152, 4096, 512, 1);
#pragma unroll
        for (int k8 = 0; k8 < 4; ++k8) ptx::mma_tf32_ss(tS, dq + (uint64_t)(k8 * 2), dk + (uint64_t)(k8 * 2), idesc_s, k8 > 0);
        ptx::mma_commit(&bar.s_full);
        ptx::mbar_wait(&bar.p_ready, (uint32_t)it & 1u);
        ptx::mbar_wait(&bar.o_empty[s], ph ^ 1u);
        ptx::tc_fence_after_sync();
        for (int k8 = 0; k8 < ksteps; ++k8)
          mma_tf32_ts(tO + (uint32_t)(s * 32), tS + (uint32_t)(k8 * 8), dv + (uint64_t)(k8 * 64), idesc_o, k8 > 0);
        ptx::mma_commit(&bar.o_full[s]);
        ptx::mma_commit(&bar.empty[s]);
      }
    }
  } else {
    const int q = warp & 3, part = (warp - 2) >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const float c = p.alpha * kLog2e;
    int it = 0;
    for (int w = blockIdx.x; w < p.items; w += gridDim.x, ++it) {
      const int z = w / p.T, qt = w - z * p.T, b = z / p.H, h = z - b * p.H;
      const int m = qt * 128 + q * 32 + lane;
      const bool live = m < p.L;
      const unsigned long long row_id = (unsigned long long)z * (unsigned long long)p.L + (unsigned long long)m;
      ptx::mbar_wait(&bar.s_full, (uint32_t)it & 1u);
      ptx::tc_fence_after_sync();
      const uint32_t acc = tS + lane_base;
      const int ch0 = part * 2;
      float mx = -3.0e38f;
      for (int ch = ch0; ch < ch0 + 2; ++ch) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(acc + (uint32_t)(ch * 32), r);
        ptx::tmem_ld_wait();
        const int nv = p.L - ch * 32;
        if (nv >= 32) {
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nv) mx = fmaxf(mx, __uint_as_float(r[j]));
        }
      }
      red[q][part][lane] = mx;
      quad_barrier(q);
      mx = fmaxf(fmaxf(red[q][0][lane], red[q][1][lane]), fmaxf(red[q][2][lane], red[q][3][lane]));
      const float mc = mx * c;
      float sum = 0.f;
      for (int ch = ch0; ch < ch0 + 2; ++ch) {  // e = exp2(s*c - max) once per element, kept in TMEM over s
        uint32_t r[32];
        ptx::tmem_ld_32x32(acc + (uint32_t)(ch * 32), r);
        ptx::tmem_ld_wait();
        const int nv = p.L - ch * 32;
        if (nv >= 32) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float e = fast_exp2(fmaf(__uint_as_float(r[j]), c, -mc));
            sum += e;
            r[j] = __float_as_uint(e);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float e = (j < nv) ? fast_exp2(fmaf(__uint_as_float(r[j]), c, -mc)) : 0.f;
            sum += e;
            r[j] = __float_as_uint(e);
          }
        }
        ptx::tmem_st_32x32(acc + (uint32_t)(ch * 32), r);
      }
      ptx::tmem_st_wait();
      quad_barrier(q);
      red[q][part][lane] = sum;
      quad_barrier(q);
      sum = (red[q][0][lane] + red[q][1][lane]) + (red[q][2][lane] + red[q][3][lane]);
      if (live && part == 0) p.lse[row_id] = (mc + log2f(sum)) * 0.6931471805599453f;
      const float keep_mul = p.dscale * kTruncComp / sum;  // normalisation, inverted dropout, truncation compensation
      const uint32_t rs = row_seed(row_id, p.seed);
      for (int ch = ch0; ch < ch0 + 2; ++ch) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(acc + (uint32_t)(ch * 32), r);
        uint32_t hsd = chunk_seed(rs, ch);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          hsd = hsd * kLcgA + kLcgC;
          r[j] = __float_as_uint(__uint_as_float(r[j]) * ((hsd >= p.thr) ? keep_mul : 0.f));
        }
        ptx::tmem_st_32x32(acc + (uint32_t)(ch * 32), r);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bar.p_ready);
      quad_barrier(q);  // the exchange buffer may be rewritten by the next item
      if (part == 0) {
        const int s = it & 1;
        ptx::mbar_wait(&bar.o_full[s], ((uint32_t)it >> 1) & 1u);
        ptx::tc_fence_after_sync();
        uint32_t r[32];
        ptx::tmem_ld_32x32(tO + (uint32_t)(s * 32) + lane_base, r);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bar.o_empty[s]);
        store_box(&tmO, staging + q * 4096, lane, r, p.round_out != 0, h * 32, qt * 128 + q * 32, b);
      }
    }
    if (part == 0 && lane == 0) ptx::bulk_wait_all();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem, 512);
  }
}

// ============================================================================ backward
// One template for both backward kernels.  KV == false: lanes = queries of tile `t`, halves walk the keys, output
// dq.  KV == true: lanes = keys of tile `t`, halves walk the queries, outputs dk and dv.
//   per item   (2 buffers): two lane-side tiles X0, X1 of 16 KB, K-major   (dq: q, dO ;  dk/dv: k, v)
//   per half   (2 stages) : Y0, Y1 K-major 16 KB each                      (dq: k, v  ;  dk/dv: q, dO)
//                           + MN-major copies for the output products      (dq: k     ;  dk/dv: dO, q)
//   TMEM: [0,128) S (then dS | P~^T)   [128,256) dP~ (then dS^T)   [256, ...) output accumulators (2 sets)
constexpr int kItemBytes = 2 * 16384;
template <bool KV>
struct BwdCfg {
  static constexpr int kHalfBytes = (KV ? 4 : 3) * 16384;
  static constexpr int kStagingOff = 2 * kItemBytes + 2 * kHalfBytes;
  static constexpr int kTabOff = kStagingOff + 4 * 4096;       // KV: lse / delta tables of 256 queries + the mask
  static constexpr int kSmem = kTabOff + (KV ? 6 * 1024 : 0) + 1024;  // stream seeds of [4 key chunks][256 queries]
  static constexpr int kOutCols = KV ? 64 : 32;                 // per accumulator set
};

template <bool KV>
__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmQKVmn,
                const __grid_constant__ CUtensorMap tmDO, const __grid_constant__ CUtensorMap tmDOmn,
                const __grid_constant__ CUtensorMap tmDQKV, const FaParams p) {
  using C = BwdCfg<KV>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ Bars bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float red[4][4][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* half0 = smem + 2 * kItemBytes;
  uint8_t* staging = smem + C::kStagingOff;
  float* tab_off = reinterpret_cast<float*>(smem + C::kTabOff);  // lse in the log2 domain
  float* tab_delta = tab_off + 256;
  uint32_t* tab_seed = reinterpret_cast<uint32_t*>(tab_delta + 256);  // [quadrant = key chunk of the tile][query]
  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmQKV);
    ptx::prefetch_tensormap(&tmQKVmn);
    ptx::prefetch_tensormap(&tmDO);
    ptx::prefetch_tensormap(&tmDOmn);
    ptx::prefetch_tensormap(&tmDQKV);
  }
  init_common(bar, &tmem_slot, warp, 512);
  const uint32_t tmem = tmem_slot;
  const uint32_t tS = tmem, tP = tmem + 128, tOut = tmem + 256;
  const int NH = p.T;  // halves of the other index (same tiling: ceil(L / 128))

  if (warp == 0) {
    if (lane == 0) {
      int it = 0, hs = 0;
      for (int w = blockIdx.x; w < p.items; w += gridDim.x, ++it) {
        const int s = it & 1;
        const int z = w / p.T, t = w - z * p.T, b = z / p.H, h = z - b * p.H;
        ptx::mbar_wait(&bar.empty[s], (((uint32_t)it >> 1) & 1u) ^ 1u);
        uint8_t* xi = smem + s * kItemBytes;
        ptx::mbar_arrive_expect_tx(&bar.full[s], kItemBytes);
        if (KV) {
          ptx::tma_load_3d(&tmQKV, &bar.full[s], xi, p.d + h * 32, t * 128, b);              // k tile
          ptx::tma_load_3d(&tmQKV, &bar.full[s], xi + 16384, 2 * p.d + h * 32, t * 128, b);  // v tile
        } else {
          ptx::tma_load_3d(&tmQKV, &bar.full[s], xi, h * 32, t * 128, b);       // q tile
          ptx::tma_load_3d(&tmDO, &bar.full[s], xi + 16384, h * 32, t * 128, b);  // dO tile
        }
        for (int hf = 0; hf < NH; ++hf, ++hs) {
          const int st = hs & 1;
          ptx::mbar_wait(&bar.hempty[st], (((uint32_t)hs >> 1) & 1u) ^ 1u);
          uint8_t* yi = half0 + st * C::kHalfBytes;
          ptx::mbar_arrive_expect_tx(&bar.hfull[st], C::kHalfBytes);
          if (KV) {
            ptx::tma_load_3d(&tmQKV, &bar.hfull[st], yi, h * 32, hf * 128, b);            // q half (K-major)
            ptx::tma_load_3d(&tmDO, &bar.hfull[st], yi + 16384, h * 32, hf * 128, b);     // dO half (K-major)
            ptx::tma_load_3d(&tmDOmn, &bar.hfull[st], yi + 32768, h * 32, hf * 128, b);   // dO half (MN-major) -> dv
            ptx::tma_load_3d(&tmQKVmn, &bar.hfull[st], yi + 49152, h * 32, hf * 128, b);  // q half (MN-major) -> dk
          } else {
            ptx::tma_load_3d(&tmQKV, &bar.hfull[st], yi, p.d + h * 32, hf * 128, b);            // k half
            ptx::tma_load_3d(&tmQKV, &bar.hfull[st], yi + 16384, 2 * p.d + h * 32, hf * 128, b);  // v half
            ptx::tma_load_3d(&tmQKVmn, &bar.hfull[st], yi + 32768, p.d + h * 32, hf * 128, b);  // k half (MN) -> dq
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = ptx::make_idesc_tf32(128, 128, 0, 0);
      const uint32_t idesc_o = ptx::make_idesc_tf32(128, 32, 0, 1);
      int it = 0, hs = 0;
      for (int w = blockIdx.x; w < p.items; w += gridDim.x, ++it) {
        const int s = it & 1;
        const uint32_t ph = ((uint32_t)it >> 1) & 1u;
        ptx::mbar_wait(&bar.full[s], ph);
        const uint32_t xa = ptx::smem_u32(smem + s * kItemBytes);
        const uint64_t dx0 = ptx::make_smem_desc(xa, 16, 1024, 2);
        const uint64_t dx1 = ptx::make_smem_desc(xa + 16384, 16, 1024, 2);
        const uint32_t out = tOut + (uint32_t)(s * C::kOutCols);
        for (int hf = 0; hf < NH; ++hf, ++hs) {
          const int st = hs & 1;
          ptx::mbar_wait(&bar.hfull[st], ((uint32_t)hs >> 1) & 1u);
          ptx::tc_fence_after_sync();
          const uint32_t ya = ptx::smem_u32(half0 + st * C::kHalfBytes);
          const uint64_t dy0 = ptx::make_smem_desc(ya, 16, 1024, 2);
          const uint64_t dy1 = ptx::make_smem_desc(ya + 16384, 16, 1024, 2);
#pragma unroll
          for (int k8 = 0; k8 < 4; ++k8)  // S (or S^T): lane-side tile 0 times half tile 0
            ptx::mma_tf32_ss(tS, dx0 + (uint64_t)(k8 * 2), dy0 + (uint64_t)(k8 * 2), idesc_s, k8 > 0);
#pragma unroll
          for (int k8 = 0; k8 < 4; ++k8)  // dP~ (or its transpose): lane-side tile 1 times half tile 1
            ptx::mma_tf32_ss(tP, dx1 + (uint64_t)(k8 * 2), dy1 + (uint64_t)(k8 * 2), idesc_s, k8 > 0);
          ptx::mma_commit(&bar.s_full);
          ptx::mbar_wait(&bar.p_ready, (uint32_t)hs & 1u);
          if (hf == 0) ptx::mbar_wait(&bar.o_empty[s], ph ^ 1u);
          ptx::tc_fence_after_sync();
          const uint64_t dm0 = ptx::make_smem_desc(ya + 32768, 4096, 512, 1);
          if (KV) {
            const uint64_t dm1 = ptx::make_smem_desc(ya + 49152, 4096, 512, 1);
#pragma unroll 4
            for (int k8 = 0; k8 < 16; ++k8)  // dv += P~^T dO
              mma_tf32_ts(out + 32, tS + (uint32_t)(k8 * 8), dm0 + (uint64_t)(k8 * 64), idesc_o, (hf > 0 || k8 > 0));
#pragma unroll 4
            for (int k8 = 0; k8 < 16; ++k8)  // dk += dS^T q
              mma_tf32_ts(out, tP + (uint32_t)(k8 * 8), dm1 + (uint64_t)(k8 * 64), idesc_o, (hf > 0 || k8 > 0));
          } else {
#pragma unroll 4
            for (int k8 = 0; k8 < 16; ++k8)  // dq += dS k
              mma_tf32_ts(out, tS + (uint32_t)(k8 * 8), dm0 + (uint64_t)(k8 * 64), idesc_o, (hf > 0 || k8 > 0));
          }
          ptx::mma_commit(&bar.hempty[st]);
        }
        ptx::mma_commit(&bar.o_full[s]);
        ptx::mma_commit(&bar.empty[s]);
      }
    }
  } else {
    const int q = warp & 3, part = (warp - 2) >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const float c = p.alpha * kLog2e;
    // ds carries alpha (chain rule of the scaled scores) and the truncation compensation: both are folded into
    // the exponent offset, so pr_a below is alpha' * P
    const float amul = p.alpha * kTruncComp;
    const float log2_amul = log2f(amul);
    const float pt_keep = p.dscale / p.alpha;  // P~^T = pr_a * pt_keep  (kept elements)
    const int et = threadIdx.x - 64;  // 0..511 among the epilogue threads
    // keys on the lanes: LCG jump to state lane + 1 of a chunk's stream
    uint32_t jump_a = 1u, jump_c = 0u;
    if (KV) {
      for (int k = 0; k <= lane; ++k) {
        jump_c = jump_c * kLcgA + kLcgC;
        jump_a *= kLcgA;
      }
    }
    int it = 0, hs = 0;
    for (int w = blockIdx.x; w < p.items; w += gridDim.x, ++it) {
      const int z = w / p.T, t = w - z * p.T, b = z / p.H, h = z - b * p.H;
      const int row = t * 128 + q * 32 + lane;  // query (dq kernel) or key (dk/dv kernel) of this thread
      const bool live = row < p.L;
      const unsigned long long slab0 = (unsigned long long)z * (unsigned long long)p.L;
      float off = 0.f, delta = 0.f;
      uint32_t rs = 0u;
      if (KV) {
        epi_barrier_all();  // every warp is done with the previous item's tables
        {
          const int m = et & 255;
          const bool ok = m < p.L;
          if (et < 256) {
            tab_off[m] = (ok ? __ldg(p.lse + slab0 + m) * kLog2e : 0.f) - log2_amul;
            tab_delta[m] = ok ? __ldg(p.delta + slab0 + m) : 0.f;
          }
          const uint32_t rsm = row_seed(slab0 + (unsigned long long)m, p.seed);
          const int qd0 = (et >> 8) * 2;  // this thread fills two of the tile's four key chunks
          tab_seed[qd0 * 256 + m] = chunk_seed(rsm, t * 4 + qd0);
          tab_seed[(qd0 + 1) * 256 + m] = chunk_seed(rsm, t * 4 + qd0 + 1);
        }
        epi_barrier_all();
      } else {
        // delta = dO . O over the head's 32 columns: each part takes 8 of them
        float part_sum = 0.f;
        if (live) {
          const long long o = ((long long)b * p.L + row) * p.d + h * 32 + part * 8;
          const float4 a0 = __ldg(reinterpret_cast<const float4*>(p.dout + o));
          const float4 a1 = __ldg(reinterpret_cast<const float4*>(p.dout + o + 4));
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.out + o));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.out + o + 4));
          part_sum = a0.x * b0.x + a0.y * b0.y + a0.z * b0.z + a0.w * b0.w + a1.x * b1.x + a1.y * b1.y + a1.z * b1.z +
                     a1.w * b1.w;
          off = __ldg(p.lse + slab0 + row) * kLog2e;
        }
        off -= log2_amul;
        red[q][part][lane] = part_sum;
        quad_barrier(q);
        delta = (red[q][0][lane] + red[q][1][lane]) + (red[q][2][lane] + red[q][3][lane]);
        quad_barrier(q);
        if (live && part == 0) p.delta[slab0 + row] = delta;
        rs = row_seed(slab0 + (unsigned long long)row, p.seed);
      }
      for (int hf = 0; hf < NH; ++hf, ++hs) {
        ptx::mbar_wait(&bar.s_full, (uint32_t)hs & 1u);
        ptx::tc_fence_after_sync();
        uint32_t r[32], g[32];
        ptx::tmem_ld_32x32(tS + lane_base + (uint32_t)(part * 32), r);
        ptx::tmem_ld_32x32(tP + lane_base + (uint32_t)(part * 32), g);
        const int o0 = hf * 128 + part * 32;  // first key (dq kernel) / query (dk/dv kernel) of this chunk
        if (KV) {
          const uint32_t* sd = tab_seed + q * 256 + o0;
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {  // per-query constants: one broadcast 128-bit load per 4 columns
            const float4 o4 = *reinterpret_cast<const float4*>(tab_off + o0 + 4 * j4);
            const float4 d4 = *reinterpret_cast<const float4*>(tab_delta + o0 + 4 * j4);
            const uint4 s4 = *reinterpret_cast<const uint4*>(sd + 4 * j4);
            const float of[4] = {o4.x, o4.y, o4.z, o4.w}, de[4] = {d4.x, d4.y, d4.z, d4.w};
            const uint32_t sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = 4 * j4 + e;
              const float pr_a = fast_exp2(fmaf(__uint_as_float(r[j]), c, -of[e]));
              const bool keep = sv[e] * jump_a + jump_c >= p.thr;
              const float ds = pr_a * fmaf(keep ? p.dscale : 0.f, __uint_as_float(g[j]), -de[e]);
              r[j] = __float_as_uint(pr_a * (keep ? pt_keep : 0.f));
              g[j] = __float_as_uint(ds);
            }
          }
          ptx::tmem_st_32x32(tS + lane_base + (uint32_t)(part * 32), r);
          ptx::tmem_st_32x32(tP + lane_base + (uint32_t)(part * 32), g);
        } else {
          uint32_t hsd = chunk_seed(rs, o0 >> 5);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            hsd = hsd * kLcgA + kLcgC;
            const float pr_a = fast_exp2(fmaf(__uint_as_float(r[j]), c, -off));
            const float mk = (hsd >= p.thr) ? p.dscale : 0.f;
            r[j] = __float_as_uint(pr_a * fmaf(mk, __uint_as_float(g[j]), -delta));
          }
          ptx::tmem_st_32x32(tS + lane_base + (uint32_t)(part * 32), r);
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bar.p_ready);
      }
      if (part == 0) {
        const int s = it & 1;
        ptx::mbar_wait(&bar.o_full[s], ((uint32_t)it >> 1) & 1u);
        ptx::tc_fence_after_sync();
        const uint32_t out = tOut + (uint32_t)(s * C::kOutCols) + lane_base;
        uint32_t r[32];
        ptx::tmem_ld_32x32(out, r);
        ptx::tmem_ld_wait();
        if (KV) {
          store_box(&tmDQKV, staging + q * 4096, lane, r, p.round_out != 0, p.d + h * 32, t * 128 + q * 32, b);  // dk
          ptx::tmem_ld_32x32(out + 32, r);
          ptx::tmem_ld_wait();
        }
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bar.o_empty[s]);
        store_box(&tmDQKV, staging + q * 4096, lane, r, p.round_out != 0, (KV ? 2 * p.d : 0) + h * 32, t * 128 + q * 32, b);
      }
    }
    if (part == 0 && lane == 0) ptx::bulk_wait_all();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem, 512);
  }
}

__global__ void attn_mask_kernel(uint8_t* mask, long long slabs, int L, uint32_t thr, unsigned long long seed) {
  const long long n = slabs * L * L;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long row_id = i / L;
    const int key = (int)(i - row_id * L);
    uint32_t hsd = chunk_seed(row_seed((unsigned long long)row_id, seed), key >> 5);
    for (int j = 0; j <= (key & 31); ++j) hsd = hsd * kLcgA + kLcgC;
    mask[i] = hsd >= thr ? 1 : 0;
  }
}

static void drop_params(FaParams& p, float drop_p, uint64_t seed) {
  p.seed = seed;
  if (drop_p > 0.f) {
    double th = (double)drop_p * 4294967296.0;
    p.thr = th >= 4294967295.0 ? 4294967295u : (uint32_t)th;
    if (p.thr == 0u) p.thr = 1u;
    p.dscale = 1.0f / (1.0f - drop_p);
  } else {
    p.thr = 0u;
    p.dscale = 1.0f;
  }
}

static TensorView3 view3(const void* ptr, long long d0, long long d1, long long d2) {
  return TensorView3{ptr, {(unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2},
                     {(unsigned long long)d0 * 4, (unsigned long long)(d0 * d1) * 4}};
}

template <typename K>
static int set_smem(K kernel, int bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    g_last_cuda_error = (int)e;
    return XM_ERR_LAUNCH;
  }
  return XM_OK;
}

}  // namespace fa
}  // namespace xm

using namespace xm;
using namespace xm::fa;

extern "C" {

int xm_attn_fused_fwd_f32(const float* qkv, float* out, float* lse, int64_t B, int64_t L, int64_t H, int64_t dh, float scale,
                          float drop_p, uint64_t seed, int round_out, void* stream) {
  if (!qkv || !out || !lse || !(drop_p >= 0.f && drop_p < 1.f) || B <= 0 || L <= 0 || H <= 0) return XM_ERR_INVALID;
  if (dh != 32 || L > 256 || H > 64 || B * H > 500000000ll) return XM_ERR_UNSUPPORTED;
  FaParams p{};
  p.L = (int)L; p.H = (int)H; p.d = (int)(H * 32);
  p.T = ceil_div(L, 128);
  p.items = (int)(B * H * p.T);
  p.alpha = scale;
  drop_params(p, drop_p, seed);
  p.round_out = round_out;
  p.lse = lse;
  const TensorView3 tq = view3(qkv, 3 * p.d, L, B), to = view3(out, p.d, L, B);
  CUtensorMap mq, mk, mv, mo;
  int rc = encode_tmap(&mq, tq, 32, 128, 0);
  if (rc == XM_OK) rc = encode_tmap(&mk, tq, 32, 256, 0);
  if (rc == XM_OK) rc = encode_tmap(&mv, tq, 32, 256, 1);
  if (rc == XM_OK) rc = encode_tmap(&mo, to, 32, 32, 0);
  if (rc == XM_OK) rc = set_smem(attn_fwd_kernel, kFwdSmem);
  if (rc != XM_OK) return rc;
  const int ctas = p.items < kNumSMs ? p.items : kNumSMs;
  attn_fwd_kernel<<<ctas, kThreads, kFwdSmem, (cudaStream_t)stream>>>(mq, mk, mv, mo, p);
  return check_launch();
}

int xm_attn_fused_bwd_f32(const float* dout, const float* qkv, const float* out, const float* lse, float* dqkv, float* delta,
                          int64_t B, int64_t L, int64_t H, int64_t dh, float scale, float drop_p, uint64_t seed,
                          int round_out, void* stream) {
  if (!dout || !qkv || !out || !lse || !dqkv || !delta || !(drop_p >= 0.f && drop_p < 1.f) || B <= 0 || L <= 0 || H <= 0)
    return XM_ERR_INVALID;
  if (dh != 32 || L > 256 || H > 64 || B * H > 500000000ll) return XM_ERR_UNSUPPORTED;
  FaParams p{};
  p.L = (int)L; p.H = (int)H; p.d = (int)(H * 32);
  p.T = ceil_div(L, 128);
  p.items = (int)(B * H * p.T);
  p.alpha = scale;
  drop_params(p, drop_p, seed);
  p.round_out = round_out;
  p.lse = const_cast<float*>(lse);
  p.delta = delta;
  p.out = out;
  p.dout = dout;
  const TensorView3 tq = view3(qkv, 3 * p.d, L, B), tdo = view3(dout, p.d, L, B), tdq = view3(dqkv, 3 * p.d, L, B);
  CUtensorMap mq, mqn, md, mdn, mo;
  int rc = encode_tmap(&mq, tq, 32, 128, 0);
  if (rc == XM_OK) rc = encode_tmap(&mqn, tq, 32, 128, 1);
  if (rc == XM_OK) rc = encode_tmap(&md, tdo, 32, 128, 0);
  if (rc == XM_OK) rc = encode_tmap(&mdn, tdo, 32, 128, 1);
  if (rc == XM_OK) rc = encode_tmap(&mo, tdq, 32, 32, 0);
  if (rc == XM_OK) rc = set_smem(attn_bwd_kernel<false>, BwdCfg<false>::kSmem);
  if (rc == XM_OK) rc = set_smem(attn_bwd_kernel<true>, BwdCfg<true>::kSmem);
  if (rc != XM_OK) return rc;
  const int ctas = p.items < kNumSMs ? p.items : kNumSMs;
  cudaStream_t st = (cudaStream_t)stream;
  attn_bwd_kernel<false><<<ctas, kThreads, BwdCfg<false>::kSmem, st>>>(mq, mqn, md, mdn, mo, p);  // dq, delta
  rc = check_launch();
  if (rc != XM_OK) return rc;
  attn_bwd_kernel<true><<<ctas, kThreads, BwdCfg<true>::kSmem, st>>>(mq, mqn, md, mdn, mo, p);  // dk, dv
  return check_launch();
}

int xm_attn_fused_mask_u8(uint8_t* mask, int64_t B, int64_t L, int64_t H, float drop_p, uint64_t seed, void* stream) {
  if (!mask || B <= 0 || L <= 0 || H <= 0 || !(drop_p >= 0.f && drop_p < 1.f)) return XM_ERR_INVALID;
  FaParams p{};
  drop_params(p, drop_p, seed);
  const long long n = B * H * L * L;
  const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  attn_mask_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(mask, B * H, (int)L, p.thr, seed);
  return check_launch();
}

}  // extern "C"
