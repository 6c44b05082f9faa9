// Fused multi-head self-attention core for sm_100a (nn.MultiheadAttention inside TemporalTransformerBlock,
// EEG_CODE/enhanced_models_v4.py:71-73,98): the probability matrix never leaves the SM.
//
//   forward   (queries on the TMEM lanes), per 128-key half:  S = q k^T (tcgen05) -> half-row max / exp / dropout
//             by 8 epilogue warps -> unnormalised P~ written BACK INTO TMEM (tcgen05.st) over S -> O_half = P~ v
//             with P~ as the TMEM-resident A operand of the second MMA; the two halves' (max, sum, O) are
//             combined flash-style in the output epilogue -> TMA store.
//   backward, dq (queries on the lanes), per 64-key chunk:  S = q k^T and dP~ = dO v^T side by side in TMEM
//             -> dS = scale * P o (mask/keep * dP~ - delta) over S in TMEM -> dq += dS k (A from TMEM).
//   backward, dk / dv (KEYS on the lanes), per 64-query chunk:  S^T = k q^T and dP~^T = v dO^T
//             -> P~^T and dS^T in place -> dv += P~^T dO, dk += dS^T q (A from TMEM).
//   delta_i = dO_i . O_i (the row sum of P~ o dP~) is formed by the dq kernel and handed to the dk/dv kernel.
//
// Every CTA needs <= 256 TMEM columns, <= 104 KB of shared memory and 320 threads, so TWO CTAs share an SM: inside a
// CTA the phases of an item are serial (MMA -> epilogue -> MMA), and the second CTA fills the tensor pipe / the
// issue slots meanwhile.  An MMA whose A operand comes from TMEM is paced by that read (64 B/clk: a 128 x 8 tf32
// slab per 64 clk whatever N is), which with N = dh = 32 is the floor of these kernels.
//
// HBM traffic per (sample, head): q, k, v, O, dO read a few times (L2) and dq, dk, dv written once -- the
// L x L matrices P~ and dS (4.2 GB each per layer at batch 4096) are never stored.  Supported: dh == 32, L <= 512.
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

constexpr int kEpiWarps = 8;  // 4 TMEM lane quadrants x 2 column parts
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr uint32_t kGold = 0x9E3779B1u;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct FaParams {
  int L, H, d;      // sequence length, heads, d = H*32
  int T;            // 128-row tiles per slab = ceil(L / 128)
  int items;        // B*H*T
  float alpha;      // softmax scale
  uint32_t thr;     // keep <=> stream state >= thr (0: no dropout)
  float dscale;     // 1 / (1 - p)
  unsigned long long seed;
  int round_out;
  float* lse;        // (B*H, L) natural-log logsumexp of the scaled scores
  float* delta;      // (B*H, L)
  const float* out;  // (B, L, d) forward output (backward: delta = dO . O)
  const float* dout; // (B, L, d)
  long long* trace;  // debug (xm_debug_set_attn_trace): CTA 0 appends clock64() at phase boundaries, 4096 slots per role
  float* bias_part;  // backward, may be NULL: (2 * 148 * 4, 3d) zero-initialised partial column sums of dqkv, one row
                     // per (CTA, output warp) -- the bias gradient of the in-projection without a pass over dqkv
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

static long long* g_attn_trace = nullptr;
// Phase tracing (clock64() stamps of CTA 0, see xm_debug_set_attn_trace) is compiled in only with -DXM_FA_TRACE:
// even a never-taken branch per phase costs registers in these register-bound epilogues.
#ifdef XM_FA_TRACE
#define FA_TRACE(role, n)                                                                   \
  do {                                                                                      \
    if (p.trace != nullptr && blockIdx.x == 0 && (n) < 4096) p.trace[(role) * 4096 + (n)++] = clock64(); \
  } while (0)
#else
#define FA_TRACE(role, n) ((void)(n))
#endif

// Forward, one 32-column chunk of a row in registers: e' = exp2(s*c - mcs) (mcs carries -log2(keep_mul), so the kept
// value IS e'), row sum accumulated two lanes wide, dropped / padded entries zeroed.  The arithmetic uses the packed
// f32x2 instructions of sm_100 (one FFMA2 / FADD2 per two elements): these epilogues are FMA-pipe issue bound.
template <bool TAIL>
XM_DEVICE void fwd_chunk(uint32_t (&r)[32], float c, float mcs, uint32_t hsd, uint32_t thr, int nv, float2& sum2) {
  const float2 c2 = make_float2(c, c), m2 = make_float2(-mcs, -mcs);
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    const float2 x = __ffma2_rn(make_float2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), c2, m2);
    hsd = hsd * kLcgA + kLcgC;
    const bool k0 = hsd >= thr;
    hsd = hsd * kLcgA + kLcgC;
    const bool k1 = hsd >= thr;
    float e0 = fast_exp2(x.x), e1 = fast_exp2(x.y);
    if (TAIL) {
      e0 = (j < nv) ? e0 : 0.f;
      e1 = (j + 1 < nv) ? e1 : 0.f;
    }
    sum2 = __fadd2_rn(sum2, make_float2(e0, e1));
    r[j] = __float_as_uint(k0 ? e0 : 0.f);
    r[j + 1] = __float_as_uint(k1 ? e1 : 0.f);
  }
}

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

XM_DEVICE void pair_barrier(int q) { asm volatile("bar.sync %0, 64;" ::"r"(q + 1) : "memory"); }  // the 2 parts of a quadrant
XM_DEVICE void epi_barrier_all() { asm volatile("bar.sync 5, 256;" ::: "memory"); }

// 32 x 32 fp32 block of this warp -> swizzled staging buffer -> TMA store (rows / columns outside the tensor clipped).
// ROWS = 32: one 4 KB box.  ROWS = 16: a 2 KB staging buffer used twice (lanes 0-15, then 16-31) with a {32, 16}
// tensor-map box -- for the kernel whose shared memory has no room for 4 KB per warp.  (A synchronous variant that
// transposed the block through the same buffer and wrote it with st.global was measured 5-10 % slower overall.)
template <int ROWS>
XM_DEVICE void store_box(const CUtensorMap* tm, uint8_t* sb, int lane, const uint32_t (&r)[32], bool round, int col, int row0,
                         int z) {
#pragma unroll
  for (int part = 0; part < 32 / ROWS; ++part) {
    if (lane == 0) ptx::bulk_wait_read<0>();  // the previous store has finished reading this buffer
    __syncwarp();
    if (lane / ROWS == part) {
      const int lr = lane % ROWS;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                               __uint_as_float(r[4 * j + 3]));
        if (round) v = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
        *reinterpret_cast<float4*>(sb + lr * 128 + ((j ^ (lr & 7)) << 4)) = v;
      }
    }
    ptx::fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      ptx::tma_store_3d(tm, sb, col, row0 + part * ROWS, z);
      ptx::bulk_commit();
    }
  }
}

struct Bars {
  uint64_t full, empty;            // per-item operand tiles (single buffer)
  uint64_t hfull[2], hempty[2];    // per-chunk operand ring
  uint64_t s_full, p_ready;        // score accumulators complete / epilogue wrote the TMEM A operands
  uint64_t o_full[2], o_empty[2];  // output accumulator sets
};

XM_DEVICE void init_common(Bars& bar, uint32_t* tmem_slot, int warp, int out_readers = 4) {
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar.full, 1);
    ptx::mbar_init(&bar.empty, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bar.hfull[i], 1);
      ptx::mbar_init(&bar.hempty[i], 1);
      ptx::mbar_init(&bar.o_full[i], 1);
      ptx::mbar_init(&bar.o_empty[i], out_readers);  // the four part-0 warps drain an output accumulator set (backward
                                                     // with bias sums: the four part-1 warps read it too)
    }
    ptx::mbar_init(&bar.s_full, 1);
    ptx::mbar_init(&bar.p_ready, kEpiWarps);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
}

// ============================================================================ forward
// smem: q tile 16 KB | 2 stages of { k block 16 KB (K-major) | v block 16 KB (MN-major) } | 4 staging boxes
// TMEM: [0,128) S / P~ of the current 128-key block ("half": L <= 256 has two);  output accumulators at 128 + ...:
// L <= 256: two sets (item parity) of two blocks;  L <= 512: one set of four blocks
constexpr int kFwdHalf = 2 * 16384;
constexpr int kFwdSmem = 16384 + 2 * kFwdHalf + 4 * 4096 + 1024;

template <bool LONG>  // LONG: 256 < L <= 512 (up to four key blocks, one output set); else at most two, two sets
__global__ void __launch_bounds__(kThreads, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmVmn,
                const __grid_constant__ CUtensorMap tmO, const FaParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ Bars bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float redm[4][2][32], reds[4][2][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* ring = smem + 16384;
  uint8_t* staging = ring + 2 * kFwdHalf;
  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmQ);
    ptx::prefetch_tensormap(&tmVmn);
    ptx::prefetch_tensormap(&tmO);
  }
  init_common(bar, &tmem_slot, warp);
  const uint32_t tmem = tmem_slot;
  const uint32_t tS = tmem, tO = tmem + 128;
  const int NH = p.T;  // 128-key blocks (1..4)
  constexpr bool two_sets = !LONG;
  constexpr int NHMAX = LONG ? 4 : 2;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0, hs = 0;
      for (int w = blockIdx.x; w < p.items; w += gridDim.x, ++it) {
        const int z = w / p.T, qt = w - z * p.T, b = z / p.H, h = z - b * p.H;
        ptx::mbar_wait(&bar.empty, ((uint32_t)it & 1u) ^ 1u);
        ptx::mbar_arrive_expect_tx(&bar.full, 16384);
        ptx::tma_load_3d(&tmQ, &bar.full, smem, h * 32, qt * 128, b);
        for (int hf = 0; hf < NH; ++hf, ++hs) {
          const int st = hs & 1;
          ptx::mbar_wait(&bar.hempty[st], (((uint32_t)hs >> 1) & 1u) ^ 1u);
          uint8_t* yi = ring + st * kFwdHalf;
          ptx::mbar_arrive_expect_tx(&bar.hfull[st], kFwdHalf);
          ptx::tma_load_3d(&tmQ, &bar.hfull[st], yi, p.d + h * 32, hf * 128, b);
          ptx::tma_load_3d(&tmVmn, &bar.hfull[st], yi + 16384, 2 * p.d + h * 32, hf * 128, b);
        }
      }
    }
  } else if (warp == 1) {
    {
      // Warp-uniform schedule: all 32 lanes wait on the barriers and form the descriptors (which then live in uniform
      // registers); each group of MMAs and its commits issue from one `elect_one` branch, back to back in SASS.  With
      // the schedule inside `if (lane == 0)` every MMA paid an R2UR move and an ELECT loop (~125 clk of issue against
      // ~80 for an N = 32 MMA, profiles/r2_mma_rate_probe.txt).
      const uint32_t idesc_s = ptx::make_idesc_tf32(128, 128, 0, 0);
      const uint32_t idesc_o = ptx::make_idesc_tf32(128, 32, 0, 1);
      const uint64_t dq = ptx::make_smem_desc(ptx::smem_u32(smem), 16, 1024, 2);
      const uint64_t dring = ptx::make_smem_desc(ptx::smem_u32(ring), 16, 1024, 2);
      const uint64_t dvring = ptx::make_smem_desc(ptx::smem_u32(ring) + 16384, 4096, 512, 1);
      int it = 0, hs = 0, tn = 0;
      for (int w = blockIdx.x; w < p.items; w += gridDim.x, ++it) {
        const int s = two_sets ? (it & 1) : 0;
        const uint32_t ouse = (uint32_t)(two_sets ? (it >> 1) : it);
        ptx::mbar_wait(&bar.full, (uint32_t)it & 1u);
        for (int hf = 0; hf < NH; ++hf, ++hs) {
          const int st = hs & 1;
          FA_TRACE(0, tn);
          ptx::mbar_wait(&bar.hfull[st], ((uint32_t)hs >> 1) & 1u);
          FA_TRACE(0, tn);
          ptx::tc_fence_after_sync();
          const uint64_t dk = dring + (uint64_t)(st * (kFwdHalf >> 4));
          const uint64_t dv = dvring + (uint64_t)(st * (kFwdHalf >> 4));
          if (ptx::elect_one()) {
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8) ptx::mma_tf32_ss(tS, dq + (uint64_t)(k8 * 2), dk + (uint64_t)(k8 * 2), idesc_s, k8 > 0);
            ptx::mma_commit(&bar.s_full);
            if (hf == NH - 1) ptx::mma_commit(&bar.empty);  // the q tile is not read again: prefetch the next item's
          }
          __syncwarp();
          FA_TRACE(0, tn);
          ptx::mbar_wait(&bar.p_ready, (uint32_t)hs & 1u);
          FA_TRACE(0, tn);
          if (hf == 0) ptx::mbar_wait(&bar.o_empty[s], (ouse & 1u) ^ 1u);
          ptx::tc_fence_after_sync();
          const int nk = min(16, (p.L - hf * 128 + 7) >> 3);  // K = 8 slabs with a valid key
          const uint32_t acc = tO + (uint32_t)(s * 64 + hf * 32);
          if (ptx::elect_one()) {
            if (nk == 16) {
#pragma unroll
              for (int k8 = 0; k8 < 16; ++k8) mma_tf32_ts(acc, tS + (uint32_t)(k8 * 8), dv + (uint64_t)(k8 * 64), idesc_o, k8 > 0);
            } else {
              for (int k8 = 0; k8 < nk; ++k8) mma_tf32_ts(acc, tS + (uint32_t)(k8 * 8), dv + (uint64_t)(k8 * 64), idesc_o, k8 > 0);
            }
            ptx::mma_commit(&bar.hempty[st]);
            if (hf == NH - 1) ptx::mma_commit(&bar.o_full[s]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    const int q = warp & 3, part = (warp - 2) >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const float c = p.alpha * kLog2e;
    const float keep_mul = p.dscale * kTruncComp;  // inverted dropout and truncation compensation ride in the exponent
    const float log2_keep = log2f(keep_mul), inv_keep = 1.0f / keep_mul;
    int it = 0, hs = 0, tn = 0;
    for (int w = blockIdx.x; w < p.items; w += gridDim.x, ++it) {
      const int z = w / p.T, qt = w - z * p.T, b = z / p.H, h = z - b * p.H;
      const int m = qt * 128 + q * 32 + lane;
      const bool live = m < p.L;
      const unsigned long long row_id = (unsigned long long)z * (unsigned long long)p.L + (unsigned long long)m;
      const uint32_t rs = row_seed(row_id, p.seed);
      float mh[NHMAX], lh[NHMAX];  // per block: max * c (log2 domain), sum of exp2
#pragma unroll
      for (int k = 0; k < NHMAX; ++k) mh[k] = lh[k] = 0.f;
      for (int hf = 0; hf < NH; ++hf, ++hs) {
        if (warp == 2 && lane == 0) FA_TRACE(1, tn);
        ptx::mbar_wait(&bar.s_full, (uint32_t)hs & 1u);
        if (warp == 2 && lane == 0) FA_TRACE(1, tn);
        ptx::tc_fence_after_sync();
        const uint32_t acc = tS + lane_base + (uint32_t)(part * 64);
        const int col0 = hf * 128 + part * 64;  // first key of this thread's 64 columns
        // the thread's 64 scores stay in registers from the max to the exp pass: one TMEM read per element
        uint32_t r0[32], r1[32];
        ptx::tmem_ld_32x32(acc, r0);
        ptx::tmem_ld_32x32(acc + 32u, r1);
        ptx::tmem_ld_wait();
        if (warp == 2 && lane == 0) FA_TRACE(1, tn);
        const int nv0 = p.L - col0, nv1 = nv0 - 32;
        float mx = -3.0e38f;
        if (nv1 >= 32) {
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, fmaxf(__uint_as_float(r0[j]), __uint_as_float(r1[j])));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (j < nv0) mx = fmaxf(mx, __uint_as_float(r0[j]));
            if (j < nv1) mx = fmaxf(mx, __uint_as_float(r1[j]));
          }
        }
        redm[q][part][lane] = mx;
        pair_barrier(q);
        if (warp == 2 && lane == 0) FA_TRACE(1, tn);
        mx = fmaxf(redm[q][0][lane], redm[q][1][lane]);  // >= one valid key per half, so finite
        const float mc = mx * c;
        float2 sum2 = make_float2(0.f, 0.f);
        if (nv0 >= 32) fwd_chunk<false>(r0, c, mc - log2_keep, chunk_seed(rs, col0 >> 5), p.thr, nv0, sum2);
        else fwd_chunk<true>(r0, c, mc - log2_keep, chunk_seed(rs, col0 >> 5), p.thr, nv0, sum2);
        ptx::tmem_st_32x32(acc, r0);
        if (nv1 >= 32) fwd_chunk<false>(r1, c, mc - log2_keep, chunk_seed(rs, (col0 >> 5) + 1), p.thr, nv1, sum2);
        else fwd_chunk<true>(r1, c, mc - log2_keep, chunk_seed(rs, (col0 >> 5) + 1), p.thr, nv1, sum2);
        ptx::tmem_st_32x32(acc + 32u, r1);
        const float sum = (sum2.x + sum2.y) * inv_keep;  // back to the sum of the plain exp2(s*c - mc)
        ptx::tmem_st_wait();
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bar.p_ready);
        if (warp == 2 && lane == 0) FA_TRACE(1, tn);
        reds[q][part][lane] = sum;
        pair_barrier(q);
        {
          const float lsum = reds[q][0][lane] + reds[q][1][lane];
#pragma unroll
          for (int k = 0; k < NHMAX; ++k)  // static indices keep the arrays in registers
            if (k == hf) { mh[k] = mc; lh[k] = lsum; }
        }
      }
      if (part == 0) {
        // combine the blocks: out = sum_h O_h a_h / sum_h l_h a_h,  a_h = exp2(m_h - max_h m_h)
        const int s = two_sets ? (it & 1) : 0;
        const uint32_t ouse = (uint32_t)(two_sets ? (it >> 1) : it);
        float mtot = mh[0];
#pragma unroll
        for (int k = 1; k < NHMAX; ++k)
          if (k < NH) mtot = fmaxf(mtot, mh[k]);
        float ah[NHMAX], denom = 0.f;
#pragma unroll
        for (int k = 0; k < NHMAX; ++k) {
          ah[k] = (k < NH) ? fast_exp2(mh[k] - mtot) : 0.f;
          denom = fmaf(lh[k], ah[k], denom);
        }
        if (live) p.lse[row_id] = (mtot + log2f(denom)) * kLn2;
        const float inv = 1.0f / denom;
        if (warp == 2 && lane == 0) FA_TRACE(1, tn);
        ptx::mbar_wait(&bar.o_full[s], ouse & 1u);
        if (warp == 2 && lane == 0) FA_TRACE(1, tn);
        ptx::tc_fence_after_sync();
        uint32_t r[32];
        ptx::tmem_ld_32x32(tO + (uint32_t)(s * 64) + lane_base, r);
        ptx::tmem_ld_wait();
        {
          const float a0 = ah[0] * inv;
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * a0);
        }
#pragma unroll
        for (int k = 1; k < NHMAX; ++k) {
          if (k < NH) {
            uint32_t u[32];
            ptx::tmem_ld_32x32(tO + (uint32_t)(s * 64 + k * 32) + lane_base, u);
            ptx::tmem_ld_wait();
            const float ak = ah[k] * inv;
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(fmaf(__uint_as_float(u[j]), ak, __uint_as_float(r[j])));
          }
        }
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bar.o_empty[s]);
        store_box<32>(&tmO, staging + q * 4096, lane, r, p.round_out != 0, h * 32, qt * 128 + q * 32, b);
      }
    }
    if (part == 0 && lane == 0) ptx::bulk_wait_all();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem, 256);
  }
}

// ============================================================================ backward
// One template for both backward kernels.  KV == false: lanes = queries of tile `t`, chunks walk the keys, output
// dq.  KV == true: lanes = keys of tile `t`, chunks walk the queries, outputs dk and dv.
//   per item  (single buffer): two lane-side tiles X0, X1 of 16 KB, K-major       (dq: q, dO ;  dk/dv: k, v)
//   per chunk (2 stages)     : Y0, Y1 of 64 rows, K-major, 8 KB each              (dq: k, v  ;  dk/dv: q, dO)
//                              + MN-major copies for the output products          (dq: k     ;  dk/dv: dO, q)
//   TMEM: [0,64) S (then dS | P~^T)   [64,128) dP~ (then dS^T)   [128, ...) two sets of output accumulators
//   Staging: 4 KB per part-0 warp (dq) or 2 KB used twice (dk/dv, whose smem also holds the per-query tables).
constexpr int kItemBytes = 2 * 16384;
template <bool KV>
struct BwdCfg {
  static constexpr int kChunkBytes = (KV ? 4 : 3) * 8192;
  static constexpr int kRingOff = kItemBytes;
  static constexpr int kStagingOff = kRingOff + 2 * kChunkBytes;
  static constexpr int kStageRows = KV ? 16 : 32;
  static constexpr int kTabOff = kStagingOff + 4 * kStageRows * 128;  // KV: lse/delta + seeds [4] of 256 queries
  static constexpr int kSmem = kTabOff + (KV ? 6 * 1024 : 0) + 1024;   // (refilled mid-item when L > 256)
  static constexpr int kOutCols = KV ? 64 : 32;  // per accumulator set
};

template <bool KV, bool LONG, bool BIAS>  // LONG (dk/dv only): 256 < L <= 512, the per-query tables are refilled mid-item
__global__ void __launch_bounds__(kThreads, 2)                   // BIAS: also emit partial column sums of the outputs
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDOx,
                const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmYmn,
                const __grid_constant__ CUtensorMap tmDOy, const __grid_constant__ CUtensorMap tmDOymn,
                const __grid_constant__ CUtensorMap tmDQKV, const FaParams p) {
  // tmX / tmDOx: qkv / dout with 128-row boxes (item tiles); tmY / tmDOy (K-major) and tmYmn / tmDOymn (MN-major):
  // the same tensors with 64-row boxes (chunk tiles)
  using C = BwdCfg<KV>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ Bars bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float red[4][2][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* ring = smem + C::kRingOff;
  uint8_t* staging = smem + C::kStagingOff;
  float* tab_off = reinterpret_cast<float*>(smem + C::kTabOff);  // lse in the log2 domain (minus log2 of the ds scale)
  float* tab_delta = tab_off + 256;
  uint32_t* tab_seed = reinterpret_cast<uint32_t*>(tab_delta + 256);  // [quadrant = key chunk of the tile][query]
  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmX);
    ptx::prefetch_tensormap(&tmDOx);
    ptx::prefetch_tensormap(&tmY);
    ptx::prefetch_tensormap(&tmYmn);
    ptx::prefetch_tensormap(&tmDOy);
    ptx::prefetch_tensormap(&tmDOymn);
    ptx::prefetch_tensormap(&tmDQKV);
  }
  init_common(bar, &tmem_slot, warp, BIAS ? 8 : 4);
  const uint32_t tmem = tmem_slot;
  const uint32_t tS = tmem, tP = tmem + 64, tOut = tmem + 128;
  const int NC = (p.L + 63) >> 6;  // 64-wide chunks of the other index

  if (warp == 0) {
    if (lane == 0) {
      int it = 0, hs = 0;
      for (int w = blockIdx.x; w < p.items; w += gridDim.x, ++it) {
        const int z = w / p.T, t = w - z * p.T, b = z / p.H, h = z - b * p.H;
        ptx::mbar_wait(&bar.empty, ((uint32_t)it & 1u) ^ 1u);
        ptx::mbar_arrive_expect_tx(&bar.full, kItemBytes);
        if (KV) {
          ptx::tma_load_3d(&tmX, &bar.full, smem, p.d + h * 32, t * 128, b);              // k tile
          ptx::tma_load_3d(&tmX, &bar.full, smem + 16384, 2 * p.d + h * 32, t * 128, b);  // v tile
        } else {
          ptx::tma_load_3d(&tmX, &bar.full, smem, h * 32, t * 128, b);             // q tile
          ptx::tma_load_3d(&tmDOx, &bar.full, smem + 16384, h * 32, t * 128, b);   // dO tile
        }
        for (int cc = 0; cc < NC; ++cc, ++hs) {
          const int st = hs & 1;
          ptx::mbar_wait(&bar.hempty[st], (((uint32_t)hs >> 1) & 1u) ^ 1u);
          uint8_t* yi = ring + st * C::kChunkBytes;
          ptx::mbar_arrive_expect_tx(&bar.hfull[st], C::kChunkBytes);
          if (KV) {
            ptx::tma_load_3d(&tmY, &bar.hfull[st], yi, h * 32, cc * 64, b);               // q chunk (K-major)
            ptx::tma_load_3d(&tmDOy, &bar.hfull[st], yi + 8192, h * 32, cc * 64, b);      // dO chunk (K-major)
            ptx::tma_load_3d(&tmDOymn, &bar.hfull[st], yi + 16384, h * 32, cc * 64, b);   // dO chunk (MN-major) -> dv
            ptx::tma_load_3d(&tmYmn, &bar.hfull[st], yi + 24576, h * 32, cc * 64, b);     // q chunk (MN-major) -> dk
          } else {
            ptx::tma_load_3d(&tmY, &bar.hfull[st], yi, p.d + h * 32, cc * 64, b);             // k chunk
            ptx::tma_load_3d(&tmY, &bar.hfull[st], yi + 8192, 2 * p.d + h * 32, cc * 64, b);  // v chunk
            ptx::tma_load_3d(&tmYmn, &bar.hfull[st], yi + 16384, p.d + h * 32, cc * 64, b);   // k chunk (MN) -> dq
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      // warp-uniform schedule, MMAs issued from elect_one branches (see the forward kernel's issuer)
      const uint32_t idesc_s = ptx::make_idesc_tf32(128, 64, 0, 0);
      const uint32_t idesc_o = ptx::make_idesc_tf32(128, 32, 0, 1);
      const uint32_t xa = ptx::smem_u32(smem);
      const uint64_t dx0 = ptx::make_smem_desc(xa, 16, 1024, 2);
      const uint64_t dx1 = ptx::make_smem_desc(xa + 16384, 16, 1024, 2);
      const uint32_t ra = ptx::smem_u32(ring);
      const uint64_t dr0 = ptx::make_smem_desc(ra, 16, 1024, 2);
      const uint64_t dr1 = ptx::make_smem_desc(ra + 8192, 16, 1024, 2);
      const uint64_t drm0 = ptx::make_smem_desc(ra + 16384, 4096, 512, 1);
      const uint64_t drm1 = ptx::make_smem_desc(ra + 24576, 4096, 512, 1);
      int it = 0, hs = 0, tn = 0;
      for (int w = blockIdx.x; w < p.items; w += gridDim.x, ++it) {
        const int s = it & 1;
        FA_TRACE(0, tn);
        ptx::mbar_wait(&bar.full, (uint32_t)it & 1u);
        const uint32_t out = tOut + (uint32_t)(s * C::kOutCols);
        for (int cc = 0; cc < NC; ++cc, ++hs) {
          const int st = hs & 1;
          FA_TRACE(0, tn);
          ptx::mbar_wait(&bar.hfull[st], ((uint32_t)hs >> 1) & 1u);
          FA_TRACE(0, tn);
          ptx::tc_fence_after_sync();
          const uint64_t so = (uint64_t)(st * (C::kChunkBytes >> 4));
          const uint64_t dy0 = dr0 + so, dy1 = dr1 + so;
          if (ptx::elect_one()) {
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8)  // S (or S^T): lane-side tile 0 times chunk tile 0
              ptx::mma_tf32_ss(tS, dx0 + (uint64_t)(k8 * 2), dy0 + (uint64_t)(k8 * 2), idesc_s, k8 > 0);
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8)  // dP~ (or its transpose): lane-side tile 1 times chunk tile 1
              ptx::mma_tf32_ss(tP, dx1 + (uint64_t)(k8 * 2), dy1 + (uint64_t)(k8 * 2), idesc_s, k8 > 0);
            ptx::mma_commit(&bar.s_full);
            if (cc == NC - 1) ptx::mma_commit(&bar.empty);  // the item tiles are not read again: prefetch the next item's
          }
          __syncwarp();
          FA_TRACE(0, tn);
          ptx::mbar_wait(&bar.p_ready, (uint32_t)hs & 1u);
          FA_TRACE(0, tn);
          if (cc == 0) ptx::mbar_wait(&bar.o_empty[s], (((uint32_t)it >> 1) & 1u) ^ 1u);
          ptx::tc_fence_after_sync();
          const uint64_t dm0 = drm0 + so, dm1 = drm1 + so;
          const uint32_t acc0 = cc > 0 ? 1u : 0u;
          if (ptx::elect_one()) {
            if (KV) {
#pragma unroll
              for (int k8 = 0; k8 < 8; ++k8)  // dv += P~^T dO
                mma_tf32_ts(out + 32, tS + (uint32_t)(k8 * 8), dm0 + (uint64_t)(k8 * 64), idesc_o, k8 > 0 ? 1u : acc0);
#pragma unroll
              for (int k8 = 0; k8 < 8; ++k8)  // dk += dS^T q
                mma_tf32_ts(out, tP + (uint32_t)(k8 * 8), dm1 + (uint64_t)(k8 * 64), idesc_o, k8 > 0 ? 1u : acc0);
            } else {
#pragma unroll
              for (int k8 = 0; k8 < 8; ++k8)  // dq += dS k
                mma_tf32_ts(out, tS + (uint32_t)(k8 * 8), dm0 + (uint64_t)(k8 * 64), idesc_o, k8 > 0 ? 1u : acc0);
            }
            ptx::mma_commit(&bar.hempty[st]);
            if (cc == NC - 1) ptx::mma_commit(&bar.o_full[s]);
          }
          __syncwarp();
        }
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
    const int et = threadIdx.x - 64;  // 0..255 among the epilogue threads
    // keys on the lanes: LCG jump to state lane + 1 of a chunk's stream
    uint32_t jump_a = 1u, jump_c = 0u;
    if (KV) {
      for (int k = 0; k <= lane; ++k) {
        jump_c = jump_c * kLcgA + kLcgC;
        jump_a *= kLcgA;
      }
    }
    int it = 0, hs = 0, tn = 0;
    // dk/dv kernel: the per-query lse / delta of the NEXT item are fetched into registers one item ahead, so the
    // table refresh at an item boundary does not wait for global memory
    constexpr int NQH = LONG ? 2 : 1;
    float nx_lse[NQH], nx_delta[NQH];  // queries et (and et + 256)
#pragma unroll
    for (int k = 0; k < NQH; ++k) nx_lse[k] = nx_delta[k] = 0.f;
    if (KV && (int)blockIdx.x < p.items) {
      const unsigned long long o = (unsigned long long)(blockIdx.x / p.T) * (unsigned long long)p.L;
#pragma unroll
      for (int k = 0; k < NQH; ++k)
        if (et + 256 * k < p.L) {
          nx_lse[k] = __ldg(p.lse + o + et + 256 * k);
          nx_delta[k] = __ldg(p.delta + o + et + 256 * k);
        }
    }
    for (int w = blockIdx.x; w < p.items; w += gridDim.x, ++it) {
      const int z = w / p.T, t = w - z * p.T, b = z / p.H, h = z - b * p.H;
      const int row = t * 128 + q * 32 + lane;  // query (dq kernel) or key (dk/dv kernel) of this thread
      const bool live = row < p.L;
      const unsigned long long slab0 = (unsigned long long)z * (unsigned long long)p.L;
      float off = 0.f, delta = 0.f;
      uint32_t rs = 0u;
      // dk/dv kernel: per-query tables (exponent offset, delta, mask-stream seeds of this tile's four key chunks) for
      // queries [256 k, 256 k + 256); refilled from the prefetch registers, which then fetch the next item's values
      auto fill_tables = [&](float& pre_lse, float& pre_delta, int k) {
        epi_barrier_all();  // every warp is done with the previous contents
        const int m = et + 256 * k;
        tab_off[et] = pre_lse * kLog2e - log2_amul;  // rows >= L: 0 (finite; their q / dO rows are zero)
        tab_delta[et] = pre_delta;
        const uint32_t rsm = row_seed(slab0 + (unsigned long long)m, p.seed);
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) tab_seed[qd * 256 + et] = chunk_seed(rsm, t * 4 + qd);
        pre_lse = 0.f;
        pre_delta = 0.f;
        const int wn = w + (int)gridDim.x;
        if (wn < p.items && m < p.L) {
          const unsigned long long on = (unsigned long long)(wn / p.T) * (unsigned long long)p.L + m;
          pre_lse = __ldg(p.lse + on);
          pre_delta = __ldg(p.delta + on);
        }
        epi_barrier_all();
      };
      if (KV) {
        fill_tables(nx_lse[0], nx_delta[0], 0);
      } else {
        // delta = dO . O over the head's 32 columns: each part takes 16 of them
        float part_sum = 0.f;
        if (live) {
          const long long o = ((long long)b * p.L + row) * p.d + h * 32 + part * 16;
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(p.dout + o) + v4);
            const float4 g = __ldg(reinterpret_cast<const float4*>(p.out + o) + v4);
            part_sum += a.x * g.x + a.y * g.y + a.z * g.z + a.w * g.w;
          }
          off = __ldg(p.lse + slab0 + row) * kLog2e;
        }
        off -= log2_amul;
        red[q][part][lane] = part_sum;
        pair_barrier(q);
        delta = red[q][0][lane] + red[q][1][lane];
        pair_barrier(q);
        if (live && part == 0) p.delta[slab0 + row] = delta;
        rs = row_seed(slab0 + (unsigned long long)row, p.seed);
      }
      if (warp == 2 && lane == 0) FA_TRACE(1, tn);
      for (int cc = 0; cc < NC; ++cc, ++hs) {
        if (KV && LONG && cc == 4) fill_tables(nx_lse[NQH - 1], nx_delta[NQH - 1], 1);  // second half of the queries
        if (warp == 2 && lane == 0) FA_TRACE(1, tn);
        ptx::mbar_wait(&bar.s_full, (uint32_t)hs & 1u);
        if (warp == 2 && lane == 0) FA_TRACE(1, tn);
        ptx::tc_fence_after_sync();
        uint32_t r[32], g[32];
        ptx::tmem_ld_32x32(tS + lane_base + (uint32_t)(part * 32), r);
        ptx::tmem_ld_32x32(tP + lane_base + (uint32_t)(part * 32), g);
        const int o0 = cc * 64 + part * 32;  // first key (dq kernel) / query (dk/dv kernel) of this chunk
        if (KV) {
          const int ot = o0 & 255;  // position inside the resident half of the tables
          const uint32_t* sd = tab_seed + q * 256 + ot;
          ptx::tmem_ld_wait();
          const float2 c2 = make_float2(c, c);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {  // per-query constants: one broadcast 128-bit load per 4 columns
            const float4 o4 = *reinterpret_cast<const float4*>(tab_off + ot + 4 * j4);
            const float4 d4 = *reinterpret_cast<const float4*>(tab_delta + ot + 4 * j4);
            const uint4 s4 = *reinterpret_cast<const uint4*>(sd + 4 * j4);
            const float2 of[2] = {make_float2(-o4.x, -o4.y), make_float2(-o4.z, -o4.w)};
            const float2 de[2] = {make_float2(-d4.x, -d4.y), make_float2(-d4.z, -d4.w)};
            const uint32_t sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int j = 4 * j4 + 2 * e;
              const float2 x = __ffma2_rn(make_float2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), c2, of[e]);
              const bool k0 = sv[2 * e] * jump_a + jump_c >= p.thr;
              const bool k1 = sv[2 * e + 1] * jump_a + jump_c >= p.thr;
              const float2 pr = make_float2(fast_exp2(x.x), fast_exp2(x.y));
              const float2 t2 = __ffma2_rn(make_float2(k0 ? p.dscale : 0.f, k1 ? p.dscale : 0.f),
                                           make_float2(__uint_as_float(g[j]), __uint_as_float(g[j + 1])), de[e]);
              const float2 ds = __fmul2_rn(pr, t2);
              const float2 pt = __fmul2_rn(pr, make_float2(k0 ? pt_keep : 0.f, k1 ? pt_keep : 0.f));
              r[j] = __float_as_uint(pt.x);
              r[j + 1] = __float_as_uint(pt.y);
              g[j] = __float_as_uint(ds.x);
              g[j + 1] = __float_as_uint(ds.y);
            }
          }
          ptx::tmem_st_32x32(tS + lane_base + (uint32_t)(part * 32), r);
          ptx::tmem_st_32x32(tP + lane_base + (uint32_t)(part * 32), g);
        } else {
          uint32_t hsd = chunk_seed(rs, o0 >> 5);
          const float2 c2 = make_float2(c, c), no2 = make_float2(-off, -off), nd2 = make_float2(-delta, -delta);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), c2, no2);
            hsd = hsd * kLcgA + kLcgC;
            const float m0 = (hsd >= p.thr) ? p.dscale : 0.f;
            hsd = hsd * kLcgA + kLcgC;
            const float m1 = (hsd >= p.thr) ? p.dscale : 0.f;
            const float2 t2 = __ffma2_rn(make_float2(m0, m1), make_float2(__uint_as_float(g[j]), __uint_as_float(g[j + 1])), nd2);
            const float2 d2 = __fmul2_rn(make_float2(fast_exp2(x.x), fast_exp2(x.y)), t2);
            r[j] = __float_as_uint(d2.x);
            r[j + 1] = __float_as_uint(d2.y);
          }
          ptx::tmem_st_32x32(tS + lane_base + (uint32_t)(part * 32), r);
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bar.p_ready);
        if (warp == 2 && lane == 0) FA_TRACE(1, tn);
      }
      if (part == 0) {
        const int s = it & 1;
        ptx::mbar_wait(&bar.o_full[s], ((uint32_t)it >> 1) & 1u);
        if (warp == 2 && lane == 0) FA_TRACE(1, tn);
        ptx::tc_fence_after_sync();
        const uint32_t out = tOut + (uint32_t)(s * C::kOutCols) + lane_base;
        uint32_t r[32];
        ptx::tmem_ld_32x32(out, r);
        ptx::tmem_ld_wait();
        uint8_t* sb = staging + q * (C::kStageRows * 128);
        if (KV) {
          store_box<C::kStageRows>(&tmDQKV, sb, lane, r, p.round_out != 0, p.d + h * 32, t * 128 + q * 32, b);  // dk
          ptx::tmem_ld_32x32(out + 32, r);
          ptx::tmem_ld_wait();
        }
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bar.o_empty[s]);
        store_box<C::kStageRows>(&tmDQKV, sb, lane, r, p.round_out != 0, (KV ? 2 * p.d : 0) + h * 32, t * 128 + q * 32, b);
      } else if (BIAS) {
        // The part-1 warps are idle between items: they read the same output accumulators (their lane quadrant) and add
        // the column sums of the 32 output rows to this warp's private row of partial sums -- the in-projection's bias
        // gradient without a pass over dqkv.  Rows past L hold products of zero-filled tiles with nonzero
        // probabilities: excluded.  RED (fire and forget): the row is private to the warp, so the sum's order is fixed.
        const int s = it & 1;
        ptx::mbar_wait(&bar.o_full[s], ((uint32_t)it >> 1) & 1u);
        ptx::tc_fence_after_sync();
        const uint32_t out = tOut + (uint32_t)(s * C::kOutCols) + lane_base;
        const bool live_row = t * 128 + q * 32 + lane < p.L;
        float* bp = p.bias_part + (long long)(blockIdx.x * 4 + q) * (3 * p.d) + h * 32 + lane;
        uint32_t r[32];
#pragma unroll
        for (int o = 0; o < (KV ? 2 : 1); ++o) {
          ptx::tmem_ld_32x32(out + (uint32_t)(o * 32), r);
          ptx::tmem_ld_wait();
          if (o == (KV ? 1 : 0)) {
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&bar.o_empty[s]);
          }
          if (!live_row) {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = 0u;
          }
          const float cs = warp_column_sums(r, lane);
          atomicAdd(bp + (KV ? (o == 0 ? p.d : 2 * p.d) : 0), cs);
        }
      }
    }
    if (part == 0 && lane == 0) ptx::bulk_wait_all();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem, 256);
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

static int fill_params(FaParams& p, int64_t B, int64_t L, int64_t H, int64_t dh, float scale, float drop_p, uint64_t seed,
                       int round_out) {
  if (!(drop_p >= 0.f && drop_p < 1.f) || B <= 0 || L <= 0 || H <= 0) return XM_ERR_INVALID;
  if (dh != 32 || L > 512 || H > 64 || B * H > 250000000ll) return XM_ERR_UNSUPPORTED;
  p.L = (int)L; p.H = (int)H; p.d = (int)(H * 32);
  p.T = ceil_div(L, 128);
  p.items = (int)(B * H * p.T);
  p.alpha = scale;
  drop_params(p, drop_p, seed);
  p.round_out = round_out;
  p.trace = g_attn_trace;
  return XM_OK;
}

}  // namespace fa
}  // namespace xm

XM_DEFINE_SEED_EPOCH_SLOT(attention_fused)

using namespace xm;
using namespace xm::fa;

extern "C" {

int xm_attn_fused_fwd_f32(const float* qkv, float* out, float* lse, int64_t B, int64_t L, int64_t H, int64_t dh, float scale,
                          float drop_p, uint64_t seed, int round_out, void* stream) {
  if (!qkv || !out || !lse) return XM_ERR_INVALID;
  FaParams p{};
  int rc = fill_params(p, B, L, H, dh, scale, drop_p, seed, round_out);
  if (rc != XM_OK) return rc;
  p.lse = lse;
  const TensorView3 tq = view3(qkv, 3 * p.d, L, B), to = view3(out, p.d, L, B);
  CUtensorMap mq, mv, mo;
  rc = encode_tmap(&mq, tq, 32, 128, 0);
  if (rc == XM_OK) rc = encode_tmap(&mv, tq, 32, 128, 1);
  if (rc == XM_OK) rc = encode_tmap(&mo, to, 32, 32, 0);
  const bool lng = L > 256;
  if (rc == XM_OK) rc = lng ? set_smem(attn_fwd_kernel<true>, kFwdSmem) : set_smem(attn_fwd_kernel<false>, kFwdSmem);
  if (rc != XM_OK) return rc;
  const int ctas = p.items < 2 * kNumSMs ? p.items : 2 * kNumSMs;
  if (lng) attn_fwd_kernel<true><<<ctas, kThreads, kFwdSmem, (cudaStream_t)stream>>>(mq, mv, mo, p);
  else attn_fwd_kernel<false><<<ctas, kThreads, kFwdSmem, (cudaStream_t)stream>>>(mq, mv, mo, p);
  return check_launch();
}

int xm_attn_fused_bwd_nblk(void) { return 2 * kNumSMs * 4; }

int xm_attn_fused_bwd_f32(const float* dout, const float* qkv, const float* out, const float* lse, float* dqkv, float* delta,
                          float* dbias_part, int64_t B, int64_t L, int64_t H, int64_t dh, float scale, float drop_p,
                          uint64_t seed, int round_out, void* stream) {
  if (!dout || !qkv || !out || !lse || !dqkv || !delta) return XM_ERR_INVALID;
  FaParams p{};
  int rc = fill_params(p, B, L, H, dh, scale, drop_p, seed, round_out);
  if (rc != XM_OK) return rc;
  p.bias_part = dbias_part;
  if (dbias_part &&
      cudaMemsetAsync(dbias_part, 0, (size_t)xm_attn_fused_bwd_nblk() * 3 * p.d * sizeof(float), (cudaStream_t)stream) != cudaSuccess) {
    g_last_cuda_error = (int)cudaGetLastError();
    return XM_ERR_LAUNCH;
  }
  p.lse = const_cast<float*>(lse);
  p.delta = delta;
  p.out = out;
  p.dout = dout;
  if ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(dout)) & 15) return XM_ERR_INVALID;
  const TensorView3 tq = view3(qkv, 3 * p.d, L, B), tdo = view3(dout, p.d, L, B), tdq = view3(dqkv, 3 * p.d, L, B);
  CUtensorMap mx, mdx, my, myn, mdy, mdyn, mo, mo16;
  rc = encode_tmap(&mx, tq, 32, 128, 0);
  if (rc == XM_OK) rc = encode_tmap(&mdx, tdo, 32, 128, 0);
  if (rc == XM_OK) rc = encode_tmap(&my, tq, 32, 64, 0);
  if (rc == XM_OK) rc = encode_tmap(&myn, tq, 32, 64, 1);
  if (rc == XM_OK) rc = encode_tmap(&mdy, tdo, 32, 64, 0);
  if (rc == XM_OK) rc = encode_tmap(&mdyn, tdo, 32, 64, 1);
  if (rc == XM_OK) rc = encode_tmap(&mo, tdq, 32, 32, 0);
  if (rc == XM_OK) rc = encode_tmap(&mo16, tdq, 32, 16, 0);
  const bool lng = L > 256;
  const int ctas = p.items < 2 * kNumSMs ? p.items : 2 * kNumSMs;
  cudaStream_t st = (cudaStream_t)stream;
  auto launch = [&](auto kernel, int smem, const CUtensorMap& out_map) {
    int r = set_smem(kernel, smem);
    if (r != XM_OK) return r;
    kernel<<<ctas, kThreads, smem, st>>>(mx, mdx, my, myn, mdy, mdyn, out_map, p);
    return check_launch();
  };
  const bool bias = dbias_part != nullptr;
  // dq, delta
  rc = bias ? launch(attn_bwd_kernel<false, false, true>, BwdCfg<false>::kSmem, mo)
            : launch(attn_bwd_kernel<false, false, false>, BwdCfg<false>::kSmem, mo);
  if (rc != XM_OK) return rc;
  // dk, dv
  if (lng) return bias ? launch(attn_bwd_kernel<true, true, true>, BwdCfg<true>::kSmem, mo16)
                       : launch(attn_bwd_kernel<true, true, false>, BwdCfg<true>::kSmem, mo16);
  return bias ? launch(attn_bwd_kernel<true, false, true>, BwdCfg<true>::kSmem, mo16)
              : launch(attn_bwd_kernel<true, false, false>, BwdCfg<true>::kSmem, mo16);
}

int xm_debug_set_attn_trace(int64_t* device_buffer) {
  xm::fa::g_attn_trace = reinterpret_cast<long long*>(device_buffer);
#ifdef XM_FA_TRACE
  return XM_OK;
#else
  return device_buffer ? XM_ERR_UNSUPPORTED : XM_OK;  // rebuild with -DXM_FA_TRACE (XM_NVCC_FLAGS) to trace
#endif
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
