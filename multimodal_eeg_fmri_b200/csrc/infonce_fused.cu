// Fused backward of the symmetric InfoNCE loss for sm_100a: the (local batch x GLOBAL batch) softmax-gradient blocks
//
//   G1[i][j] = coef * (exp(S_ij - lse_ef[i]) + exp(S_ij - lse_fe_all[j]) - 2 [j == i + diag_off]),  S = e_n f_n^T / tau
//   G2[i][j] = the same with the roles of e and f exchanged
//
// are produced tile by tile in tensor memory and consumed in place as the A operand of  dE = G1 f_n,  dF = G2 e_n:
// nothing of size (Ml x Ng) is written (the unfused chain -- xm_infonce_grad_f32, xm_split3_f32, xm_infonce_dgrad_f32 --
// materialises both blocks, 2 x 537 MB at 8 x 4096 samples, and reads them back).  No reference implementation
// (an authored definition, as the unfused kernels).
//
// Work unit = (direction, 128-row tile of the local batch, range of 128-column chunks of the global batch); one
// persistent CTA per SM walks units, both directions in one launch.  Per chunk j (structure of ffn_fused.cu):
//   M1(j)  S_j = a_hi b_hi^T + a_lo b_hi^T + a_hi b_lo^T   fp32-accurate scores: 3 groups of 16 MMAs over the tf32 splits
//                                                         (a: resident 128 KB tile, b: 16 KB k-blocks through a ring)
//   T(j)   8 transform warps: g = exp2(S c) * (rowscale_i + colscale_j) (- 2 coef on the positive), one exponential per
//          element; precise mode: hi = g with the low 13 bits cleared (what the tensor core keeps), lo = g - hi;
//          hi is written back over S, lo into a second TMEM buffer
//   M2(j)  X = G_lo B_hi + G_hi B_hi + G_hi B_lo  (precise)  |  X = G_hi B_hi  (single pass), A operand from TMEM,
//          B = the transposed unit vectors (D x Ng, K-major) made once per call by nce_prep_kernel
//   D(j)   all 16 warps add X into fp32 registers (the tensor core's own accumulation truncates: one chunk per TMEM
//          accumulation keeps the sum over the global batch fp32-accurate, as GemmParams::acc_chunk does for the
//          unfused product); the running sums are written at the end of the unit (red.add when a row tile is split).
// MMA order  M1(0) M1(1) | M2(0) M1(2) | M2(1) M1(3) | ...   TMEM: S0 [0,128) S1 [128,256) L [256,384) X [384,512).
//
// The forward (xm_infonce_lse_fused_f32) is the same kernel without M2 / D: T(j) adds exp(S_ij - 1/tau) of its columns
// to a per-thread row sum (|S| <= 1/tau for unit vectors: no running maximum needed) and picks up the positive S_ii;
// the per-thread partial sums go to a workspace and are combined in a fixed order (deterministic).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/xmodal_b200.h"
#include "gemm_engine.cuh"

namespace xm {
namespace nce {

constexpr int kD = 128, kChunk = 128;
constexpr int kTile = 16384;  // 128 rows x 32 fp32, SWIZZLE_128B
constexpr uint32_t kTile16 = kTile >> 4;
constexpr int kRing = 6;
constexpr int kXfWarps = 16;
constexpr int kThreads = 64 + 32 * kXfWarps;
constexpr int kSmem = 8 * kTile + kRing * kTile + 1024;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kTruncComp = 1.0f + 0.7213f / 2048.0f;  // see ffn_fused.cu

struct Params {
  int row_tiles, nc, cps, units;  // cps: chunks per unit
  int a_lo_col[2], b_lo_col[2];   // column of the lo block in a row of the local / global split (hi is at column 0)
  const float* lse_row[2];
  const float* colscale[2];       // scale * exp(-lse_col[j])
  float* out[2];                  // (Ml, 128)
  int diag_off;
  int atomic;   // a row tile is split over several units: 1 (two units) = outputs zeroed and accumulated with red.add
                // (0 + a + b is the same in either order); 2 (more) = unit sp writes slab sp of `part`, summed in order
  float* part[2];         // atomic == 2: (splits, Ml, 128) partial outputs per direction
  long long part_stride;  // Ml * 128
  float c;      // log2(e) / tau
  float scale;  // coef (x truncation compensation in the single-pass mode)
  // forward (MODE_LSE) only
  float* partial[2];  // (Ml, slots) partial sums of exp(S - 1/tau): slot = (column split * 2 + transform group) * 2 + column half
  float* diag;        // (Ml) S[i, i + diag_off]
  int slots;
  float inv_tau;
};
enum : int { MODE_SINGLE = 0, MODE_PRECISE = 1, MODE_LSE = 2 };

struct Bars {
  uint64_t a_full, a_empty;
  uint64_t w_full[kRing], w_empty[kRing];
  uint64_t s_full[2], a_ready[2], l_free[2];
  uint64_t x_full, x_free;
};

// Position in the CTA's chunk sequence: unit ordinal k (unit = blockIdx.x + k * gridDim.x), running index n.
struct Cursor {
  int k, n, c, c0, c1, dir, row0, sp;
  bool ok;
};
XM_DEVICE void cur_load(Cursor& cu, const Params& p) {
  const int u = (int)blockIdx.x + cu.k * (int)gridDim.x;
  cu.ok = u < p.units;
  if (!cu.ok) return;
  cu.dir = u & 1;
  const int rest = u >> 1, sp = rest / p.row_tiles;
  cu.row0 = (rest - sp * p.row_tiles) * 128;
  cu.sp = sp;
  cu.c0 = sp * p.cps;
  cu.c1 = min(p.nc, cu.c0 + p.cps);
  cu.c = cu.c0;
}
XM_DEVICE void cur_init(Cursor& cu, const Params& p) {
  cu.k = 0;
  cu.n = 0;
  cur_load(cu, p);
}
XM_DEVICE void cur_next(Cursor& cu, const Params& p) {
  ++cu.n;
  if (++cu.c == cu.c1) {
    ++cu.k;
    cur_load(cu, p);
  }
}

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

XM_DEVICE void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

XM_DEVICE void store_row32(float* dst, const float (&r)[32]) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 8 * j), "r"(__float_as_uint(r[8 * j])),
                 "r"(__float_as_uint(r[8 * j + 1])), "r"(__float_as_uint(r[8 * j + 2])), "r"(__float_as_uint(r[8 * j + 3])),
                 "r"(__float_as_uint(r[8 * j + 4])), "r"(__float_as_uint(r[8 * j + 5])), "r"(__float_as_uint(r[8 * j + 6])),
                 "r"(__float_as_uint(r[8 * j + 7]))
                 : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1)
nce_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
               const __grid_constant__ CUtensorMap tmT0, const __grid_constant__ CUtensorMap tmT1, const Params p) {
  constexpr bool PRECISE = MODE == MODE_PRECISE, LSE = MODE == MODE_LSE;
  extern __shared__ uint8_t smem_raw[];
  __shared__ Bars bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* as = smem;                // [8] k-block tiles of the local rows: hi 0..3, lo 4..7
  uint8_t* ring = smem + 8 * kTile;  // [kRing] k-block tiles of the global side
  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmA0);
    ptx::prefetch_tensormap(&tmA1);
    ptx::prefetch_tensormap(&tmB0);
    ptx::prefetch_tensormap(&tmB1);
    ptx::prefetch_tensormap(&tmT0);
    ptx::prefetch_tensormap(&tmT1);
    ptx::mbar_init(&bar.a_full, 1);
    ptx::mbar_init(&bar.a_empty, 1);
    for (int i = 0; i < kRing; ++i) {
      ptx::mbar_init(&bar.w_full[i], 1);
      ptx::mbar_init(&bar.w_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bar.s_full[i], 1);
      ptx::mbar_init(&bar.a_ready[i], kXfWarps / 2);
      ptx::mbar_init(&bar.l_free[i], 1);
    }
    ptx::mbar_init(&bar.x_full, 1);
    ptx::mbar_init(&bar.x_free, kXfWarps);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(&tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t tL = tmem + 256u, tX = tmem + 384u;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t st = 0, ph = 0;
      auto ring_load = [&](const CUtensorMap* tm, int x, int y) {
        ptx::mbar_wait(&bar.w_empty[st], ph ^ 1u);
        ptx::mbar_arrive_expect_tx(&bar.w_full[st], kTile);
        ptx::tma_load_3d(tm, &bar.w_full[st], ring + st * kTile, x, y, 0);
        if (++st == kRing) { st = 0; ph ^= 1u; }
      };
      auto load_m1 = [&](const Cursor& cu) {
        const CUtensorMap* ta = cu.dir ? &tmA1 : &tmA0;
        const CUtensorMap* tb = cu.dir ? &tmB1 : &tmB0;
        if (cu.c == cu.c0) {  // first chunk of the unit: its local rows (hi and lo k-blocks)
          ptx::mbar_wait(&bar.a_empty, ((uint32_t)cu.k & 1u) ^ 1u);
          ptx::mbar_arrive_expect_tx(&bar.a_full, 8 * kTile);
          for (int kb = 0; kb < 4; ++kb) ptx::tma_load_3d(ta, &bar.a_full, as + kb * kTile, kb * 32, cu.row0, 0);
          for (int kb = 0; kb < 4; ++kb)
            ptx::tma_load_3d(ta, &bar.a_full, as + (4 + kb) * kTile, p.a_lo_col[cu.dir] + kb * 32, cu.row0, 0);
        }
        for (int kb = 0; kb < 4; ++kb) ring_load(tb, kb * 32, cu.c * kChunk);
        for (int kb = 0; kb < 4; ++kb) ring_load(tb, p.b_lo_col[cu.dir] + kb * 32, cu.c * kChunk);
      };
      auto load_m2 = [&](const Cursor& cu) {
        const CUtensorMap* tt = cu.dir ? &tmT1 : &tmT0;
        for (int kb = 0; kb < 4; ++kb) ring_load(tt, cu.c * kChunk + kb * 32, 0);
        if (PRECISE)
          for (int kb = 0; kb < 4; ++kb) ring_load(tt, cu.c * kChunk + kb * 32, kD);
      };
      Cursor c1, c2;
      cur_init(c1, p);
      cur_init(c2, p);
      for (int i = 0; i < 2; ++i)
        if (c1.ok) {
          load_m1(c1);
          cur_next(c1, p);
        }
      while (c2.ok) {
        if (!LSE) load_m2(c2);
        cur_next(c2, p);
        if (c1.ok) {
          load_m1(c1);
          cur_next(c1, p);
        }
      }
    }
  } else if (warp == 1) {
    // warp-uniform schedule, MMAs issued from elect_one branches (see ffn_fused.cu)
    const uint32_t idesc = ptx::make_idesc_tf32(128, 128, 0, 0);
    const uint64_t da0 = ptx::make_smem_desc(ptx::smem_u32(as), 16, 1024, 2);
    const uint64_t dw0 = ptx::make_smem_desc(ptx::smem_u32(ring), 16, 1024, 2);
    uint32_t st = 0, ph = 0;
    auto mma_m1 = [&](const Cursor& cu) {
      const uint32_t b = (uint32_t)cu.n & 1u;
      const uint32_t tS = tmem + b * 128u;
      if (cu.c == cu.c0) {
        ptx::mbar_wait(&bar.a_full, (uint32_t)cu.k & 1u);
        ptx::tc_fence_after_sync();
      }
      const bool last = cu.c == cu.c1 - 1;
      if (LSE) {  // no M2 in front of this product: the transform warps have read chunk n - 2 out of this S buffer
        ptx::mbar_wait(&bar.a_ready[b], (((uint32_t)cu.n >> 1) & 1u) ^ 1u);
        ptx::tc_fence_after_sync();
      }
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {  // b_hi k-blocks: a_hi b_hi^T + a_lo b_hi^T
        ptx::mbar_wait(&bar.w_full[st], ph);
        ptx::tc_fence_after_sync();
        const uint64_t dh = da0 + (uint64_t)(kb * kTile16), dl = da0 + (uint64_t)((4 + kb) * kTile16);
        const uint64_t db = dw0 + (uint64_t)(st * kTile16);
        if (ptx::elect_one()) {
#pragma unroll
          for (int k8 = 0; k8 < 4; ++k8)
            ptx::mma_tf32_ss(tS, dh + (uint64_t)(k8 * 2), db + (uint64_t)(k8 * 2), idesc, (kb | k8) ? 1u : 0u);
#pragma unroll
          for (int k8 = 0; k8 < 4; ++k8) ptx::mma_tf32_ss(tS, dl + (uint64_t)(k8 * 2), db + (uint64_t)(k8 * 2), idesc, 1u);
          ptx::mma_commit(&bar.w_empty[st]);
        }
        __syncwarp();
        if (++st == kRing) { st = 0; ph ^= 1u; }
      }
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {  // b_lo k-blocks: a_hi b_lo^T
        ptx::mbar_wait(&bar.w_full[st], ph);
        ptx::tc_fence_after_sync();
        const uint64_t dh = da0 + (uint64_t)(kb * kTile16);
        const uint64_t db = dw0 + (uint64_t)(st * kTile16);
        if (ptx::elect_one()) {
#pragma unroll
          for (int k8 = 0; k8 < 4; ++k8) ptx::mma_tf32_ss(tS, dh + (uint64_t)(k8 * 2), db + (uint64_t)(k8 * 2), idesc, 1u);
          ptx::mma_commit(&bar.w_empty[st]);
          if (kb == 3) {
            ptx::mma_commit(&bar.s_full[b]);
            if (last) ptx::mma_commit(&bar.a_empty);  // the unit's local rows are not read again
          }
        }
        __syncwarp();
        if (++st == kRing) { st = 0; ph ^= 1u; }
      }
    };
    auto mma_m2 = [&](const Cursor& cu) {
      const uint32_t n = (uint32_t)cu.n, b = n & 1u;
      const uint32_t tG = tmem + b * 128u;
      ptx::mbar_wait(&bar.a_ready[b], (n >> 1) & 1u);  // the chunk's G is in tensor memory
      ptx::mbar_wait(&bar.x_free, (n & 1u) ^ 1u);      // the previous chunk's X has been added to the running sums
      ptx::tc_fence_after_sync();
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {  // B_hi k-blocks
        ptx::mbar_wait(&bar.w_full[st], ph);
        ptx::tc_fence_after_sync();
        const uint64_t db = dw0 + (uint64_t)(st * kTile16);
        if (ptx::elect_one()) {
          if (PRECISE) {
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8)
              mma_tf32_ts(tX, tL + (uint32_t)(kb * 32 + k8 * 8), db + (uint64_t)(k8 * 2), idesc, (kb | k8) ? 1u : 0u);
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8) mma_tf32_ts(tX, tG + (uint32_t)(kb * 32 + k8 * 8), db + (uint64_t)(k8 * 2), idesc, 1u);
          } else {
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8)
              mma_tf32_ts(tX, tG + (uint32_t)(kb * 32 + k8 * 8), db + (uint64_t)(k8 * 2), idesc, (kb | k8) ? 1u : 0u);
          }
          ptx::mma_commit(&bar.w_empty[st]);
          if (kb == 3) ptx::mma_commit(PRECISE ? &bar.l_free[b] : &bar.x_full);
        }
        __syncwarp();
        if (++st == kRing) { st = 0; ph ^= 1u; }
      }
      if (PRECISE) {
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {  // B_lo k-blocks
          ptx::mbar_wait(&bar.w_full[st], ph);
          ptx::tc_fence_after_sync();
          const uint64_t db = dw0 + (uint64_t)(st * kTile16);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8) mma_tf32_ts(tX, tG + (uint32_t)(kb * 32 + k8 * 8), db + (uint64_t)(k8 * 2), idesc, 1u);
            ptx::mma_commit(&bar.w_empty[st]);
            if (kb == 3) ptx::mma_commit(&bar.x_full);
          }
          __syncwarp();
          if (++st == kRing) { st = 0; ph ^= 1u; }
        }
      }
    };
    Cursor c1, c2;
    cur_init(c1, p);
    cur_init(c2, p);
    for (int i = 0; i < 2; ++i)
      if (c1.ok) {
        mma_m1(c1);
        cur_next(c1, p);
      }
    while (c2.ok) {
      if (!LSE) mma_m2(c2);
      cur_next(c2, p);
      if (c1.ok) {
        mma_m1(c1);
        cur_next(c1, p);
      }
    }
  } else {
    const int q = warp & 3;            // TMEM lane quadrant this warp may access
    const int pidx = (warp - 2) >> 2;  // 0..3: the 32-column block of X this warp accumulates
    const int g = pidx >> 1;           // transform group = S buffer
    const int sub = pidx & 1;          // which 64 of the chunk's 128 columns this warp transforms
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int rit = q * 32 + lane;     // row inside the tile
    if (LSE) {
      // forward: row sums of exp(S - 1/tau) over this warp's 64 columns of its group's chunks
      const float cmax = p.inv_tau * kLog2e;
      float acc = 0.f;
      int acc_k = -1, acc_dir = 0, acc_row = 0, acc_sp = 0;
      auto flush = [&]() {
        if (acc_k >= 0) p.partial[acc_dir][(long long)acc_row * p.slots + (acc_sp * 2 + g) * 2 + sub] = acc;
      };
      Cursor cu;
      cur_init(cu, p);
      for (; cu.ok; cur_next(cu, p)) {
        if (((uint32_t)cu.n & 1u) != (uint32_t)g) continue;
        if (acc_k != cu.k) {
          flush();
          acc = 0.f;
          acc_k = cu.k;
          acc_dir = cu.dir;
          acc_row = cu.row0 + rit;
          acc_sp = cu.sp;
        }
        ptx::mbar_wait(&bar.s_full[g], ((uint32_t)cu.n >> 1) & 1u);
        ptx::tc_fence_after_sync();
        const bool diag_chunk = cu.c * kChunk == cu.row0 + p.diag_off;
        const uint32_t tS = tmem + (uint32_t)(g * 128) + lane_base;
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          const int col0 = sub * 64 + blk * 32;
          uint32_t r[32];
          ptx::tmem_ld_32x32(tS + (uint32_t)col0, r);
          ptx::tmem_ld_wait();
          float a0 = 0.f, a1 = 0.f;
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            a0 += approx_ex2(fmaf(__uint_as_float(r[e]), p.c, -cmax));
            a1 += approx_ex2(fmaf(__uint_as_float(r[e + 1]), p.c, -cmax));
          }
          acc += a0 + a1;
          if (diag_chunk && (col0 >> 5) == q && cu.dir == 0) {  // S_ii is the same in both directions
            float dv = 0.f;
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (e == lane) dv = __uint_as_float(r[e]);
            p.diag[cu.row0 + rit] = dv * p.inv_tau;
          }
        }
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bar.a_ready[g]);  // this S buffer may be overwritten
      }
      flush();
    } else {
    float xs[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) xs[e] = 0.f;
    float rowscale = 0.f;
    int scale_k = -1;

    auto transform = [&](const Cursor& cu) {
      const uint32_t n = (uint32_t)cu.n;
      if (scale_k != cu.k) {  // scale * exp(-lse_row[i]) of this unit's row
        scale_k = cu.k;
        rowscale = p.scale * approx_ex2(-kLog2e * __ldg(p.lse_row[cu.dir] + cu.row0 + rit));
      }
      ptx::mbar_wait(&bar.s_full[g], (n >> 1) & 1u);
      ptx::tc_fence_after_sync();
      const int j0 = cu.c * kChunk;
      const float* cs = p.colscale[cu.dir] + j0;
      const bool diag_chunk = j0 == cu.row0 + p.diag_off;
      const uint32_t tS = tmem + (uint32_t)(g * 128) + lane_base;
#pragma unroll
      for (int blk = 0; blk < 4; ++blk) {
        const int col0 = sub * 64 + blk * 16;
        uint32_t r[16];
        ptx::tmem_ld_32x16(tS + (uint32_t)col0, r);
        float sc[16];
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(cs + col0 + e));
          sc[e] = v.x + rowscale;
          sc[e + 1] = v.y + rowscale;
          sc[e + 2] = v.z + rowscale;
          sc[e + 3] = v.w + rowscale;
        }
        ptx::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(approx_ex2(__uint_as_float(r[e]) * p.c) * sc[e]);
        if (diag_chunk && (col0 >> 5) == q) {  // the positives of this warp's rows lie in this block for half of its lanes
          const int de = rit - col0;
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (e == de) r[e] = __float_as_uint(__uint_as_float(r[e]) - 2.0f * p.scale);
        }
        if (PRECISE) {
          uint32_t lo[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const uint32_t hi = r[e] & 0xFFFFE000u;  // exactly what the tensor core keeps of a tf32 operand
            lo[e] = __float_as_uint(__uint_as_float(r[e]) - __uint_as_float(hi));
            r[e] = hi;
          }
          if (blk == 0 && n > 0) {  // the previous chunk's G_lo products have been read out of the shared lo buffer
            ptx::mbar_wait(&bar.l_free[(n - 1u) & 1u], ((n - 1u) >> 1) & 1u);
            ptx::tc_fence_after_sync();
          }
          tmem_st_32x16(tL + lane_base + (uint32_t)col0, lo);
        }
        tmem_st_32x16(tS + (uint32_t)col0, r);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bar.a_ready[g]);
    };
    auto drain = [&](const Cursor& cu) {
      ptx::mbar_wait(&bar.x_full, (uint32_t)cu.n & 1u);
      ptx::tc_fence_after_sync();
      uint32_t r[32];
      ptx::tmem_ld_32x32(tX + lane_base + (uint32_t)(pidx * 32), r);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bar.x_free);
#pragma unroll
      for (int e = 0; e < 32; ++e) xs[e] += __uint_as_float(r[e]);
      if (cu.c == cu.c1 - 1) {  // end of the unit: this thread's 32 columns of one output row
        float* dst = (p.atomic == 2 ? p.part[cu.dir] + cu.sp * p.part_stride : p.out[cu.dir]) + (long long)(cu.row0 + rit) * kD + pidx * 32;
        if (p.atomic == 1) {
#pragma unroll
          for (int e = 0; e < 32; ++e) atomicAdd(dst + e, xs[e]);
        } else {
          store_row32(dst, xs);
        }
#pragma unroll
        for (int e = 0; e < 32; ++e) xs[e] = 0.f;
      }
    };
    // chunk n + 1 is transformed (by its group) before chunk n is drained: T(n + 1) overlaps M2(n), D(n) falls into M1(n + 2)
    Cursor ct, cd;
    cur_init(ct, p);
    cur_init(cd, p);
    if (ct.ok && g == 0) transform(ct);
    if (ct.ok) cur_next(ct, p);
    while (cd.ok) {
      if (ct.ok && (ct.n & 1) == g) transform(ct);
      drain(cd);
      if (ct.ok) cur_next(ct, p);
      cur_next(cd, p);
    }
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem, 512);
  }
}

// Per call: colscale[d][j] = scale * exp(-lse_col[d][j]) and the transposed global unit vectors bt[d] (2 * 128, Ng) =
// [hi^T ; lo^T] (the B operand of M2 is K-major: contraction index j contiguous).
__global__ void __launch_bounds__(256)
nce_prep_kernel(const float* __restrict__ b0, const float* __restrict__ b1, int lo0, int lo1, const float* __restrict__ lse0,
                const float* __restrict__ lse1, float* __restrict__ bt0, float* __restrict__ bt1, float* __restrict__ cs0,
                float* __restrict__ cs1, int Ng, float scale) {
  __shared__ float tile[32][33];
  const int jt = Ng / 32;
  const int n_t = 2 * 8 * jt;  // direction x (hi | lo) x 4 column blocks of 32 x jt row blocks
  for (int t = blockIdx.x; t < n_t; t += gridDim.x) {
    const int dir = t / (8 * jt), rem = t - dir * 8 * jt, part = rem / jt, jb = rem - part * jt;
    const float* src = dir ? b1 : b0;
    float* dst = dir ? bt1 : bt0;
    const int lo = dir ? lo1 : lo0;
    const int col0 = (part >> 2 ? lo : 0) + (part & 3) * 32;  // column of the split row
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    __syncthreads();
    for (int r = ty; r < 32; r += 8) tile[r][tx] = src[(long long)(jb * 32 + r) * (3 * kD) + col0 + tx];
    __syncthreads();
    for (int r = ty; r < 32; r += 8) dst[(long long)((part >> 2) * kD + (part & 3) * 32 + r) * Ng + jb * 32 + tx] = tile[tx][r];
  }
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < 2 * Ng; j += gridDim.x * blockDim.x) {
    const int dir = j >= Ng, jj = j - dir * Ng;
    (dir ? cs1 : cs0)[jj] = scale * exp2f(-kLog2e * (dir ? lse1 : lse0)[jj]);
  }
}

// lse[d][i] = 1/tau + log(sum of the row's partial sums), fixed summation order
__global__ void __launch_bounds__(256)
nce_lse_finalize_kernel(const float* __restrict__ part0, const float* __restrict__ part1, float* __restrict__ lse0,
                        float* __restrict__ lse1, int Ml, int slots, float inv_tau) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * Ml; i += gridDim.x * blockDim.x) {
    const int dir = i >= Ml, row = i - dir * Ml;
    const float* pr = (dir ? part1 : part0) + (long long)row * slots;
    float s = 0.f;
    for (int k = 0; k < slots; ++k) s += pr[k];
    (dir ? lse1 : lse0)[row] = inv_tau + logf(s);
  }
}

// out[d][i] = sum over sp of part[d][sp][i], fixed order (row tiles split over more than two units)
__global__ void __launch_bounds__(256)
nce_sum_parts_kernel(const float* __restrict__ p0, const float* __restrict__ p1, float* __restrict__ o0, float* __restrict__ o1,
                     long long n4, int splits) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < 2 * n4; i += (long long)gridDim.x * blockDim.x) {
    const int dir = i >= n4;
    const long long j = i - dir * n4;
    const float4* src = reinterpret_cast<const float4*>(dir ? p1 : p0);
    float4 acc = src[j];
    for (int s = 1; s < splits; ++s) {
      const float4 v = src[(long long)s * n4 + j];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(dir ? o1 : o0)[j] = acc;
  }
}

// units = direction x row tile x column split, the split chosen so that one wave of CTAs covers the SMs
static void plan(Params& p, int64_t Ml, int64_t Ng) {
  p.row_tiles = (int)(Ml / 128);
  p.nc = (int)(Ng / 128);
  int split = kNumSMs / (2 * p.row_tiles);
  if (split < 1) split = 1;
  if (split > p.nc) split = p.nc;
  p.cps = (p.nc + split - 1) / split;
  split = (p.nc + p.cps - 1) / p.cps;
  p.units = 2 * p.row_tiles * split;
  p.atomic = split > 2 ? 2 : (split > 1 ? 1 : 0);
  p.slots = 4 * split;
  // e3 = [hi | lo | hi], f3 = [hi | hi | lo] (xm_l2norm_split_fwd_f32, which = 0 / 1)
  p.a_lo_col[0] = kD;      // local e
  p.b_lo_col[0] = 2 * kD;  // global f
  p.a_lo_col[1] = 2 * kD;  // local f
  p.b_lo_col[1] = kD;      // global e
}

static TensorView3 view2(const void* ptr, long long cols, long long rows) {
  return TensorView3{ptr, {(unsigned long long)cols, (unsigned long long)rows, 1ull},
                     {(unsigned long long)cols * 4ull, (unsigned long long)cols * (unsigned long long)rows * 4ull}};
}

}  // namespace nce
}  // namespace xm

using namespace xm;

extern "C" {

int xm_infonce_bwd_fused_supported(int64_t Ml, int64_t Ng, int64_t D, int64_t diag_off) {
  return D == nce::kD && Ml > 0 && Ng > 0 && Ml % 128 == 0 && Ng % 128 == 0 && diag_off >= 0 && diag_off % 128 == 0 &&
         diag_off + Ml <= Ng && Ng < ((int64_t)1 << 30);
}

int64_t xm_infonce_bwd_fused_workspace(int64_t Ml, int64_t Ng, int64_t D) {
  if (Ml <= 0 || Ng <= 0 || Ml % 128 || Ng % 128) return 0;
  nce::Params p{};
  nce::plan(p, Ml, Ng);
  return 2 * Ng + 4 * D * Ng + (p.atomic == 2 ? 2 * (int64_t)(p.slots / 4) * Ml * D : 0);
}

int xm_infonce_bwd_fused_f32(const float* e3, const float* f3, const float* e3_all, const float* f3_all, const float* lse_ef,
                             const float* lse_fe, const float* lse_ef_all, const float* lse_fe_all, float* de, float* df,
                             int64_t Ml, int64_t Ng, int64_t D, float inv_tau, int64_t diag_off, float coef, int precise,
                             float* workspace, void* stream) {
  if (!e3 || !f3 || !e3_all || !f3_all || !lse_ef || !lse_fe || !lse_ef_all || !lse_fe_all || !de || !df || !workspace)
    return XM_ERR_INVALID;
  if (!xm_infonce_bwd_fused_supported(Ml, Ng, D, diag_off)) return XM_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(de) | reinterpret_cast<uintptr_t>(df) | reinterpret_cast<uintptr_t>(workspace)) & 31)
    return XM_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  float* cs0 = workspace;       // direction 0 (dE): columns are the global f rows -> lse_fe_all
  float* cs1 = workspace + Ng;  // direction 1 (dF): lse_ef_all
  float* bt0 = workspace + 2 * Ng;             // f^T [hi; lo]
  float* bt1 = bt0 + 2 * nce::kD * Ng;         // e^T [hi; lo]
  nce::Params p{};
  nce::plan(p, Ml, Ng);
  p.lse_row[0] = lse_ef;
  p.lse_row[1] = lse_fe;
  p.colscale[0] = cs0;
  p.colscale[1] = cs1;
  p.out[0] = de;
  p.out[1] = df;
  p.diag_off = (int)diag_off;
  p.c = inv_tau * nce::kLog2e;
  p.scale = precise ? coef : coef * nce::kTruncComp;
  {
    const int tiles = 2 * 8 * (int)(Ng / 32);
    const int blocks = tiles < kNumSMs * 8 ? tiles : kNumSMs * 8;
    nce::nce_prep_kernel<<<blocks, 256, 0, st>>>(f3_all, e3_all, 2 * nce::kD, nce::kD, lse_fe_all, lse_ef_all, bt0, bt1, cs0, cs1,
                                                 (int)Ng, p.scale);
    int rc = check_launch();
    if (rc != XM_OK) return rc;
  }
  if (p.atomic == 2) {
    p.part[0] = bt1 + 2 * nce::kD * Ng;
    p.part_stride = Ml * nce::kD;
    p.part[1] = p.part[0] + (p.slots / 4) * p.part_stride;
  }
  if (p.atomic == 1) {
    if (cudaMemsetAsync(de, 0, (size_t)Ml * nce::kD * 4, st) != cudaSuccess ||
        cudaMemsetAsync(df, 0, (size_t)Ml * nce::kD * 4, st) != cudaSuccess) {
      g_last_cuda_error = (int)cudaGetLastError();
      return XM_ERR_LAUNCH;
    }
  }
  CUtensorMap ma0, ma1, mb0, mb1, mt0, mt1;
  int rc = encode_tmap(&ma0, nce::view2(e3, 3 * D, Ml), 32, 128, 0);
  if (rc == XM_OK) rc = encode_tmap(&ma1, nce::view2(f3, 3 * D, Ml), 32, 128, 0);
  if (rc == XM_OK) rc = encode_tmap(&mb0, nce::view2(f3_all, 3 * D, Ng), 32, 128, 0);
  if (rc == XM_OK) rc = encode_tmap(&mb1, nce::view2(e3_all, 3 * D, Ng), 32, 128, 0);
  if (rc == XM_OK) rc = encode_tmap(&mt0, nce::view2(bt0, Ng, 2 * D), 32, 128, 0);
  if (rc == XM_OK) rc = encode_tmap(&mt1, nce::view2(bt1, Ng, 2 * D), 32, 128, 0);
  if (rc != XM_OK) return rc;
  const int ctas = p.units < kNumSMs ? p.units : kNumSMs;
  cudaError_t e;
  if (precise) {
    e = cudaFuncSetAttribute(nce::nce_kernel<nce::MODE_PRECISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, nce::kSmem);
    if (e == cudaSuccess) nce::nce_kernel<nce::MODE_PRECISE><<<ctas, nce::kThreads, nce::kSmem, st>>>(ma0, ma1, mb0, mb1, mt0, mt1, p);
  } else {
    e = cudaFuncSetAttribute(nce::nce_kernel<nce::MODE_SINGLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, nce::kSmem);
    if (e == cudaSuccess) nce::nce_kernel<nce::MODE_SINGLE><<<ctas, nce::kThreads, nce::kSmem, st>>>(ma0, ma1, mb0, mb1, mt0, mt1, p);
  }
  if (e != cudaSuccess) {
    g_last_cuda_error = (int)e;
    return XM_ERR_LAUNCH;
  }
  rc = check_launch();
  if (rc != XM_OK || p.atomic != 2) return rc;
  const long long n4 = Ml * nce::kD / 4;
  nce::nce_sum_parts_kernel<<<(int)((2 * n4 + 255) / 256 < 1184 ? (2 * n4 + 255) / 256 : 1184), 256, 0, st>>>(p.part[0], p.part[1], de, df, n4,
                                                                                                              p.slots / 4);
  return check_launch();
}

int64_t xm_infonce_lse_fused_workspace(int64_t Ml, int64_t Ng) {
  if (Ml <= 0 || Ng <= 0 || Ml % 128 || Ng % 128) return 0;
  nce::Params p{};
  nce::plan(p, Ml, Ng);
  return 2 * Ml * p.slots;
}

int xm_infonce_lse_fused_f32(const float* e3, const float* f3, const float* e3_all, const float* f3_all, float* lse_ef,
                             float* lse_fe, float* diag, int64_t Ml, int64_t Ng, int64_t D, float inv_tau, int64_t diag_off,
                             float* workspace, void* stream) {
  if (!e3 || !f3 || !e3_all || !f3_all || !lse_ef || !lse_fe || !diag || !workspace) return XM_ERR_INVALID;
  if (!xm_infonce_bwd_fused_supported(Ml, Ng, D, diag_off)) return XM_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  nce::Params p{};
  nce::plan(p, Ml, Ng);
  p.partial[0] = workspace;
  p.partial[1] = workspace + Ml * p.slots;
  p.diag = diag;
  p.diag_off = (int)diag_off;
  p.inv_tau = inv_tau;
  p.c = inv_tau * nce::kLog2e;
  CUtensorMap ma0, ma1, mb0, mb1;
  int rc = encode_tmap(&ma0, nce::view2(e3, 3 * D, Ml), 32, 128, 0);
  if (rc == XM_OK) rc = encode_tmap(&ma1, nce::view2(f3, 3 * D, Ml), 32, 128, 0);
  if (rc == XM_OK) rc = encode_tmap(&mb0, nce::view2(f3_all, 3 * D, Ng), 32, 128, 0);
  if (rc == XM_OK) rc = encode_tmap(&mb1, nce::view2(e3_all, 3 * D, Ng), 32, 128, 0);
  if (rc != XM_OK) return rc;
  // every (row, slot) of the workspace is written: each unit's 16 transform warps cover 2 groups x 2 column halves.  A
  // group that gets no chunk of a unit (single-chunk units) leaves its slots untouched: clear them first.
  if (cudaMemsetAsync(workspace, 0, (size_t)(2 * Ml * p.slots) * 4, st) != cudaSuccess) {
    g_last_cuda_error = (int)cudaGetLastError();
    return XM_ERR_LAUNCH;
  }
  const int ctas = p.units < kNumSMs ? p.units : kNumSMs;
  cudaError_t e = cudaFuncSetAttribute(nce::nce_kernel<nce::MODE_LSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, nce::kSmem);
  if (e != cudaSuccess) {
    g_last_cuda_error = (int)e;
    return XM_ERR_LAUNCH;
  }
  nce::nce_kernel<nce::MODE_LSE><<<ctas, nce::kThreads, nce::kSmem, st>>>(ma0, ma1, mb0, mb1, mb0, mb1, p);
  rc = check_launch();
  if (rc != XM_OK) return rc;
  const int blocks = (int)((2 * Ml + 255) / 256);
  nce::nce_lse_finalize_kernel<<<blocks, 256, 0, st>>>(p.partial[0], p.partial[1], lse_ef, lse_fe, (int)Ml, p.slots, inv_tau);
  return check_launch();
}

}  // extern "C"
