// Fused feed-forward branch of the pre-norm transformer block for sm_100a (EEG_CODE/enhanced_models_v4.py:79-80,
// 102-105: linear2(Dropout(GELU(linear1(x))))), d_model = 128, hidden = a multiple of 128 (512 in the reference).
// The (rows x hidden) intermediate -- 2.1 GB per block at batch 4096 x 250 tokens, which the unfused path wrote or
// read 11 times per block -- never reaches HBM in the forward and is written exactly once (as the two operands of
// the weight gradients) in the backward.
//
//   forward   y = tf32(Dropout(act(x W1^T + b1))) W2^T + b2                                         xm_ffn_fused_fwd_f32
//   backward  given x, dY:  H = x W1^T + b1 (recomputed),  G = dY W2,
//             A  = Dropout(act(H))                (rows, hidden)   operand of dW2 = dY^T A
//             dH = G * act'(H) * mask / (1 - p)   (rows, hidden)   operand of dW1 = dH^T x
//             dX = dH W1,   db1 partial column sums of dH                                          xm_ffn_fused_dgrad_f32
//   (tf32 operands are made by the tensor core's truncation of values pre-scaled by 1 + 0.7213 * 2^-11: see kTruncComp)
//
// One persistent CTA per SM walks 128-row tiles; per tile the hidden axis is processed in chunks of 128 units:
//   warp 0      TMA producer: the x (and dY) tile of the tile, then a ring of 16 KB weight k-blocks (from L2)
//   warp 1      one thread issues every tcgen05.mma (kind::tf32, M = N = 128, fp32 accumulators in TMEM)
//   warps 2-17  16 transform warps (4 TMEM lane quadrants x 4 blocks of 32 columns), all on every chunk:
//               tcgen05.ld the chunk's accumulator, bias + activation + dropout in registers (packed f32x2
//               arithmetic), tcgen05.st the result back IN PLACE, where it is the A operand of the next product
// forward MMA order:   M1(0) M1(1) | M2(0) M1(2) | M2(1) M1(3) | M2(2) | M2(3)          M1: H_c = x W1_c^T
//                                                                                        M2: Y += A_c W2[:, c]^T
// backward MMA order:  M1(0) M3(0) M1(1) | M4(0) M3(1) M1(2) | M4(1) M3(2) M1(3) | ...   M3: G_c = dY W2t_c^T
//                                                                                        M4: dX += dH_c W1t[:, c]^T
// tcgen05.mma executes in issue order, so re-using a TMEM buffer between MMAs needs no barrier; every M2 / M4 waits
// for the transform warps of its chunk.  TMEM: H0 [0,128) H1 [128,256); forward Y [256,384); backward G [256,384),
// dX [384,512).  A / dH / y / dX leave the registers as 256-bit global stores: each lane owns one 128-byte line of a
// row, so every store instruction writes 32 complete sectors.
//
// Dropout mask: a pure function of (seed, row, hidden unit), identical in both kernels and in xm_ffn_fused_mask_u8:
// per (row, 32-unit group g) a stream seed h0 = mix(rs(row) + (g + 1) * 0x9E3779B1), rs = high word of
// hash_u64(row, seed); unit g*32 + j takes LCG state h_{j+1} and is kept iff h_{j+1} >= p * 2^32.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/xmodal_b200.h"
#include "gemm_engine.cuh"

namespace xm {
namespace ffn {

constexpr int kD = 128, kChunk = 128, kMaxChunks = 8;
constexpr int kTile = 16384;  // 128 rows x 32 fp32, SWIZZLE_128B
constexpr int kXfWarps = 16;
constexpr int kThreads = 64 + 32 * kXfWarps;
constexpr int kFwdThreads = kThreads + 128;  // forward: + 4 output warps (one per TMEM lane quadrant)
constexpr int kFwdRing = 5, kBwdRing = 9;
constexpr int kBiasBytes = (kMaxChunks * kChunk + kD) * 4;  // b1 (+ b2) staged in shared memory
constexpr int kFwdSmem = 2 * 4 * kTile + kFwdRing * kTile + kBiasBytes + 1024;
constexpr int kStage = 4096;  // per-warp staging buffer of the data-gradient kernel: one 32 x 32 fp32 block
constexpr int kBwdSmem = kBwdRing * kTile + kXfWarps * kStage + kBiasBytes + 1024;
constexpr uint32_t kGold = 0x9E3779B1u;
constexpr uint32_t kLcgA = 747796405u, kLcgC = 2891336453u;
enum : int { OP_M1 = 0, OP_M2 = 1, OP_M3 = 2, OP_M4 = 3 };

struct Params {
  long long M;
  int tiles, nc, hidden, act;
  const float* b1;
  const float* b2;
  float* y;        // forward: (M, 128)
  float* dx;       // backward: (M, 128)
  float* db1_part; // backward: (tiles * 4, hidden) partial column sums of dH, one row per (tile, lane quadrant)
  float dscale;
  uint32_t thr;
  unsigned long long seed;
  long long* trace;  // debug (xm_debug_set_ffn_trace): CTA 0 logs per-op clocks, 3 roles x 8192 slots
  int dbg;           // tracing instance only, timing experiments (results are then WRONG): bit 0 = the transform warps
                     // skip their arithmetic, bit 1 = the producer loads each ring stage once and only re-signals it
                     // (forward), bit 2 = the data-gradient kernel does not store A / dH
};

// mbarrier wait that, in the tracing instance of a kernel, adds the cycles it spent to `acc`
template <bool TRACE>
XM_DEVICE void twait(uint64_t* bar, uint32_t parity, long long& acc) {
  if (TRACE) {
    const long long t0 = clock64();
    ptx::mbar_wait(bar, parity);
    acc += clock64() - t0;
  } else {
    ptx::mbar_wait(bar, parity);
  }
}

XM_DEVICE uint32_t mask_mix(uint32_t x) {
  x *= 0x7FEB352Du;
  x ^= x >> 15;
  x *= 0x846CA68Bu;
  return x ^ (x >> 16);
}
XM_DEVICE uint32_t row_seed(unsigned long long row, unsigned long long seed) { return (uint32_t)(hash_u64(row, seed) >> 32); }
XM_DEVICE uint32_t group_seed(uint32_t rs, int g) { return mask_mix(rs + (uint32_t)(g + 1) * kGold); }

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

// The MMA warp runs its schedule WARP-UNIFORMLY (all 32 lanes wait on the barriers and compute the operand
// descriptors, which the compiler then keeps in uniform registers) and issues each k-block's four MMAs + commit from
// an `elect_one` branch: in SASS the four UTCHMMA follow each other without an instruction in between.  With the
// whole schedule inside `if (lane == 0)` every descriptor went through vector registers and R2UR moves, ~125 cycles
// of issue per MMA against 76 for the MMA itself (profiles/r2_ffn_trace_v2.json, profiles/r2_mma_rate_probe.txt).
constexpr uint32_t kTile16 = kTile >> 4;  // descriptor start-address units (16 B) per 16 KB tile

// The per-tile MMA schedule, written once into shared memory (kind << 8 | chunk).
XM_DEVICE int build_fwd_ops(int* ops, int nc) {
  int n = 0;
  ops[n++] = OP_M1 << 8;
  if (nc > 1) ops[n++] = (OP_M1 << 8) | 1;
  for (int c = 0; c < nc; ++c) {
    ops[n++] = (OP_M2 << 8) | c;
    if (c + 2 < nc) ops[n++] = (OP_M1 << 8) | (c + 2);
  }
  return n;
}
XM_DEVICE int build_bwd_ops(int* ops, int nc) {
  int n = 0;
  ops[n++] = OP_M1 << 8;
  ops[n++] = OP_M3 << 8;
  if (nc > 1) ops[n++] = (OP_M1 << 8) | 1;
  for (int c = 0; c < nc; ++c) {
    ops[n++] = (OP_M4 << 8) | c;
    if (c + 1 < nc) ops[n++] = (OP_M3 << 8) | (c + 1);
    if (c + 2 < nc) ops[n++] = (OP_M1 << 8) | (c + 2);
  }
  return n;
}

// ---- packed activation arithmetic (f32x2: one FFMA2 / FMUL2 per two elements; these transforms are issue bound)
XM_DEVICE float2 f2(float a, float b) { return make_float2(a, b); }
XM_DEVICE float2 splat(float a) { return make_float2(a, a); }
// Values written back to TMEM (and the A / dH operands written for the weight gradients) are NOT rounded to tf32:
// the tensor core truncates them, and their scale carries (1 + 0.7213 * 2^-11), the mean relative truncation loss
// over a binade, so the truncation is zero-mean like round-to-nearest (attention_fused.cu uses the same device).
constexpr float kTruncComp = 1.0f + 0.7213f / 2048.0f;
// cs * (standard normal cdf, pdf) of two values (Abramowitz & Stegun 26.2.17, as xm_common.cuh:normal_cdf_pdf)
XM_DEVICE void normal_cdf_pdf2(float2 x, float cs, float2& cdf, float2& pdf) {
  const float2 e = __fmul2_rn(__fmul2_rn(x, x), splat(-0.72134752f));
  pdf = __fmul2_rn(f2(approx_ex2(e.x), approx_ex2(e.y)), splat(0.3989422804f * cs));
  const float2 d = __ffma2_rn(f2(fabsf(x.x), fabsf(x.y)), splat(0.2316419f), splat(1.0f));
  const float2 t = f2(approx_rcp(d.x), approx_rcp(d.y));
  float2 poly = __ffma2_rn(t, splat(1.330274429f), splat(-1.821255978f));
  poly = __ffma2_rn(t, poly, splat(1.781477937f));
  poly = __ffma2_rn(t, poly, splat(-0.356563782f));
  poly = __ffma2_rn(t, poly, splat(0.319381530f));
  const float2 q = __fmul2_rn(pdf, __fmul2_rn(t, poly));           // cs * (1 - Phi(|x|))
  const float2 h = __fadd2_rn(splat(0.5f * cs), f2(-q.x, -q.y));   // cs * (Phi(|x|) - 0.5)
  cdf = __fadd2_rn(f2(copysignf(h.x, x.x), copysignf(h.y, x.y)), splat(0.5f * cs));
}

// Forward transform of 32 accumulator columns of one row, in place: cs * Dropout(act(acc + bias)), cs = truncation
// compensation (* 1 / (1 - p) with dropout).  `bias`: shared memory.
template <bool DROP>
XM_DEVICE void fwd_transform(uint32_t (&r)[32], const float* bias, int act, float cs, uint32_t thr, uint32_t h) {
#pragma unroll
  for (int e = 0; e < 32; e += 4) {
    const float4 bb = *reinterpret_cast<const float4*>(bias + e);
    float2 v0 = __fadd2_rn(f2(__uint_as_float(r[e]), __uint_as_float(r[e + 1])), f2(bb.x, bb.y));
    float2 v1 = __fadd2_rn(f2(__uint_as_float(r[e + 2]), __uint_as_float(r[e + 3])), f2(bb.z, bb.w));
    if (act == XM_ACT_GELU) {
      float2 c0, c1, p0, p1;
      normal_cdf_pdf2(v0, cs, c0, p0);
      normal_cdf_pdf2(v1, cs, c1, p1);
      v0 = __fmul2_rn(v0, c0);
      v1 = __fmul2_rn(v1, c1);
    } else {
      v0 = __fmul2_rn(f2(fmaxf(v0.x, 0.f), fmaxf(v0.y, 0.f)), splat(cs));
      v1 = __fmul2_rn(f2(fmaxf(v1.x, 0.f), fmaxf(v1.y, 0.f)), splat(cs));
    }
    if (DROP) {
      h = h * kLcgA + kLcgC;
      v0.x = h >= thr ? v0.x : 0.f;
      h = h * kLcgA + kLcgC;
      v0.y = h >= thr ? v0.y : 0.f;
      h = h * kLcgA + kLcgC;
      v1.x = h >= thr ? v1.x : 0.f;
      h = h * kLcgA + kLcgC;
      v1.y = h >= thr ? v1.y : 0.f;
    }
    r[e] = __float_as_uint(v0.x);
    r[e + 1] = __float_as_uint(v0.y);
    r[e + 2] = __float_as_uint(v1.x);
    r[e + 3] = __float_as_uint(v1.y);
  }
}

// Backward transform, in place: rh = H (pre-bias) -> A = cs * Dropout(act(.)), rg = G -> dH = cs * G * act'(.) * mask.
template <bool DROP>
XM_DEVICE void bwd_transform(uint32_t (&rh)[32], uint32_t (&rg)[32], const float* bias, int act, float cs, uint32_t thr, uint32_t h) {
#pragma unroll
  for (int e = 0; e < 32; e += 2) {
    const float2 bb = *reinterpret_cast<const float2*>(bias + e);
    const float2 v = __fadd2_rn(f2(__uint_as_float(rh[e]), __uint_as_float(rh[e + 1])), bb);
    float2 a, d;  // cs * act(v), cs * act'(v)
    if (act == XM_ACT_GELU) {
      float2 cdf, pdf;
      normal_cdf_pdf2(v, cs, cdf, pdf);
      a = __fmul2_rn(v, cdf);
      d = __ffma2_rn(v, pdf, cdf);
    } else {
      a = __fmul2_rn(f2(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f)), splat(cs));
      d = f2(v.x > 0.f ? cs : 0.f, v.y > 0.f ? cs : 0.f);
    }
    float2 g = __fmul2_rn(f2(__uint_as_float(rg[e]), __uint_as_float(rg[e + 1])), d);
    if (DROP) {
      h = h * kLcgA + kLcgC;
      const bool k0 = h >= thr;
      h = h * kLcgA + kLcgC;
      const bool k1 = h >= thr;
      a = f2(k0 ? a.x : 0.f, k1 ? a.y : 0.f);
      g = f2(k0 ? g.x : 0.f, k1 ? g.y : 0.f);
    }
    rh[e] = __float_as_uint(a.x);
    rh[e + 1] = __float_as_uint(a.y);
    rg[e] = __float_as_uint(g.x);
    rg[e + 1] = __float_as_uint(g.y);
  }
}

// 32 consecutive fp32 of one row (128 B, one full line per lane) -> global memory as four 256-bit stores: every
// store instruction writes 32 complete sectors (16-B stores would write each sector in two halves).
XM_DEVICE void store_row32(float* dst, const uint32_t (&r)[32]) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 8 * j), "r"(r[8 * j]), "r"(r[8 * j + 1]),
                 "r"(r[8 * j + 2]), "r"(r[8 * j + 3]), "r"(r[8 * j + 4]), "r"(r[8 * j + 5]), "r"(r[8 * j + 6]), "r"(r[8 * j + 7])
                 : "memory");
}

// 32 x 32 fp32 block of this warp (one row per lane) -> the warp's 4 KB swizzled staging buffer -> one asynchronous
// TMA store (rows / columns outside the tensor are clipped).  The warp only waits until the PREVIOUS store has read
// the buffer out of shared memory, never for the global write.
XM_DEVICE void stage_block(const CUtensorMap* tm, uint8_t* sb, int lane, const uint32_t (&r)[32], int col, int row0) {
  if (lane == 0) ptx::bulk_wait_read<0>();
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<uint4*>(sb + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
  ptx::fence_proxy_async();
  __syncwarp();
  if (lane == 0) {
    ptx::tma_store_3d(tm, sb, col, row0, 0);
    ptx::bulk_commit();
  }
}

struct FwdBars {
  uint64_t x_full[2], x_empty[2];
  uint64_t w_full[kFwdRing], w_empty[kFwdRing];
  uint64_t h_full[2], a_ready[2];  // per H buffer (a barrier may only have waiters that consume EVERY phase: a waiter
                                   // that skips phases aliases on the parity -- all 16 transform warps wait on these)
  uint64_t y_full[2], y_free[2];   // per Y accumulator (tile parity); waited on by the four output warps / the MMA warp
};

// Forward.  The MMA warp runs ONE software pipeline over the CTA's whole chunk sequence n = 0, 1, ... (tile = n / nc,
// chunk = n % nc), crossing tile boundaries:  M1(0) M1(1) | M2(0) M1(2) | M2(1) M1(3) | ...  H buffer = n & 1.
// ALL 16 transform warps work on every chunk (4 lane quadrants x 4 blocks of 32 columns): with two H buffers the
// transform of chunk n has exactly the time of [M2(n - 1) M1(n + 1)] -- M1(n) ends where that slot starts and M2(n) opens
// the next one -- so its LATENCY, not its throughput, decides whether the MMA warp waits.  (Two groups of 8 warps
// alternating chunks had the same throughput but never overlapped: each group's chunk took ~6.5k clk of a 4.9k slot,
// profiles/r2_ffn_fwd_trace_v4.json; ncu: tensor pipe 35 %, issue 45 %, nothing saturated.)
// The tile's output leaves through FOUR EXTRA WARPS and two Y accumulators (tile parity): when the transform warps
// wrote y themselves every tile boundary cost ~8k of 23.7k clk -- M2 of the last chunk had to finish, the row-per-lane
// global stores block their warp, and the next tile's first chunk waited behind both (trace of that version).
template <bool TRACE>
__global__ void __launch_bounds__(kFwdThreads, 1)
ffn_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
               const __grid_constant__ CUtensorMap tmW2, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ FwdBars bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* xs = smem;                    // [2][4] x k-block tiles
  uint8_t* ring = smem + 2 * 4 * kTile;  // [kFwdRing] weight k-block tiles
  float* sb1 = reinterpret_cast<float*>(ring + kFwdRing * kTile);  // [hidden] + [128]: b1, b2
  float* sb2 = sb1 + p.hidden;
  for (int i = threadIdx.x; i < p.hidden + kD; i += blockDim.x) sb1[i] = i < p.hidden ? p.b1[i] : p.b2[i - p.hidden];
  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmX);
    ptx::prefetch_tensormap(&tmW1);
    ptx::prefetch_tensormap(&tmW2);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bar.x_full[i], 1);
      ptx::mbar_init(&bar.x_empty[i], 1);
      ptx::mbar_init(&bar.h_full[i], 1);
      ptx::mbar_init(&bar.a_ready[i], kXfWarps);
      ptx::mbar_init(&bar.y_full[i], 1);
      ptx::mbar_init(&bar.y_free[i], 4);
    }
    for (int i = 0; i < kFwdRing; ++i) {
      ptx::mbar_init(&bar.w_full[i], 1);
      ptx::mbar_init(&bar.w_empty[i], 1);
    }
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
  const int my_tiles = p.tiles > (int)blockIdx.x ? (p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int N = my_tiles * p.nc;  // chunks this CTA processes

  if (warp == 0) {
    if (lane == 0) {
      uint32_t st = 0, ph = 0;  // ring stage and its phase parity
      long long tw = 0, tx = 0;
      int tn = 0;
      auto log = [&](int kind, int n, long long t0) {
        if (TRACE && blockIdx.x == 0 && tn < 8192 - 6) {
          long long* t = p.trace + tn;
          t[0] = kind; t[1] = n; t[2] = t0; t[3] = clock64(); t[4] = tw; t[5] = tx;
          tn += 6; tw = 0; tx = 0;
        }
      };
      auto load_m1 = [&](int n) {
        const long long t0 = TRACE ? clock64() : 0;
        const int it = n / p.nc, c = n - it * p.nc;
        if (c == 0) {  // first use of this tile's x
          const int xb = it & 1, tile = (int)blockIdx.x + it * (int)gridDim.x;
          twait<TRACE>(&bar.x_empty[xb], (((uint32_t)it >> 1) & 1u) ^ 1u, tx);
          ptx::mbar_arrive_expect_tx(&bar.x_full[xb], 4 * kTile);
          for (int kb = 0; kb < 4; ++kb)
            ptx::tma_load_3d(&tmX, &bar.x_full[xb], xs + (xb * 4 + kb) * kTile, kb * 32, tile * 128, 0);
        }
        for (int kb = 0; kb < 4; ++kb) {
          twait<TRACE>(&bar.w_empty[st], ph ^ 1u, tw);
          if (TRACE && (p.dbg & 2) && (ph || n > 1)) {
            ptx::mbar_arrive(&bar.w_full[st]);
          } else {
            ptx::mbar_arrive_expect_tx(&bar.w_full[st], kTile);
            ptx::tma_load_3d(&tmW1, &bar.w_full[st], ring + st * kTile, kb * 32, c * kChunk, 0);  // W1[128c.., 32kb..]
          }
          if (++st == kFwdRing) { st = 0; ph ^= 1u; }
        }
        log(OP_M1, n, t0);
      };
      auto load_m2 = [&](int n) {
        const long long t0 = TRACE ? clock64() : 0;
        const int c = n % p.nc;
        for (int kb = 0; kb < 4; ++kb) {
          twait<TRACE>(&bar.w_empty[st], ph ^ 1u, tw);
          if (TRACE && (p.dbg & 2)) {
            ptx::mbar_arrive(&bar.w_full[st]);
          } else {
            ptx::mbar_arrive_expect_tx(&bar.w_full[st], kTile);
            ptx::tma_load_3d(&tmW2, &bar.w_full[st], ring + st * kTile, c * kChunk + kb * 32, 0, 0);  // W2[:, 128c + 32kb..]
          }
          if (++st == kFwdRing) { st = 0; ph ^= 1u; }
        }
        log(OP_M2, n, t0);
      };
      if (N > 0) load_m1(0);
      if (N > 1) load_m1(1);
      for (int n = 0; n < N; ++n) {
        load_m2(n);
        if (n + 2 < N) load_m1(n + 2);
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t idesc = ptx::make_idesc_tf32(128, 128, 0, 0);
      const uint64_t dx0 = ptx::make_smem_desc(ptx::smem_u32(xs), 16, 1024, 2);    // x tile 0, k-block 0
      const uint64_t dw0 = ptx::make_smem_desc(ptx::smem_u32(ring), 16, 1024, 2);  // ring stage 0
      uint32_t st = 0, ph = 0;  // ring stage and its phase parity
      long long tw = 0, ta = 0;
      int tn = 0;
      auto log = [&](int kind, int n, long long t0) {
        if (TRACE && blockIdx.x == 0 && lane == 0 && tn < 8192 - 6) {
          long long* t = p.trace + 8192 + tn;
          t[0] = kind; t[1] = n; t[2] = t0; t[3] = clock64(); t[4] = tw; t[5] = ta;
          tn += 6; tw = 0; ta = 0;
        }
      };
      auto mma_m1 = [&](int n) {
        const long long t0 = TRACE ? clock64() : 0;
        const int it = n / p.nc, c = n - it * p.nc, xb = it & 1, b = n & 1;
        const uint32_t tH = tmem + (uint32_t)(b * 128);
        if (c == 0) {
          twait<TRACE>(&bar.x_full[xb], ((uint32_t)it >> 1) & 1u, ta);
          ptx::tc_fence_after_sync();
        }
        const uint64_t dxa = dx0 + (uint64_t)((uint32_t)xb * 4u * kTile16);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          twait<TRACE>(&bar.w_full[st], ph, tw);
          ptx::tc_fence_after_sync();
          const uint64_t da = dxa + (uint64_t)(kb * kTile16), db = dw0 + (uint64_t)(st * kTile16);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8)
              ptx::mma_tf32_ss(tH, da + (uint64_t)(k8 * 2), db + (uint64_t)(k8 * 2), idesc, (kb | k8) ? 1u : 0u);
            ptx::mma_commit(&bar.w_empty[st]);
            if (kb == 3) {
              ptx::mma_commit(&bar.h_full[b]);
              if (c == p.nc - 1) ptx::mma_commit(&bar.x_empty[xb]);
            }
          }
          __syncwarp();
          if (++st == kFwdRing) { st = 0; ph ^= 1u; }
        }
        log(OP_M1, n, t0);
      };
      auto mma_m2 = [&](int n) {
        const long long t0 = TRACE ? clock64() : 0;
        const int it = n / p.nc, c = n - it * p.nc, b = n & 1;
        const uint32_t tH = tmem + (uint32_t)(b * 128);
        const uint32_t tY = tmem + 256u + (uint32_t)((it & 1) * 128);
        twait<TRACE>(&bar.a_ready[b], ((uint32_t)n >> 1) & 1u, ta);  // the transform warps have written A over H
        ptx::tc_fence_after_sync();
        if (c == 0) {
          twait<TRACE>(&bar.y_free[it & 1], (((uint32_t)it >> 1) & 1u) ^ 1u, ta);  // tile it - 2 has been read out of this Y
          ptx::tc_fence_after_sync();
        }
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          twait<TRACE>(&bar.w_full[st], ph, tw);
          ptx::tc_fence_after_sync();
          const uint64_t db = dw0 + (uint64_t)(st * kTile16);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8)
              mma_tf32_ts(tY, tH + (uint32_t)(kb * 32 + k8 * 8), db + (uint64_t)(k8 * 2), idesc, (kb | k8) ? 1u : (c ? 1u : 0u));
            ptx::mma_commit(&bar.w_empty[st]);
            if (kb == 3 && c == p.nc - 1) ptx::mma_commit(&bar.y_full[it & 1]);  // the tile's output is complete
          }
          __syncwarp();
          if (++st == kFwdRing) { st = 0; ph ^= 1u; }
        }
        log(OP_M2, n, t0);
      };
      if (N > 0) mma_m1(0);
      if (N > 1) mma_m1(1);
      for (int n = 0; n < N; ++n) {
        mma_m2(n);
        if (n + 2 < N) mma_m1(n + 2);
      }
    }
  } else if (warp < 2 + kXfWarps) {
    const int q = warp & 3;            // TMEM lane quadrant this warp may access
    const int cb = (warp - 2) >> 2;    // which 32 of the chunk's 128 columns
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const float cs = kTruncComp * p.dscale;
    int tn = 0;
    for (int n = 0; n < N; ++n) {
      const int it = n / p.nc, c = n - it * p.nc, g = n & 1;
      const long long row = ((long long)blockIdx.x + (long long)it * gridDim.x) * 128 + q * 32 + lane;
      const long long t0 = TRACE ? clock64() : 0;
      ptx::mbar_wait(&bar.h_full[g], ((uint32_t)n >> 1) & 1u);
      const long long t1 = TRACE ? clock64() : 0;
      ptx::tc_fence_after_sync();
      {
        const int col0 = cb * 32;
        const uint32_t addr = tmem + (uint32_t)(g * 128 + col0) + lane_base;
        uint32_t r[32];
        ptx::tmem_ld_32x32(addr, r);
        ptx::tmem_ld_wait();
        const float* bias = sb1 + c * kChunk + col0;
        if (TRACE && (p.dbg & 1)) {
        } else if (p.thr) {
          const uint32_t rs = row_seed((unsigned long long)row, p.seed);
          fwd_transform<true>(r, bias, p.act, cs, p.thr, group_seed(rs, (c * kChunk + col0) >> 5));
        } else {
          fwd_transform<false>(r, bias, p.act, cs, 0u, 0u);
        }
        ptx::tmem_st_32x32(addr, r);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bar.a_ready[g]);
      if (TRACE && blockIdx.x == 0 && q == 0 && cb == 0 && lane == 0 && tn < 4096 - 6) {
        long long* t = p.trace + 16384 + tn;
        t[0] = 4; t[1] = n; t[2] = t0; t[3] = t1; t[4] = clock64(); t[5] = 0;
        tn += 6;
      }
    }
  } else {
    // ---- output warps: y = Y + b2, one TMEM lane quadrant (32 rows) each, 32 columns at a time
    const int q = warp & 3;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    for (int it = 0; it < my_tiles; ++it) {
      const long long row = ((long long)blockIdx.x + (long long)it * gridDim.x) * 128 + q * 32 + lane;
      const uint32_t tY = tmem + 256u + (uint32_t)((it & 1) * 128) + lane_base;
      ptx::mbar_wait(&bar.y_full[it & 1], ((uint32_t)it >> 1) & 1u);
      ptx::tc_fence_after_sync();
#pragma unroll 1
      for (int cb = 0; cb < 4; ++cb) {
        uint32_t r0[32];
        ptx::tmem_ld_32x32(tY + (uint32_t)(cb * 32), r0);
        ptx::tmem_ld_wait();
        if (cb == 3) {
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&bar.y_free[it & 1]);  // Y is in registers: tile it + 2 may overwrite it
        }
#pragma unroll
        for (int e = 0; e < 32; ++e) r0[e] = __float_as_uint(__uint_as_float(r0[e]) + sb2[cb * 32 + e]);
        if (row < p.M) store_row32(p.y + row * kD + cb * 32, r0);
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

struct BwdBars {
  uint64_t w_full[kBwdRing], w_empty[kBwdRing];
  uint64_t g_full, d_ready;
  uint64_t x_done, xacc_free;
};

// Data gradient.  MMA order per tile as in the header comment; G is single-buffered, so the transform of chunk c
// sits between M3(c) and [M4(c), M3(c + 1)] (see the transform branch below).
template <bool TRACE>
__global__ void __launch_bounds__(kThreads, 1)
ffn_dgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                 const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2t,
                 const __grid_constant__ CUtensorMap tmW1t, const __grid_constant__ CUtensorMap tmA,
                 const __grid_constant__ CUtensorMap tmDH, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ BwdBars bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int ops[3 * kMaxChunks];
  __shared__ int n_ops_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  // Shared memory: NO operand tile is resident.  The x and dY k-blocks are streamed through the ring in front of the
  // W1 / W2t k-blocks they are multiplied with (each is re-read from L2 once per chunk: +384 KB per tile), which
  // leaves room for a 9-slot ring and one 4 KB staging buffer per transform warp, from which the A / dH blocks leave
  // as asynchronous TMA stores.  Stores issued from registers (st.global) blocked the transform warps and slowed the
  // concurrent MMAs: 0.78 of 2.1 ms (profiles/r2_ffn_dgrad_experiments.txt).
  uint8_t* ring = smem;                     // [kBwdRing] 16 KB slots
  uint8_t* staging = ring + kBwdRing * kTile;
  float* sb1 = reinterpret_cast<float*>(staging + kXfWarps * kStage);
  for (int i = threadIdx.x; i < p.hidden; i += blockDim.x) sb1[i] = p.b1[i];
  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmDH);
    ptx::prefetch_tensormap(&tmX);
    ptx::prefetch_tensormap(&tmDY);
    ptx::prefetch_tensormap(&tmW1);
    ptx::prefetch_tensormap(&tmW2t);
    ptx::prefetch_tensormap(&tmW1t);
    for (int i = 0; i < kBwdRing; ++i) {
      ptx::mbar_init(&bar.w_full[i], 1);
      ptx::mbar_init(&bar.w_empty[i], 1);
    }
    ptx::mbar_init(&bar.g_full, 1);
    ptx::mbar_init(&bar.x_done, 1);
    ptx::mbar_init(&bar.d_ready, kXfWarps);
    ptx::mbar_init(&bar.xacc_free, kXfWarps);
    ptx::fence_mbar_init();
    n_ops_s = build_bwd_ops(ops, p.nc);
  }
  if (warp == 1) {
    ptx::tmem_alloc(&tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t tG = tmem + 256u, tX = tmem + 384u;
  const int n_ops = n_ops_s;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t st = 0, ph = 0;
      int it = 0;
      auto slot = [&](const CUtensorMap* tm, int c0, int c1) {  // one 16 KB k-block tile into the next ring slot
        ptx::mbar_wait(&bar.w_empty[st], ph ^ 1u);
        ptx::mbar_arrive_expect_tx(&bar.w_full[st], kTile);
        ptx::tma_load_3d(tm, &bar.w_full[st], ring + st * kTile, c0, c1, 0);
        if (++st == kBwdRing) { st = 0; ph ^= 1u; }
      };
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        for (int op = 0; op < n_ops; ++op) {
          const int kind = ops[op] >> 8, c = ops[op] & 255;
          for (int kb = 0; kb < 4; ++kb) {
            if (kind == OP_M1) {
              slot(&tmX, kb * 32, tile * 128);          // x[tile rows, in 32kb..]
              slot(&tmW1, kb * 32, c * kChunk);         // W1[128c.., in 32kb..]
            } else if (kind == OP_M3) {
              slot(&tmDY, kb * 32, tile * 128);         // dY[tile rows, out 32kb..]
              slot(&tmW2t, kb * 32, c * kChunk);        // W2t[128c.., out 32kb..]
            } else {
              slot(&tmW1t, c * kChunk + kb * 32, 0);    // W1t[:, 128c + 32kb..]
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      // warp-uniform schedule, single-thread issue from elect_one branches (see the note above build_fwd_ops)
      const uint32_t idesc = ptx::make_idesc_tf32(128, 128, 0, 0);
      const uint64_t dw0 = ptx::make_smem_desc(ptx::smem_u32(ring), 16, 1024, 2);
      uint32_t st = 0, ph = 0, cu = 0;  // cu: chunks whose M4 has been issued
      long long tw = 0, ta = 0;
      int tn = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        for (int op = 0; op < n_ops; ++op) {
          const long long t0 = TRACE ? clock64() : 0;
          const int kind = ops[op] >> 8, c = ops[op] & 255;
          const uint32_t tH = tmem + (uint32_t)((c & 1) * 128);
          if (kind == OP_M4) {
            twait<TRACE>(&bar.d_ready, cu & 1u, ta);  // the transform warps have read H_c, G_c and written dH_c over G_c
            ++cu;
            ptx::tc_fence_after_sync();
            if (c == 0) {
              twait<TRACE>(&bar.xacc_free, ((uint32_t)it & 1u) ^ 1u, ta);  // the previous tile's dX has been read out
              ptx::tc_fence_after_sync();
            }
          }
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {
            uint32_t sa = 0;  // M1 / M3: the slot holding the x / dY k-block, followed by the slot of the weight k-block
            if (kind != OP_M4) {
              twait<TRACE>(&bar.w_full[st], ph, tw);
              sa = st;
              if (++st == kBwdRing) { st = 0; ph ^= 1u; }
            }
            twait<TRACE>(&bar.w_full[st], ph, tw);
            ptx::tc_fence_after_sync();
            const uint64_t db = dw0 + (uint64_t)(st * kTile16);
            if (ptx::elect_one()) {
              if (kind == OP_M4) {
#pragma unroll
                for (int k8 = 0; k8 < 4; ++k8)
                  mma_tf32_ts(tX, tG + (uint32_t)(kb * 32 + k8 * 8), db + (uint64_t)(k8 * 2), idesc, (kb | k8) ? 1u : (c ? 1u : 0u));
              } else {
                const uint64_t da = dw0 + (uint64_t)(sa * kTile16);
                const uint32_t acc = kind == OP_M1 ? tH : tG;
#pragma unroll
                for (int k8 = 0; k8 < 4; ++k8)
                  ptx::mma_tf32_ss(acc, da + (uint64_t)(k8 * 2), db + (uint64_t)(k8 * 2), idesc, (kb | k8) ? 1u : 0u);
                ptx::mma_commit(&bar.w_empty[sa]);
              }
              ptx::mma_commit(&bar.w_empty[st]);
              if (kb == 3) {
                if (kind == OP_M3) ptx::mma_commit(&bar.g_full);  // H_c (issued earlier) and G_c are complete
                if (kind == OP_M4 && c == p.nc - 1) ptx::mma_commit(&bar.x_done);
              }
            }
            __syncwarp();
            if (++st == kBwdRing) { st = 0; ph ^= 1u; }
          }
          const int n = it * p.nc + c;
          if (TRACE && blockIdx.x == 0 && lane == 0 && tn < 8192 - 6) {
            long long* t = p.trace + 8192 + tn;
            t[0] = kind; t[1] = n; t[2] = t0; t[3] = clock64(); t[4] = tw; t[5] = ta;
            tn += 6; tw = 0; ta = 0;
          }
        }
      }
    }
  } else {
    // All 16 transform warps work on EVERY chunk (4 lane quadrants x 4 column groups of 32): G is single-buffered, so
    // M3(c) -> transform(c) -> [M4(c), M3(c + 1)] is a serial chain and the transform has to be as short as possible.
    // A warp hands its part of dH_c back first and streams its A / dH block out to global memory afterwards, while
    // the MMA warp already runs M4(c), M3(c + 1) and M1(c + 2).
    const int q = warp & 3;
    const int part = (warp - 2) >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const float cs = kTruncComp * p.dscale;
    uint8_t* sb = staging + (warp - 2) * kStage;
    uint32_t cu = 0;
    int tn = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
      const long long row = (long long)tile * 128 + q * 32 + lane;
      const uint32_t rs = p.thr ? row_seed((unsigned long long)row, p.seed) : 0u;
      for (int c = 0; c < p.nc; ++c, ++cu) {
        const long long t0 = TRACE ? clock64() : 0;
        ptx::mbar_wait(&bar.g_full, cu & 1u);
        const long long t1 = TRACE ? clock64() : 0;
        ptx::tc_fence_after_sync();
        const int col0 = part * 32;
        uint32_t ra[32], rd[32];
        ptx::tmem_ld_32x32(tmem + (uint32_t)((c & 1) * 128 + col0) + lane_base, ra);
        ptx::tmem_ld_32x32(tG + (uint32_t)col0 + lane_base, rd);
        ptx::tmem_ld_wait();
        const float* bias = sb1 + c * kChunk + col0;
        if (TRACE && (p.dbg & 1)) {
        } else if (p.thr)
          bwd_transform<true>(ra, rd, bias, p.act, cs, p.thr, group_seed(rs, (c * kChunk + col0) >> 5));
        else
          bwd_transform<false>(ra, rd, bias, p.act, cs, 0u, 0u);
        ptx::tmem_st_32x32(tG + (uint32_t)col0 + lane_base, rd);
        ptx::tmem_st_wait();
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bar.d_ready);
        const long long t2 = TRACE ? clock64() : 0;
        const int col = c * kChunk + col0;
        if (!(TRACE && (p.dbg & 4))) {
          stage_block(&tmA, sb, lane, ra, col, tile * 128 + q * 32);   // rows >= M are clipped by the TMA unit
          stage_block(&tmDH, sb, lane, rd, col, tile * 128 + q * 32);
        }
        const float csum = warp_column_sums(rd, lane) * (1.0f / kTruncComp);  // rows >= M carry dY = 0, hence dH = 0
        if (p.db1_part != nullptr) p.db1_part[((long long)tile * 4 + q) * p.hidden + col + lane] = csum;
        if (TRACE && blockIdx.x == 0 && q == 0 && (part & 1) == 0 && lane == 0 && tn < 4096 - 6) {
          long long* t = p.trace + 16384 + (part >> 1) * 4096 + tn;
          t[0] = 4; t[1] = cu; t[2] = t0; t[3] = t1; t[4] = t2; t[5] = clock64();
          tn += 6;
        }
      }
      // ---- tile done: dX (32 columns per warp)
      ptx::mbar_wait(&bar.x_done, (uint32_t)it & 1u);
      ptx::tc_fence_after_sync();
      {
        uint32_t r[32];
        ptx::tmem_ld_32x32(tX + (uint32_t)(part * 32) + lane_base, r);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bar.xacc_free);
        if (row < p.M) store_row32(p.dx + row * kD + part * 32, r);
      }
    }
    if (lane == 0) ptx::bulk_wait_all();  // staged blocks fully written before the CTA (and its shared memory) retires
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem, 512);
  }
}

__global__ void ffn_mask_kernel(uint8_t* mask, long long M, int hidden, uint32_t thr, unsigned long long seed) {
  const long long n = M * (hidden / 32);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / (hidden / 32);
    const int g = (int)(i % (hidden / 32));
    uint32_t h = group_seed(row_seed((unsigned long long)row, seed), g);
    for (int j = 0; j < 32; ++j) {
      h = h * kLcgA + kLcgC;
      mask[row * hidden + g * 32 + j] = (thr == 0u || h >= thr) ? 1 : 0;
    }
  }
}

static TensorView3 view2(const void* ptr, long long cols, long long rows) {
  return TensorView3{ptr, {(unsigned long long)cols, (unsigned long long)rows, 1ull},
                     {(unsigned long long)cols * 4ull, (unsigned long long)cols * (unsigned long long)rows * 4ull}};
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

static int fill(Params& p, int64_t M, int64_t D, int64_t hidden, int act, float drop_p, uint64_t seed) {
  if (M <= 0 || !(drop_p >= 0.f && drop_p < 1.f)) return XM_ERR_INVALID;
  if (D != kD || hidden <= 0 || hidden % kChunk != 0 || hidden / kChunk > kMaxChunks || (act != XM_ACT_GELU && act != XM_ACT_RELU) ||
      M > ((int64_t)1 << 31) - 256)
    return XM_ERR_UNSUPPORTED;
  p.M = M;
  p.tiles = (int)((M + 127) / 128);
  p.hidden = (int)hidden;
  p.nc = (int)(hidden / kChunk);
  p.act = act;
  p.dscale = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
  p.thr = drop_p > 0.f ? (uint32_t)((double)drop_p * 4294967296.0) : 0u;
  p.seed = seed;
  return XM_OK;
}

}  // namespace ffn
}  // namespace xm

XM_DEFINE_SEED_EPOCH_SLOT(ffn_fused)

using namespace xm;

static long long* g_ffn_trace = nullptr;
static int g_ffn_dbg = 0;

extern "C" {

int xm_ffn_fused_supported(int64_t D, int64_t hidden, int act) {
  return D == ffn::kD && hidden > 0 && hidden % ffn::kChunk == 0 && hidden / ffn::kChunk <= ffn::kMaxChunks &&
         (act == XM_ACT_GELU || act == XM_ACT_RELU);
}

int xm_ffn_fused_nblk(int64_t M) { return (int)((M + 127) / 128) * 4; }

int xm_ffn_fused_fwd_f32(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, float* y, int64_t M,
                         int64_t D, int64_t hidden, int act, float drop_p, uint64_t seed, void* stream) {
  if (!x || !w1 || !b1 || !w2 || !b2 || !y) return XM_ERR_INVALID;
  if (reinterpret_cast<uintptr_t>(y) & 31) return XM_ERR_INVALID;
  ffn::Params p{};
  int rc = ffn::fill(p, M, D, hidden, act, drop_p, seed);
  if (rc != XM_OK) return rc;
  p.b1 = b1;
  p.b2 = b2;
  p.y = y;
  CUtensorMap mx, m1, m2;
  rc = encode_tmap(&mx, ffn::view2(x, D, M), 32, 128, 0);
  if (rc == XM_OK) rc = encode_tmap(&m1, ffn::view2(w1, D, hidden), 32, 128, 0);
  if (rc == XM_OK) rc = encode_tmap(&m2, ffn::view2(w2, hidden, D), 32, 128, 0);
  const int ctas = p.tiles < kNumSMs ? p.tiles : kNumSMs;
  if (g_ffn_trace != nullptr) {  // debug instance: CTA 0 logs per-op clocks (xm_debug_set_ffn_trace)
    p.trace = g_ffn_trace;
    p.dbg = g_ffn_dbg;
    if (rc == XM_OK) rc = ffn::set_smem(ffn::ffn_fwd_kernel<true>, ffn::kFwdSmem);
    if (rc != XM_OK) return rc;
    ffn::ffn_fwd_kernel<true><<<ctas, ffn::kFwdThreads, ffn::kFwdSmem, (cudaStream_t)stream>>>(mx, m1, m2, p);
    return check_launch();
  }
  if (rc == XM_OK) rc = ffn::set_smem(ffn::ffn_fwd_kernel<false>, ffn::kFwdSmem);
  if (rc != XM_OK) return rc;
  ffn::ffn_fwd_kernel<false><<<ctas, ffn::kFwdThreads, ffn::kFwdSmem, (cudaStream_t)stream>>>(mx, m1, m2, p);
  return check_launch();
}

int xm_debug_set_ffn_trace(int64_t* device_buffer) {
  g_ffn_trace = reinterpret_cast<long long*>(device_buffer);
  return XM_OK;
}
int xm_debug_set_ffn_flags(int flags) {
  g_ffn_dbg = flags;
  return XM_OK;
}

int xm_ffn_fused_dgrad_f32(const float* x, const float* dy, const float* w1, const float* b1, const float* w2t, const float* w1t,
                           float* a, float* dh, float* dx, float* db1_part, int64_t M, int64_t D, int64_t hidden, int act,
                           float drop_p, uint64_t seed, void* stream) {
  if (!x || !dy || !w1 || !b1 || !w2t || !w1t || !a || !dh || !dx) return XM_ERR_INVALID;
  ffn::Params p{};
  int rc = ffn::fill(p, M, D, hidden, act, drop_p, seed);
  if (rc != XM_OK) return rc;
  p.b1 = b1;
  p.dx = dx;
  p.db1_part = db1_part;
  if (reinterpret_cast<uintptr_t>(dx) & 31) return XM_ERR_INVALID;
  CUtensorMap mx, my, m1, m2t, m1t, ma, mdh;
  rc = encode_tmap(&mx, ffn::view2(x, D, M), 32, 128, 0);
  if (rc == XM_OK) rc = encode_tmap(&my, ffn::view2(dy, D, M), 32, 128, 0);
  if (rc == XM_OK) rc = encode_tmap(&m1, ffn::view2(w1, D, hidden), 32, 128, 0);
  if (rc == XM_OK) rc = encode_tmap(&m2t, ffn::view2(w2t, D, hidden), 32, 128, 0);
  if (rc == XM_OK) rc = encode_tmap(&m1t, ffn::view2(w1t, hidden, D), 32, 128, 0);
  if (rc == XM_OK) rc = encode_tmap(&ma, ffn::view2(a, hidden, M), 32, 32, 0);
  if (rc == XM_OK) rc = encode_tmap(&mdh, ffn::view2(dh, hidden, M), 32, 32, 0);
  const int ctas = p.tiles < kNumSMs ? p.tiles : kNumSMs;
  if (g_ffn_trace != nullptr) {
    p.trace = g_ffn_trace;
    p.dbg = g_ffn_dbg;
    if (rc == XM_OK) rc = ffn::set_smem(ffn::ffn_dgrad_kernel<true>, ffn::kBwdSmem);
    if (rc != XM_OK) return rc;
    ffn::ffn_dgrad_kernel<true><<<ctas, ffn::kThreads, ffn::kBwdSmem, (cudaStream_t)stream>>>(mx, my, m1, m2t, m1t, ma, mdh, p);
    return check_launch();
  }
  if (rc == XM_OK) rc = ffn::set_smem(ffn::ffn_dgrad_kernel<false>, ffn::kBwdSmem);
  if (rc != XM_OK) return rc;
  ffn::ffn_dgrad_kernel<false><<<ctas, ffn::kThreads, ffn::kBwdSmem, (cudaStream_t)stream>>>(mx, my, m1, m2t, m1t, ma, mdh, p);
  return check_launch();
}

int xm_ffn_fused_mask_u8(uint8_t* mask, int64_t M, int64_t hidden, float drop_p, uint64_t seed, void* stream) {
  if (!mask || M <= 0 || hidden <= 0 || hidden % 32 != 0 || !(drop_p >= 0.f && drop_p < 1.f)) return XM_ERR_INVALID;
  const uint32_t thr = drop_p > 0.f ? (uint32_t)((double)drop_p * 4294967296.0) : 0u;
  const long long n = M * (hidden / 32);
  const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  ffn::ffn_mask_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(mask, M, (int)hidden, thr, seed);
  return check_launch();
}

}  // extern "C"
