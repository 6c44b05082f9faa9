// Band power as a tensor-core DFT (sm_100a).  Only the rFFT bins inside the requested bands are needed (26 for
// theta / alpha / beta at 1 kHz, nfft 1024), so per window
//      X[c, j] = sum_n x[c, n] * T[j, n],   T[j, n] = taper[n] * {cos, -sin}(2 pi k_j n / nfft)
// is a (channels x win) x (win x 64) product: M = 128 channels of one window on the TMEM lanes, N = 32 cosine + 32
// sine columns, K = the window's samples.  fp32 accuracy on the tf32 tensor cores by the 3-pass split
//      x = hi + lo (hi = x rounded to tf32, lo = x - hi),  T = Thi + Tlo,      X = hi Thi + hi Tlo + lo Thi.
// Four transform warps read the TMA-filled sample tile from shared memory, split it in registers and write hi and lo
// to tensor memory (tcgen05.st); all three products take their A operand from there, so nothing depends on how the
// tensor core would have converted raw fp32 samples.  The K loop is cut into <= 6 chunks with their own
// accumulators, summed in fp32 registers: the tensor core's accumulation truncates (~3e-8 per K = 8 step), which
// over 1024 samples x 3 passes would otherwise bias |X|^2 by more than the 1e-5 this path promises.
// The A-operand read (64 B/clk) paces the kernel: 12 MMAs x 64 clk per 32-sample block, ~87 ms for 1 M windows of
// 128 x 1024 samples, next to an HBM floor of 80 ms (525 GB) and 327 ms for the Stockham FFT kernel.
//
// The window samples are read in place from rec (R, C, n) through a 3-D tensor map (coordinate = window start),
// never gathered.  Eligible shapes (else the Stockham FFT kernel in spectral.cu runs): <= 32 bins in total, <= 8
// bands, win <= 1024, n % 4 == 0, hop % 4 == 0.  HBM-bound target: every sample is read hop/win ... once per window.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "../../include/xmodal_b200.h"
#include "gemm_engine.cuh"

namespace xm {
namespace bp {

constexpr int kBins = 32;       // cosine columns [0, 32), sine columns [32, 64)
constexpr int kBands = 8;
constexpr int kStages = 5;
constexpr int kStageBytes = 16384 + 2 * 8192;  // x tile (128 ch x 32 samples) | Thi (64 x 32) | Tlo (64 x 32)
constexpr int kSmem = kStages * kStageBytes + 1024;
constexpr int kThreads = 64 + 4 * 32;
constexpr int kMaxChunks = 6;

struct Tables {        // written by the twiddle kernel behind the two (64, Kpad) matrices
  float scale[kBins];  // s_k / (nfft * sum taper^2) per column
  int band_lo[kBands], band_hi[kBands];  // [lo, hi) positions of band b in the concatenated bin list
};

struct Params {
  int C, n_win, hop, KB, chunk_len, n_chunks, n_bands, ct;  // ct = 128-channel tiles per window
  long long items;                                           // windows * ct
  const Tables* tab;
  float* power;  // (windows, C, n_bands)
};

// T (2 x 64 x Kpad): rows j < 32 cosine, 32 + j sine of the j-th requested bin; hi / lo tf32 split (round to nearest).
__global__ void twiddle_kernel(const float* __restrict__ taper, int win, int nfft, int Kpad, const int* __restrict__ band_bins,
                               int n_bands, float inv_norm, float* __restrict__ thi, float* __restrict__ tlo,
                               Tables* __restrict__ tab) {
  const int total = 64 * Kpad;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int row = i / Kpad, n = i - row * Kpad;
    const int j = row & 31;
    int k = -1, pos = 0;  // the j-th bin of the concatenated [lo, hi) ranges
    for (int b = 0; b < n_bands; ++b) {
      const int lo = band_bins[2 * b], cnt = max(band_bins[2 * b + 1] - lo, 0);
      if (k < 0 && j < pos + cnt) k = lo + (j - pos);
      pos += cnt;
    }
    float v = 0.f;
    if (k >= 0 && n < win) {
      const long long kn = ((long long)k * n) % nfft;
      double s, c;
      sincospi(2.0 * (double)kn / (double)nfft, &s, &c);
      v = (float)((double)taper[n] * (row < 32 ? c : -s));
    }
    const float h = round_tf32(v);
    thi[i] = h;
    tlo[i] = round_tf32(v - h);
  }
  if (blockIdx.x == 0 && threadIdx.x < kBins) {
    const int j = threadIdx.x;
    int k = -1, pos = 0;
    for (int b = 0; b < n_bands; ++b) {
      const int lo = band_bins[2 * b], cnt = max(band_bins[2 * b + 1] - lo, 0);
      if (k < 0 && j < pos + cnt) k = lo + (j - pos);
      if (j == 0) {
        tab->band_lo[b] = pos;
        tab->band_hi[b] = pos + cnt;
      }
      pos += cnt;
    }
    tab->scale[j] = k < 0 ? 0.f : ((k == 0 || 2 * k == nfft) ? 1.f : 2.f) * inv_norm;
    if (j == 0)
      for (int b = n_bands; b < kBands; ++b) tab->band_lo[b] = tab->band_hi[b] = 0;
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

struct Bars {
  uint64_t full[kStages], empty[kStages];
  uint64_t lo_ready[2], lo_free[2];
  uint64_t acc_full, acc_free;
};

// TMEM: accumulator of chunk c at columns [64 c, 64 c + 64), c < 6; operand buffers b = 0, 1 at 384 + 64 b: hi (32), lo (32)
__global__ void __launch_bounds__(kThreads, 1)
bandpower_dft_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmThi,
                     const __grid_constant__ CUtensorMap tmTlo, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ Bars bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmX);
    ptx::prefetch_tensormap(&tmThi);
    ptx::prefetch_tensormap(&tmTlo);
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&bar.full[i], 1);
      ptx::mbar_init(&bar.empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bar.lo_ready[i], 4);
      ptx::mbar_init(&bar.lo_free[i], 1);
    }
    ptx::mbar_init(&bar.acc_full, 1);
    ptx::mbar_init(&bar.acc_free, 4);
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
  const uint32_t tOp = tmem + 64 * kMaxChunks;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (long long w = blockIdx.x; w < p.items; w += gridDim.x) {
        const long long g = w / p.ct;            // window
        const int ct = (int)(w - g * p.ct);      // channel tile
        const int r = (int)(g / p.n_win);
        const int start = (int)(g - (long long)r * p.n_win) * p.hop;
        for (int kb = 0; kb < p.KB; ++kb) {
          ptx::mbar_wait(&bar.empty[s], ph ^ 1u);
          uint8_t* st = smem + s * kStageBytes;
          ptx::mbar_arrive_expect_tx(&bar.full[s], kStageBytes);
          ptx::tma_load_3d(&tmX, &bar.full[s], st, start + kb * 32, ct * 128, r);
          ptx::tma_load_3d(&tmThi, &bar.full[s], st + 16384, kb * 32, 0, 0);
          ptx::tma_load_3d(&tmTlo, &bar.full[s], st + 24576, kb * 32, 0, 0);
          if (++s == kStages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    {
      // warp-uniform schedule; each k-block's twelve MMAs and two commits issue from one elect_one branch, back to back
      // in SASS (inside `if (lane == 0)` every MMA paid an R2UR move and an ELECT loop, profiles/r2_mma_rate_probe.txt)
      const uint32_t idesc = ptx::make_idesc_tf32(128, 64, 0, 0);
      const uint64_t dh0 = ptx::make_smem_desc(ptx::smem_u32(smem) + 16384, 16, 1024, 2);
      const uint64_t dl0 = ptx::make_smem_desc(ptx::smem_u32(smem) + 24576, 16, 1024, 2);
      int s = 0;
      uint32_t ph = 0;
      uint32_t kbg = 0;  // k-blocks issued so far (lo buffer = kbg & 1)
      int it = 0;
      for (long long w = blockIdx.x; w < p.items; w += gridDim.x, ++it) {
        ptx::mbar_wait(&bar.acc_free, ((uint32_t)it & 1u) ^ 1u);  // the epilogue has read the previous window's sums
        ptx::tc_fence_after_sync();
        for (int kb = 0; kb < p.KB; ++kb, ++kbg) {
          const int chunk = kb / p.chunk_len;
          const bool first = (kb - chunk * p.chunk_len) == 0;
          const uint32_t acc = tmem + (uint32_t)(chunk * 64);
          ptx::mbar_wait(&bar.full[s], ph);
          ptx::tc_fence_after_sync();
          const uint64_t dh = dh0 + (uint64_t)((uint32_t)s * (uint32_t)(kStageBytes >> 4));
          const uint64_t dl = dl0 + (uint64_t)((uint32_t)s * (uint32_t)(kStageBytes >> 4));
          const uint32_t lb = kbg & 1u;
          ptx::mbar_wait(&bar.lo_ready[lb], (kbg >> 1) & 1u);  // hi and lo of this block are in tensor memory
          ptx::tc_fence_after_sync();
          const uint32_t th = tOp + lb * 64u, tl = th + 32u;
          const uint32_t acc0 = first ? 0u : 1u;
          if (ptx::elect_one()) {
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8)  // hi * Thi
              mma_tf32_ts(acc, th + (uint32_t)(k8 * 8), dh + (uint64_t)(k8 * 2), idesc, k8 == 0 ? acc0 : 1u);
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8)  // hi * Tlo
              mma_tf32_ts(acc, th + (uint32_t)(k8 * 8), dl + (uint64_t)(k8 * 2), idesc, 1u);
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8)  // lo * Thi
              mma_tf32_ts(acc, tl + (uint32_t)(k8 * 8), dh + (uint64_t)(k8 * 2), idesc, 1u);
            ptx::mma_commit(&bar.empty[s]);
            ptx::mma_commit(&bar.lo_free[lb]);
            if (kb == p.KB - 1) ptx::mma_commit(&bar.acc_full);
          }
          __syncwarp();
          if (++s == kStages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else {
    const int q = warp & 3;  // TMEM lane quadrant of this warp
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int row = q * 32 + lane;  // channel inside the tile
    int s = 0;
    uint32_t ph = 0, kbg = 0;
    int it = 0;
    for (long long w = blockIdx.x; w < p.items; w += gridDim.x, ++it) {
      const long long g = w / p.ct;
      const int ct = (int)(w - g * p.ct);
      for (int kb = 0; kb < p.KB; ++kb, ++kbg) {
        const uint32_t lb = kbg & 1u;
        ptx::mbar_wait(&bar.full[s], ph);
        ptx::mbar_wait(&bar.lo_free[lb], ((kbg >> 1) & 1u) ^ 1u);  // the MMA that read this lo tile has retired
        ptx::tc_fence_after_sync();
        const uint8_t* xt = smem + s * kStageBytes + row * 128;
        uint32_t rh[32], rl[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {  // 16-B chunk j of row `row` sits at chunk j ^ (row & 7) (SWIZZLE_128B)
          const float4 v = *reinterpret_cast<const float4*>(xt + ((j ^ (row & 7)) << 4));
          const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float h = round_tf32(xs[e]);
            rh[4 * j + e] = __float_as_uint(h);
            rl[4 * j + e] = __float_as_uint(xs[e] - h);
          }
        }
        ptx::tmem_st_32x32(tOp + lb * 64u + lane_base, rh);
        ptx::tmem_st_32x32(tOp + lb * 64u + 32u + lane_base, rl);
        ptx::tmem_st_wait();
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bar.lo_ready[lb]);
        if (++s == kStages) { s = 0; ph ^= 1u; }
      }
      // ---- window done: X = sum of the chunk accumulators, |X|^2 -> band sums
      ptx::mbar_wait(&bar.acc_full, (uint32_t)it & 1u);
      ptx::tc_fence_after_sync();
      float re[32], im[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) re[j] = im[j] = 0.f;
#pragma unroll
      for (int c = 0; c < kMaxChunks; ++c) {
        if (c < p.n_chunks) {
          uint32_t a[32], b[32];
          ptx::tmem_ld_32x32(tmem + (uint32_t)(c * 64) + lane_base, a);
          ptx::tmem_ld_32x32(tmem + (uint32_t)(c * 64 + 32) + lane_base, b);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            re[j] += __uint_as_float(a[j]);
            im[j] += __uint_as_float(b[j]);
          }
        }
      }
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bar.acc_free);
      const int ch = ct * 128 + row;
      if (ch < p.C) {
        float* out = p.power + ((long long)g * p.C + ch) * p.n_bands;
#pragma unroll
        for (int j = 0; j < 32; ++j) re[j] = fmaf(re[j], re[j], im[j] * im[j]) * __ldg(&p.tab->scale[j]);
        for (int b = 0; b < p.n_bands; ++b) {
          const int lo = __ldg(&p.tab->band_lo[b]), hi = __ldg(&p.tab->band_hi[b]);
          float sum = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) sum += (j >= lo && j < hi) ? re[j] : 0.f;
          out[b] = sum;
        }
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

// ------------------------------------------------------------------------------------------------------------------
// Overlap kernel (hop == win / 2, win % 64 == 0: Welch-style half-overlapping windows, BASELINE config 5; v1 above
// serves every other shape).  What the measurements of v1 and three intermediate variants say
// (profiles/r2_bandpower_experiments.txt):
//   * an MMA whose A operand comes from tensor memory costs 80 clk PER 64 COLUMNS of N (N-stacking the twiddles does
//     not make it cheaper); an MMA with both operands in shared memory costs (A + B bytes) / ~105 B/clk, and that port is
//     shared with the TMA writes and the transform warps' loads: v1 is MMA-bound (12 x 80 clk per 32-sample block),
//     an all-shared-memory variant is port-bound (120 KB per block);
//   * switching the accumulator between MMAs costs nothing measurable; the loads are not the limit (the producer waits on
//     `empty` 60 % of the time).
// So the way forward is fewer MMA-cycles and fewer shared-memory bytes PER WINDOW: a 32-sample block of a recording
// belongs to TWO windows (second half of window w, first half of window w + 1), and here it is loaded, split and
// multiplied ONCE for both:
//      D[128 ch, 256] (+)= x_blk (raw fp32 tile in shared memory = hi by truncation) * [Thi(kb+H) ; Thi(kb) ; Tlo(kb+H) ; Tlo(kb)]^T
//      L[128 ch, 128] (+)= lo_blk (tensor memory)                                     * [Thi(kb+H) ; Thi(kb)]^T
// one N = 256 MMA (tensor-core math rate, 128 clk) and one N = 128 MMA (160 clk) per k8: 1152 clk per block for two
// windows instead of 2 x 960, 128 KB of shared-memory traffic instead of 2 x 72 KB, and every sample crosses L2 -> SM once.
// Columns: [0, 64) older window x Thi, [64, 128) newer window x Thi, [128, 192) older x Tlo, [192, 256) newer x Tlo.
// The lo products go to an accumulator of their own (columns [320, 448)): they are 2^-11 of the sums, so their accumulation
// error is irrelevant, and keeping them out of the main accumulator halves ITS accumulation steps -- the tensor core
// truncates at every step, which is this path's error floor (measured max relative error against a float64 computation,
// tools/bp_err.py: v1 4.7e-6; here 4.6e-6 with one drain per hop of 16 blocks, 2.5e-6 with XM_BP_CHUNK=8 at -3 % speed).
// Four epilogue warps drain the accumulators at every chunk end (a hop, or `XM_BP_CHUNK` blocks) and keep the two live
// windows' partial sums in registers (setmaxnreg: 232 registers for them, 64 / 112 for the other roles -- with one budget
// for all, one of the four arrays lived in local memory); at a hop boundary the older window is finished (band sums ->
// HBM) and the newer one becomes the older one.  A CTA walks an item = (recording, channel tile, run of <= 16 windows)
// along the time axis; the first / last hop of an item computes one half-window that belongs to its neighbour and is
// dropped (1/16 overhead).
constexpr int kThreads2 = 3 * 128;  // warpgroup 0: TMA producer, MMA issuer (+ 2 idle warps); 1: transform; 2: epilogue
constexpr int kStages2 = 4;
constexpr int kStageBytes2 = 16384 + 4 * 8192;  // x | Thi(old) | Thi(new) | Tlo(old) | Tlo(new)
constexpr int kSmem2 = kStages2 * kStageBytes2 + 1024;
constexpr float kTruncComp = 1.0f + 0.7213f / 2048.0f;  // makes the tensor core's truncation of lo zero-mean

struct Params2 {
  int C, n_win, HB, chunk_len, chunks_per_seg, n_bands, ct, wpi, ipr;  // HB = blocks per hop; wpi = windows per item; ipr = items per recording
  long long items;                                                      // n_rec * ct * ipr
  const Tables* tab;
  float* power;
};

struct Bars2 {
  uint64_t full[kStages2], empty[kStages2];
  uint64_t lo_ready[2], lo_free[2];
  uint64_t acc_full, acc_free;
};

struct Item2 {
  int r, ct, w0, nw;  // recording, channel tile, first window, windows in this item
};
XM_DEVICE Item2 decode_item2(const Params2& p, long long it) {
  Item2 q;
  const int piece = (int)(it % p.ipr);
  const long long rc = it / p.ipr;
  q.ct = (int)(rc % p.ct);
  q.r = (int)(rc / p.ct);
  q.w0 = piece * p.wpi;
  q.nw = min(p.wpi, p.n_win - q.w0);
  return q;
}

// TMEM: main accumulator columns [0, 256); lo operand buffers b = 0, 1 at 256 + 32 b; lo-product accumulator [320, 448)
__global__ void __launch_bounds__(kThreads2, 1)
bandpower_dft2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmThi,
                      const __grid_constant__ CUtensorMap tmTlo, const Params2 p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ Bars2 bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmX);
    ptx::prefetch_tensormap(&tmThi);
    ptx::prefetch_tensormap(&tmTlo);
    for (int i = 0; i < kStages2; ++i) {
      ptx::mbar_init(&bar.full[i], 1);
      ptx::mbar_init(&bar.empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bar.lo_ready[i], 4);
      ptx::mbar_init(&bar.lo_free[i], 1);
    }
    ptx::mbar_init(&bar.acc_full, 1);
    ptx::mbar_init(&bar.acc_free, 4);
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
  const uint32_t tOp = tmem + 256u, tLoAcc = tmem + 320u;

  // Register budget per warpgroup (setmaxnreg): the epilogue warps hold the partial sums of two windows (128 registers)
  // plus two 32-column loads; 168 registers for everybody spilled one of the four arrays to local memory.
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 64;" ::: "memory");
  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (long long it = blockIdx.x; it < p.items; it += gridDim.x) {
        const Item2 q = decode_item2(p, it);
        const int nblk = (q.nw + 1) * p.HB;
        const int b0 = q.w0 * p.HB;
        for (int i = 0; i < nblk; ++i) {
          const int kb = i % p.HB;  // k-block of the newer window; the older one is at kb + HB
          ptx::mbar_wait(&bar.empty[s], ph ^ 1u);
          uint8_t* st = smem + s * kStageBytes2;
          ptx::mbar_arrive_expect_tx(&bar.full[s], kStageBytes2);
          ptx::tma_load_3d(&tmX, &bar.full[s], st, (b0 + i) * 32, q.ct * 128, q.r);
          ptx::tma_load_3d(&tmThi, &bar.full[s], st + 16384, (kb + p.HB) * 32, 0, 0);
          ptx::tma_load_3d(&tmThi, &bar.full[s], st + 24576, kb * 32, 0, 0);
          ptx::tma_load_3d(&tmTlo, &bar.full[s], st + 32768, (kb + p.HB) * 32, 0, 0);
          ptx::tma_load_3d(&tmTlo, &bar.full[s], st + 40960, kb * 32, 0, 0);
          if (++s == kStages2) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc256 = ptx::make_idesc_tf32(128, 256, 0, 0), idesc128 = ptx::make_idesc_tf32(128, 128, 0, 0);
    const uint64_t dx0 = ptx::make_smem_desc(ptx::smem_u32(smem), 16, 1024, 2);
    const uint64_t dt0 = ptx::make_smem_desc(ptx::smem_u32(smem) + 16384, 16, 1024, 2);
    int s = 0;
    uint32_t ph = 0, kbg = 0, cg = 0;  // cg = chunks started so far
    for (long long it = blockIdx.x; it < p.items; it += gridDim.x) {
      const Item2 q = decode_item2(p, it);
      const int nblk = (q.nw + 1) * p.HB;
      for (int i = 0; i < nblk; ++i, ++kbg) {
        const int kb = i % p.HB;
        const bool first = (kb % p.chunk_len) == 0;
        const bool last = ((kb + 1) % p.chunk_len) == 0 || kb == p.HB - 1;
        if (first) {  // the epilogue warps have drained the previous chunk
          ptx::mbar_wait(&bar.acc_free, (cg & 1u) ^ 1u);
          ptx::tc_fence_after_sync();
          ++cg;
        }
        ptx::mbar_wait(&bar.full[s], ph);
        const uint32_t lb = kbg & 1u;
        ptx::mbar_wait(&bar.lo_ready[lb], (kbg >> 1) & 1u);  // lo of this block is in tensor memory
        ptx::tc_fence_after_sync();
        const uint64_t so = (uint64_t)((uint32_t)s * (uint32_t)(kStageBytes2 >> 4));
        const uint64_t dt = dt0 + so, dx = dx0 + so;
        const uint32_t tl = tOp + lb * 32u;
        const uint32_t acc0 = first ? 0u : 1u;
        if (ptx::elect_one()) {
#pragma unroll
          for (int k8 = 0; k8 < 4; ++k8)  // x * [Thi(old) ; Thi(new) ; Tlo(old) ; Tlo(new)], A = the raw sample tile
            ptx::mma_tf32_ss(tmem, dx + (uint64_t)(k8 * 2), dt + (uint64_t)(k8 * 2), idesc256, k8 == 0 ? acc0 : 1u);
#pragma unroll
          for (int k8 = 0; k8 < 4; ++k8)  // lo * [Thi(old) ; Thi(new)] into its own accumulator, one run per hop
            mma_tf32_ts(tLoAcc, tl + (uint32_t)(k8 * 8), dt + (uint64_t)(k8 * 2), idesc128, (k8 == 0 && kb == 0) ? 0u : 1u);
          ptx::mma_commit(&bar.empty[s]);
          ptx::mma_commit(&bar.lo_free[lb]);
          if (last) ptx::mma_commit(&bar.acc_full);
        }
        __syncwarp();
        if (++s == kStages2) { s = 0; ph ^= 1u; }
      }
    }
  }  // (warps 2, 3: idle)
  } else if (warp < 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 112;" ::: "memory");
    const int q4 = warp & 3;  // TMEM lane quadrant of this warp
    const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
    const int row = q4 * 32 + lane;  // channel inside the tile
    int s = 0;
    uint32_t ph = 0, kbg = 0;
    for (long long it = blockIdx.x; it < p.items; it += gridDim.x) {
      const Item2 q = decode_item2(p, it);
      const int nblk = (q.nw + 1) * p.HB;
      for (int i = 0; i < nblk; ++i, ++kbg) {
        const uint32_t lb = kbg & 1u;
        ptx::mbar_wait(&bar.full[s], ph);
        const uint8_t* xt = smem + s * kStageBytes2 + row * 128;
        uint32_t rl[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {  // 16-B chunk j of row `row` sits at chunk j ^ (row & 7) (SWIZZLE_128B)
          const uint4 v = *reinterpret_cast<const uint4*>(xt + ((j ^ (row & 7)) << 4));
          const uint32_t xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 4; ++e)
            rl[4 * j + e] = __float_as_uint((__uint_as_float(xs[e]) - __uint_as_float(xs[e] & 0xFFFFE000u)) * kTruncComp);
        }
        ptx::mbar_wait(&bar.lo_free[lb], ((kbg >> 1) & 1u) ^ 1u);  // the MMAs that read this buffer have retired
        ptx::tc_fence_after_sync();
        ptx::tmem_st_32x32(tOp + lb * 32u + lane_base, rl);
        ptx::tmem_st_wait();
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bar.lo_ready[lb]);
        if (++s == kStages2) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;" ::: "memory");
    const int q4 = warp & 3;
    const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
    const int row = q4 * 32 + lane;
    uint32_t cg = 0;  // chunks drained so far
    for (long long it = blockIdx.x; it < p.items; it += gridDim.x) {
      const Item2 q = decode_item2(p, it);
      const int ch = q.ct * 128 + row;
      float ore[32], oim[32], nre[32], nim[32];  // partial sums of the older / newer live window
#pragma unroll
      for (int j = 0; j < 32; ++j) ore[j] = oim[j] = nre[j] = nim[j] = 0.f;
      for (int seg = 0; seg <= q.nw; ++seg) {
        for (int c = 0; c < p.chunks_per_seg; ++c, ++cg) {
          ptx::mbar_wait(&bar.acc_full, cg & 1u);
          ptx::tc_fence_after_sync();
          uint32_t a[32], b[32];
#define XM_BP_DRAIN(col_a, col_b, dst)                                   \
  ptx::tmem_ld_32x32(tmem + (uint32_t)(col_a) + lane_base, a);           \
  ptx::tmem_ld_32x32(tmem + (uint32_t)(col_b) + lane_base, b);           \
  ptx::tmem_ld_wait();                                                   \
  _Pragma("unroll") for (int j = 0; j < 32; ++j) dst[j] += __uint_as_float(a[j]) + __uint_as_float(b[j]);
          XM_BP_DRAIN(0, 128, ore)
          XM_BP_DRAIN(32, 160, oim)
          XM_BP_DRAIN(64, 192, nre)
          XM_BP_DRAIN(96, 224, nim)
          if (c == p.chunks_per_seg - 1) {  // the hop's lo products
            ptx::tmem_ld_32x32(tmem + 320u + lane_base, a);
            ptx::tmem_ld_32x32(tmem + 352u + lane_base, b);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              ore[j] += __uint_as_float(a[j]);
              oim[j] += __uint_as_float(b[j]);
            }
            ptx::tmem_ld_32x32(tmem + 384u + lane_base, a);
            ptx::tmem_ld_32x32(tmem + 416u + lane_base, b);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              nre[j] += __uint_as_float(a[j]);
              nim[j] += __uint_as_float(b[j]);
            }
          }
#undef XM_BP_DRAIN
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&bar.acc_free);
        }
        // hop boundary: the older window (w0 + seg - 1) is complete; the newer one becomes the older one
        if (seg >= 1 && ch < p.C) {
          const long long g = (long long)q.r * p.n_win + q.w0 + seg - 1;
          float* out = p.power + (g * p.C + ch) * p.n_bands;
#pragma unroll
          for (int j = 0; j < 32; ++j) ore[j] = fmaf(ore[j], ore[j], oim[j] * oim[j]) * __ldg(&p.tab->scale[j]);  // (overwritten below)
          for (int b2 = 0; b2 < p.n_bands; ++b2) {
            const int lo = __ldg(&p.tab->band_lo[b2]), hi = __ldg(&p.tab->band_hi[b2]);
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) sum += (j >= lo && j < hi) ? ore[j] : 0.f;
            out[b2] = sum;
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          ore[j] = nre[j];
          oim[j] = nim[j];
          nre[j] = nim[j] = 0.f;
        }
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

}  // namespace bp
}  // namespace xm

using namespace xm;

extern "C" {

int64_t xm_bandpower_dft_workspace_floats(int64_t win) {
  const int64_t Kpad = (win + 31) / 32 * 32;
  return 2 * 64 * Kpad + (int64_t)(sizeof(bp::Tables) + 3) / 4 + 64;
}

int xm_bandpower_dft_supported(int64_t C, int64_t n_samples, int64_t win, int64_t hop, int64_t nfft, int n_bands,
                               int total_bins) {
  return C >= 1 && win >= 1 && win <= 1024 && nfft >= win && (n_samples % 4) == 0 && (hop % 4) == 0 && n_bands >= 1 &&
         n_bands <= bp::kBands && total_bins >= 1 && total_bins <= bp::kBins;
}

int xm_bandpower_dft_f32(const float* rec, int64_t n_rec, int64_t C, int64_t n_samples, int64_t win, int64_t hop,
                         int64_t nfft, float fs, const float* taper, float taper_sumsq, const int32_t* band_bins,
                         int n_bands, int total_bins, float* workspace, float* power, void* stream) {
  (void)fs;
  if (!rec || !taper || !band_bins || !power || !workspace || n_rec <= 0 || hop <= 0 || n_samples < win ||
      !(taper_sumsq > 0.f))
    return XM_ERR_INVALID;
  if (!xm_bandpower_dft_supported(C, n_samples, win, hop, nfft, n_bands, total_bins)) return XM_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(rec) | reinterpret_cast<uintptr_t>(workspace)) & 15) return XM_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const int KB = (int)((win + 31) / 32), Kpad = KB * 32;
  float* thi = workspace;
  float* tlo = workspace + 64ll * Kpad;
  bp::Tables* tab = reinterpret_cast<bp::Tables*>(workspace + 128ll * Kpad);
  bp::twiddle_kernel<<<64, 256, 0, st>>>(taper, (int)win, (int)nfft, Kpad, band_bins, n_bands,
                                          1.0f / ((float)nfft * taper_sumsq), thi, tlo, tab);
  int rc = check_launch();
  if (rc != XM_OK) return rc;
  const long long n_win = (n_samples - win) / hop + 1;
  bp::Params p{};
  p.C = (int)C;
  p.n_win = (int)n_win;
  p.hop = (int)hop;
  p.KB = KB;
  p.chunk_len = (KB + bp::kMaxChunks - 1) / bp::kMaxChunks;
  p.n_chunks = (KB + p.chunk_len - 1) / p.chunk_len;
  p.n_bands = n_bands;
  p.ct = (int)((C + 127) / 128);
  p.items = n_rec * n_win * p.ct;
  p.tab = tab;
  p.power = power;
  const TensorView3 tx{rec, {(unsigned long long)n_samples, (unsigned long long)C, (unsigned long long)n_rec},
                       {(unsigned long long)n_samples * 4, (unsigned long long)(C * n_samples) * 4}};
  const TensorView3 th{thi, {(unsigned long long)Kpad, 64, 1}, {(unsigned long long)Kpad * 4, (unsigned long long)Kpad * 64 * 4}};
  const TensorView3 tl{tlo, {(unsigned long long)Kpad, 64, 1}, {(unsigned long long)Kpad * 4, (unsigned long long)Kpad * 64 * 4}};
  CUtensorMap mx, mh, ml;
  rc = encode_tmap(&mx, tx, 32, 128, 0);
  if (rc == XM_OK) rc = encode_tmap(&mh, th, 32, 64, 0);
  if (rc == XM_OK) rc = encode_tmap(&ml, tl, 32, 64, 0);
  if (rc != XM_OK) return rc;
  // half-overlapping windows on 64-sample-aligned lengths: the overlap kernel (XM_BP_DFT_KERNEL=1 keeps v1 for A/B runs)
  static const bool force_v1 = [] { const char* e = getenv("XM_BP_DFT_KERNEL"); return e != nullptr && e[0] == '1'; }();
  const bool overlap = !force_v1 && win % 64 == 0 && hop * 2 == win && n_win >= 2;
  cudaError_t e;
  if (!overlap) {
    const int ctas = (int)(p.items < kNumSMs ? p.items : kNumSMs);
    e = cudaFuncSetAttribute(bp::bandpower_dft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bp::kSmem);
    if (e == cudaSuccess) bp::bandpower_dft_kernel<<<ctas, bp::kThreads, bp::kSmem, st>>>(mx, mh, ml, p);
  } else {
    bp::Params2 q{};
    q.C = (int)C;
    q.n_win = (int)n_win;
    q.HB = KB / 2;
    static const int max_chunk = [] { const char* c = getenv("XM_BP_CHUNK"); const int v = c ? atoi(c) : 16; return v < 1 ? 1 : v; }();  // A/B knob
    q.chunks_per_seg = (q.HB + max_chunk - 1) / max_chunk;
    q.chunk_len = (q.HB + q.chunks_per_seg - 1) / q.chunks_per_seg;
    q.chunks_per_seg = (q.HB + q.chunk_len - 1) / q.chunk_len;
    q.n_bands = n_bands;
    q.ct = p.ct;
    // windows per item: every item pays one extra hop, and the items should fill the 148 CTAs' rounds evenly
    long long best_cost = -1;
    for (int w = 4; w <= 64; ++w) {
      const int wpi = (int)(n_win < w ? n_win : w);
      const long long ipr = (n_win + wpi - 1) / wpi, items = n_rec * p.ct * ipr;
      const long long rounds = (items + kNumSMs - 1) / kNumSMs, cost = rounds * (wpi + 1);
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        q.wpi = wpi;
      }
    }
    q.ipr = (int)((n_win + q.wpi - 1) / q.wpi);
    q.items = n_rec * q.ct * q.ipr;
    q.tab = tab;
    q.power = power;
    const int ctas = (int)(q.items < kNumSMs ? q.items : kNumSMs);
    e = cudaFuncSetAttribute(bp::bandpower_dft2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bp::kSmem2);
    if (e == cudaSuccess) bp::bandpower_dft2_kernel<<<ctas, bp::kThreads2, bp::kSmem2, st>>>(mx, mh, ml, q);
  }
  if (e != cudaSuccess) {
    g_last_cuda_error = (int)e;
    return XM_ERR_LAUNCH;
  }
  return check_launch();
}

}  // extern "C"
