// EEG preprocessing on the device: window index generation, window gather, and the fused
// window-gather + taper + real FFT + one-sided PSD + band-power reduction.
//
// Band-power kernel: one warp per (window, channel) row.  The row is read straight from the
// recording (the gather is the address computation; windows are never materialised), packed as
// N2 = nfft/2 complex points, transformed by radix-8/4/2 Stockham passes that exchange through
// a padded per-warp shared-memory buffer, then only the bins inside the requested bands are
// unpacked (real-FFT post-processing), squared and reduced with warp shuffles.
// Algorithmic traffic: C*win*4 B read + C*n_bands*4 B written per window.
#include "spectral_core.cuh"
#include "xm_common.cuh"

namespace xm {

constexpr int kBpWarps = 8;      // rows in flight per CTA
constexpr int kMaxBands = 8;

__global__ void window_index_kernel(long long n_rec, long long n_win, long long hop, const long long* __restrict__ rec_labels,
                                    const long long* __restrict__ rec_subjects, long long* __restrict__ starts,
                                    long long* __restrict__ rec_ids, long long* __restrict__ labels,
                                    long long* __restrict__ subjects) {
  const long long total = n_rec * n_win;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const long long r = g / n_win, w = g - r * n_win;
    if (starts) starts[g] = w * hop;
    if (rec_ids) rec_ids[g] = r;
    if (labels && rec_labels) labels[g] = rec_labels[r];
    if (subjects && rec_subjects) subjects[g] = rec_subjects[r];
  }
}

// one block per (window, channel) row; coalesced copy.  out (G, C, ld_out >= win)
__global__ void window_gather_kernel(const float* __restrict__ rec, long long C, long long n_samples, long long n_win,
                                     long long win, long long hop, float* __restrict__ out, long long ld_out,
                                     int round_out) {
  const long long row = blockIdx.x;  // g*C + c
  const long long g = row / C, c = row - g * C;
  const long long r = g / n_win, w = g - r * n_win;
  const float* src = rec + (r * C + c) * n_samples + w * hop;
  float* dst = out + row * ld_out;
  for (long long t = threadIdx.x; t < win; t += blockDim.x) {
    const float v = src[t];
    dst[t] = round_out ? round_tf32(v) : v;
  }
}

// Transposing gather to channels-last: out (G, win, ld_out >= C).  32x32 tiles through shared
// memory so both the read (along time) and the write (along channels) are coalesced.
// grid (G, ceil(win/32), ceil(C/32)), block 32 x 8.
__global__ void window_gather_nwc_kernel(const float* __restrict__ rec, long long C, long long n_samples,
                                         long long n_win, long long win, long long hop, float* __restrict__ out,
                                         long long ld_out, int round_out) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long g = blockIdx.x;
  const long long r = g / n_win, w = g - r * n_win;
  const long long t0 = blockIdx.y * 32ll, c0 = blockIdx.z * 32ll;
  const float* src = rec + (r * C) * n_samples + w * hop;
#pragma unroll
  for (int j = ty; j < 32; j += 8) {
    const long long c = c0 + j, t = t0 + tx;
    tile[j][tx] = (c < C && t < win) ? src[c * n_samples + t] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int j = ty; j < 32; j += 8) {
    const long long t = t0 + j, c = c0 + tx;
    if (t < win && c < C) {
      const float v = tile[tx][j];
      float* o = out + (g * win + t) * ld_out + c;
      if (round_out == 2) {  // tf32 split along the channel axis [hi | lo]: operand of a 3-pass conv, which reads
        const float hi = round_tf32(v);  // [hi | lo | hi] by wrapping its third channel block back onto the first
        o[0] = hi;
        o[C] = round_tf32(v - hi);
      } else {
        o[0] = round_out ? round_tf32(v) : v;
      }
    }
  }
}

// Vectorised channels-last gather: a CTA transposes a (32 channels x 128 samples) tile through shared memory -- 512-B
// row segments in (one float4 per lane), whole 128-B channel segments out (float4 stores; the [hi | lo] split writes two
// of them).  The scalar 32 x 32 kernel above moved 4 bytes per thread and instruction and ran at 3.1 TB/s.
__global__ void __launch_bounds__(256)
window_gather_nwc_v4_kernel(const float* __restrict__ rec, long long C, long long n_samples, long long n_win, long long win,
                            long long hop, float* __restrict__ out, long long ld_out, int round_out) {
  __shared__ float tile[32][129];  // row stride 129: bank = (channel + sample) mod 32 in both phases
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long g = blockIdx.x;
  const long long r = g / n_win, w = g - r * n_win;
  const long long t0 = blockIdx.y * 128ll, c0 = blockIdx.z * 32ll;
  const float* src = rec + (r * C) * n_samples + w * hop + t0;
#pragma unroll
  for (int j = warp; j < 32; j += 8) {  // channel row j of the tile: 128 samples = 32 lanes x float4
    const long long c = c0 + j, t = t0 + 4 * lane;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < C && t < win) v = *reinterpret_cast<const float4*>(src + c * n_samples + 4 * lane);  // win % 4 == 0: all or nothing
    tile[j][4 * lane + 0] = v.x;
    tile[j][4 * lane + 1] = v.y;
    tile[j][4 * lane + 2] = v.z;
    tile[j][4 * lane + 3] = v.w;
  }
  __syncthreads();
  const int cq = lane & 7, tq = lane >> 3;  // 8 lanes x float4 = the tile's 32 channels of one sample; 4 samples per warp pass
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int tt = warp * 16 + i * 4 + tq;  // warp covers samples [16 warp, 16 warp + 16)
    const long long t = t0 + tt, c = c0 + 4 * cq;
    if (t < win && c < C) {  // C % 4 == 0: all or nothing
      float4 v = make_float4(tile[4 * cq + 0][tt], tile[4 * cq + 1][tt], tile[4 * cq + 2][tt], tile[4 * cq + 3][tt]);
      float* o = out + (g * win + t) * ld_out + c;
      if (round_out == 2) {
        const float4 hi = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
        *reinterpret_cast<float4*>(o) = hi;
        *reinterpret_cast<float4*>(o + C) = make_float4(round_tf32(v.x - hi.x), round_tf32(v.y - hi.y), round_tf32(v.z - hi.z),
                                                        round_tf32(v.w - hi.w));
      } else {
        if (round_out) v = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
        *reinterpret_cast<float4*>(o) = v;
      }
    }
  }
}


// Ping-pong Stockham passes: pass p gathers from one per-warp buffer (pass 0: from global,
// applying the taper and the zero padding) and scatters to the other, so a pass with more than
// 32 butterflies never overwrites a source another lane group still has to read.
// Raw samples of one row held in registers: lane j keeps, for butterfly j0 + j of pass 0 (radix R0),
// the R0 packed complex inputs n = j + r*NB, i.e. samples (2n, 2n+1).  Loaded one row AHEAD of the row
// being transformed, so the HBM latency of row i+1 hides behind the butterflies of row i.
template <int N2>
struct RowRegs {
  static constexpr int R0 = fft::radix_for(N2);
  static constexpr int NB = N2 / R0;
  static constexpr int NJ = (NB + 31) / 32;
  float2 v[NJ][R0];
};

template <int N2>
XM_DEVICE void load_row(RowRegs<N2>& rr, const float* __restrict__ row, int win, int lane) {
  constexpr int R0 = RowRegs<N2>::R0, NB = RowRegs<N2>::NB, NJ = RowRegs<N2>::NJ;
  const bool aligned8 = (reinterpret_cast<uintptr_t>(row) & 7) == 0;
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj) {
    const int j = jj * 32 + lane;
#pragma unroll
    for (int r = 0; r < R0; ++r) {
      const int n = j + r * NB;
      float2 x = make_float2(0.f, 0.f);
      if ((NB >= 32 || j < NB) && 2 * n < win) {
        if (2 * n + 1 < win) {
          if (aligned8) x = __ldg(reinterpret_cast<const float2*>(row + 2 * n));
          else { x.x = __ldg(row + 2 * n); x.y = __ldg(row + 2 * n + 1); }
        } else {
          x.x = __ldg(row + 2 * n);
        }
      }
      rr.v[jj][r] = x;
    }
  }
}

// Ping-pong Stockham passes: pass p gathers from one per-warp buffer (pass 0: from the row registers,
// applying the taper; samples beyond `win` are the zero padding) and scatters to the other, so a pass
// with more than 32 butterflies never overwrites a source another lane group still has to read.
template <int N2, int Ns>
struct PassesPP {
  static constexpr int rem = N2 / Ns;
  static constexpr int R = fft::radix_for(rem);
  static XM_DEVICE void run(const RowRegs<N2>& rr, const float* __restrict__ taper, int win,
                            float* are, float* aim, float* bre, float* bim, const float* __restrict__ tw_re,
                            const float* __restrict__ tw_im, int lane) {
    // reads (are, aim) [or the row registers when Ns == 1], writes (bre, bim)
    constexpr int NB = N2 / R;
#pragma unroll
    for (int j0 = 0; j0 < NB; j0 += 32) {
      const int j = j0 + lane;
      if (NB >= 32 || j < NB) {
        float re[R], im[R];
        if (Ns == 1) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int n = j + r * NB;
            float2 tp = make_float2(0.f, 0.f);
            if (2 * n + 1 < win) tp = *reinterpret_cast<const float2*>(taper + 2 * n);
            else if (2 * n < win) tp.x = taper[2 * n];
            const float2 x = rr.v[j0 / 32][r];
            re[r] = x.x * tp.x; im[r] = x.y * tp.y;
          }
        } else {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int i = fft::pad_idx(j + r * NB);
            re[r] = are[i]; im[r] = aim[i];
          }
        }
        fft::twiddle_and_butterfly<R>(re, im, j % Ns, Ns, N2, tw_re, tw_im);
        const int b = fft::scatter_base(j, Ns, R);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int i = fft::pad_idx(b + r * Ns);
          bre[i] = re[r]; bim[i] = im[r];
        }
      }
    }
    __syncwarp();
    PassesPP<N2, Ns * R>::run(rr, taper, win, bre, bim, are, aim, tw_re, tw_im, lane);
  }
  // number of passes from this Ns on (to know which buffer holds the result)
  static constexpr int count = 1 + PassesPP<N2, Ns * R>::count;
};
template <int N2>
struct PassesPP<N2, N2> {
  static XM_DEVICE void run(const RowRegs<N2>&, const float*, int, float*, float*, float*, float*, const float*,
                            const float*, int) {}
  static constexpr int count = 0;
};

template <int N2>
__global__ void __launch_bounds__(kBpWarps * 32)
bandpower_kernel(const float* __restrict__ rec, long long n_rows, long long C, long long n_samples, long long n_win,
                 int win, long long hop, const float* __restrict__ taper, float scale,
                 const int* __restrict__ band_bins, int n_bands, float* __restrict__ power) {
  constexpr int PL = N2 + (N2 >> 5) + 1;
  extern __shared__ float smem[];
  float* tw_re = smem;             // N2
  float* tw_im = smem + N2;        // N2
  float* bufs = smem + 2 * N2;     // per warp: 4 * PL (A.re, A.im, B.re, B.im)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int m = threadIdx.x; m < N2; m += blockDim.x) {
    float s, c;
    sincospif(-2.0f * (float)m / (float)N2, &s, &c);
    tw_re[m] = c; tw_im[m] = s;
  }
  __syncthreads();
  float* are = bufs + warp * 4 * PL;
  float* aim = are + PL;
  float* bre = aim + PL;
  float* bim = bre + PL;
  // the first pass writes B, the second A, ...: result buffer by pass-count parity
  constexpr int NP = PassesPP<N2, 1>::count;
  float* zre = (NP & 1) ? bre : are;
  float* zim = (NP & 1) ? bim : aim;

  int lo = 1 << 30, hi = 0;
  for (int b = 0; b < n_bands; ++b) {
    lo = min(lo, band_bins[2 * b]);
    hi = max(hi, band_bins[2 * b + 1]);
  }
  const int nfft = 2 * N2;

  const long long row_stride = (long long)gridDim.x * kBpWarps;
  auto row_ptr = [&](long long row) {
    const long long g = row / C, c = row - g * C;
    const long long r = g / n_win, w = g - r * n_win;
    return rec + (r * C + c) * n_samples + w * hop;
  };
  long long row = blockIdx.x * (long long)kBpWarps + warp;
  RowRegs<N2> cur;
  if (row < n_rows) load_row<N2>(cur, row_ptr(row), win, lane);
  for (; row < n_rows; row += row_stride) {
    RowRegs<N2> nxt;
    const bool has_next = row + row_stride < n_rows;
    if (has_next) load_row<N2>(nxt, row_ptr(row + row_stride), win, lane);  // in flight during this row's FFT
    PassesPP<N2, 1>::run(cur, taper, win, are, aim, bre, bim, tw_re, tw_im, lane);

    float acc[kMaxBands];
#pragma unroll
    for (int b = 0; b < kMaxBands; ++b) acc[b] = 0.f;
    for (int k = lo + lane; k < hi; k += 32) {
      // X[k] = 0.5*[(Z[k] + conj(Z[N2-k])) - i*e^{-2 pi i k/nfft} * (Z[k] - conj(Z[N2-k]))]
      const int k1 = (k == N2) ? 0 : k;
      const int k2 = (k == 0 || k == N2) ? 0 : N2 - k;
      const float ar = zre[fft::pad_idx(k1)], ai = zim[fft::pad_idx(k1)];
      const float br = zre[fft::pad_idx(k2)], bi = -zim[fft::pad_idx(k2)];
      const float er = 0.5f * (ar + br), ei = 0.5f * (ai + bi);
      const float dr = 0.5f * (ar - br), di = 0.5f * (ai - bi);
      float s, co;
      sincospif(-2.0f * (float)k / (float)nfft, &s, &co);
      // -i * (co + i s) * (dr + i di) = (co*di + s*dr) + i*(s*di - co*dr)
      const float xr = er + (co * di + s * dr);
      const float xi = ei + (s * di - co * dr);
      float pw = xr * xr + xi * xi;
      if (k != 0 && k != N2) pw *= 2.0f;
#pragma unroll
      for (int b = 0; b < kMaxBands; ++b)
        if (b < n_bands && k >= band_bins[2 * b] && k < band_bins[2 * b + 1]) acc[b] += pw;
    }
#pragma unroll
    for (int b = 0; b < kMaxBands; ++b) {
      if (b < n_bands) {
        const float t = warp_sum(acc[b]);
        if (lane == 0) power[row * n_bands + b] = t * scale;
      }
    }
    __syncwarp();
    if (has_next) cur = nxt;
  }
}

template <int N2>
static int launch_bandpower(const float* rec, long long n_rows, long long C, long long n_samples, long long n_win, int win,
                            long long hop, const float* taper, float scale, const int* band_bins, int n_bands,
                            float* power, cudaStream_t st) {
  constexpr int PL = N2 + (N2 >> 5) + 1;
  const size_t smem = (size_t)(2 * N2 + kBpWarps * 4 * PL) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(bandpower_kernel<N2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    g_last_cuda_error = (int)e;
    return XM_ERR_LAUNCH;
  }
  long long blocks = (n_rows + kBpWarps - 1) / kBpWarps;
  const long long cap = (long long)kNumSMs * 8;  // persistent-ish: a few CTAs per SM, grid-stride over rows
  if (blocks > cap) blocks = cap;
  bandpower_kernel<N2><<<(unsigned)blocks, kBpWarps * 32, smem, st>>>(rec, n_rows, C, n_samples, n_win, win, hop, taper,
                                                                      scale, band_bins, n_bands, power);
  return check_launch();
}

}  // namespace xm

using namespace xm;

extern "C" {

int xm_window_index_i64(int64_t n_rec, int64_t n_samples, int64_t win, int64_t hop, const int64_t* rec_labels,
                        const int64_t* rec_subjects, int64_t* starts, int64_t* rec_ids, int64_t* labels,
                        int64_t* subjects, void* stream) {
  if (n_rec <= 0 || win <= 0 || hop <= 0 || n_samples < win) return XM_ERR_INVALID;
  const long long n_win = (n_samples - win) / hop + 1;
  const long long total = n_rec * n_win;
  long long blocks = (total + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  window_index_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      n_rec, n_win, hop, (const long long*)rec_labels, (const long long*)rec_subjects, (long long*)starts,
      (long long*)rec_ids, (long long*)labels, (long long*)subjects);
  return check_launch();
}

int xm_window_gather_f32(const float* rec, int64_t n_rec, int64_t C, int64_t n_samples, int64_t win, int64_t hop,
                         float* out, int64_t ld_out, int channels_last, int round_tf32, void* stream) {
  if (!rec || !out || n_rec <= 0 || C <= 0 || win <= 0 || hop <= 0 || n_samples < win) return XM_ERR_INVALID;
  const long long n_win = (n_samples - win) / hop + 1;
  const long long G = n_rec * n_win;
  if (round_tf32 == 2 && !channels_last) return XM_ERR_UNSUPPORTED;
  if (channels_last) {
    if (ld_out < (round_tf32 == 2 ? 2 * C : C)) return XM_ERR_INVALID;
    if (G > 2147483647ll || ceil_div(win, 32) > 65535 || ceil_div(C, 32) > 65535) return XM_ERR_UNSUPPORTED;
    const bool v4 = ((n_samples | win | C | ld_out) & 3) == 0 && (n_win == 1 || (hop & 3) == 0) &&
                    ((reinterpret_cast<uintptr_t>(rec) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (v4) {
      dim3 grid4((unsigned)G, ceil_div(win, 128), ceil_div(C, 32));
      window_gather_nwc_v4_kernel<<<grid4, 256, 0, (cudaStream_t)stream>>>(rec, C, n_samples, n_win, win, hop, out, ld_out,
                                                                           round_tf32);
      return check_launch();
    }
    dim3 grid((unsigned)G, ceil_div(win, 32), ceil_div(C, 32));
    window_gather_nwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(rec, C, n_samples, n_win, win, hop, out, ld_out,
                                                                     round_tf32);
    return check_launch();
  }
  if (ld_out < win) return XM_ERR_INVALID;
  const long long rows = G * C;
  if (rows > 2147483647ll) return XM_ERR_UNSUPPORTED;
  window_gather_kernel<<<(unsigned)rows, 128, 0, (cudaStream_t)stream>>>(rec, C, n_samples, n_win, win, hop, out, ld_out,
                                                                         round_tf32);
  return check_launch();
}

int xm_bandpower_f32(const float* rec, int64_t n_rec, int64_t C, int64_t n_samples, int64_t win, int64_t hop,
                     int64_t nfft, float fs, const float* taper, float taper_sumsq, const int32_t* band_bins,
                     int n_bands, float* power, void* stream) {
  (void)fs;  // PSD * bin width: fs cancels; kept in the signature for the reference formula
  if (!rec || !taper || !band_bins || !power || n_rec <= 0 || C <= 0 || win <= 0 || hop <= 0 || n_samples < win)
    return XM_ERR_INVALID;
  if (n_bands <= 0 || n_bands > kMaxBands || nfft < win || !(taper_sumsq > 0.f)) return XM_ERR_INVALID;
  if (reinterpret_cast<uintptr_t>(taper) & 7) return XM_ERR_INVALID;
  const long long n_win = (n_samples - win) / hop + 1;
  const long long rows = n_rec * n_win * C;
  const float scale = 1.0f / ((float)nfft * taper_sumsq);
  cudaStream_t st = (cudaStream_t)stream;
  switch (nfft) {
    case 64: return launch_bandpower<32>(rec, rows, C, n_samples, n_win, (int)win, hop, taper, scale, band_bins, n_bands, power, st);
    case 128: return launch_bandpower<64>(rec, rows, C, n_samples, n_win, (int)win, hop, taper, scale, band_bins, n_bands, power, st);
    case 256: return launch_bandpower<128>(rec, rows, C, n_samples, n_win, (int)win, hop, taper, scale, band_bins, n_bands, power, st);
    case 512: return launch_bandpower<256>(rec, rows, C, n_samples, n_win, (int)win, hop, taper, scale, band_bins, n_bands, power, st);
    case 1024: return launch_bandpower<512>(rec, rows, C, n_samples, n_win, (int)win, hop, taper, scale, band_bins, n_bands, power, st);
    case 2048: return launch_bandpower<1024>(rec, rows, C, n_samples, n_win, (int)win, hop, taper, scale, band_bins, n_bands, power, st);
  }
  return XM_ERR_UNSUPPORTED;
}

}  // extern "C"
