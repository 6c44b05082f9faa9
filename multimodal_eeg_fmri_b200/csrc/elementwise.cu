// HBM-bound kernels of the paired step: fMRI ROI aggregation, z-scoring, train-mode
// BatchNorm (+GELU/ReLU, +MaxPool1d(2), +dropout) forward/backward on channels-last (B, T, C)
// activations (and (B, C) nn.Linear outputs, T = 1), LayerNorm(+act, +dropout) on (M, D) rows,
// L2 row normalisation, small reductions.  Every kernel is a streaming pass with coalesced
// accesses and fp32 math; batch statistics are combined in fp64.
#include "xm_common.cuh"

namespace xm {

static int grid_rows(long long rows) { return (int)(rows < 1 ? 1 : (rows > 2147483647ll ? 2147483647ll : rows)); }

// ------------------------------------------------------------------ block reduce helpers
template <typename T>
XM_DEVICE T block_sum(T v, T* sm) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sm[w] = v;
  __syncthreads();
  T r = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : (T)0;
  if (w == 0) r = warp_sum(r);
  if (threadIdx.x == 0) sm[0] = r;
  __syncthreads();
  return sm[0];
}

// ------------------------------------------------------------------ fMRI ROI mean/std over TR
// fMRI_CODE/fmri_utils.py:140-147: nan_to_num then concat(mean(axis=0), std(axis=0)) (ddof=0).
__global__ void roi_meanstd_kernel(const float* __restrict__ x, long long TR, long long ROI, float* __restrict__ out) {
  const long long b = blockIdx.y;
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r >= ROI) return;
  const float* col = x + b * TR * ROI + r;
  float s = 0.f;
  for (long long t = 0; t < TR; ++t) {
    float v = col[t * ROI];
    s += (v != v) ? 0.f : v;
  }
  const float mean = s / (float)TR;
  float q = 0.f;
  for (long long t = 0; t < TR; ++t) {
    float v = col[t * ROI];
    v = (v != v) ? 0.f : v;
    const float d = v - mean;
    q += d * d;
  }
  out[b * 2 * ROI + r] = mean;
  out[b * 2 * ROI + ROI + r] = sqrtf(q / (float)TR);
}

// Vectorised variant (ROI % 4 == 0, 16-B aligned): one CTA per sample, thread (c, g) owns the float4 of ROIs 4c..4c+3 and
// the TRs g, g + G, ...; the G partial sums per column meet in shared memory.  Both passes (mean, then the centred sum
// of squares, as the reference's two-pass std) read 128-bit words; the second pass hits in L1 / L2.  The scalar kernel
// above moved 4 bytes per thread and instruction with 72 of 128 threads of its second block idle at ROI = 200 (1.9 TB/s).
constexpr int kRoiThreads = 256;
__global__ void __launch_bounds__(kRoiThreads)
roi_meanstd_v4_kernel(const float* __restrict__ x, int TR, int ROI4, int G, float* __restrict__ out) {
  extern __shared__ float4 part[];  // [G][ROI4]
  const int c = threadIdx.x % ROI4, g = threadIdx.x / ROI4;
  const long long b = blockIdx.x;
  const float4* base = reinterpret_cast<const float4*>(x) + b * (long long)TR * ROI4 + c;
  const bool active = g < G;
  auto clean = [](float4 v) {
    return make_float4(v.x != v.x ? 0.f : v.x, v.y != v.y ? 0.f : v.y, v.z != v.z ? 0.f : v.z, v.w != v.w ? 0.f : v.w);
  };
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (active) {
#pragma unroll 4
    for (int t = g; t < TR; t += G) {
      const float4 v = clean(base[(long long)t * ROI4]);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    part[g * ROI4 + c] = s;
  }
  __syncthreads();
  float4 mean = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = 0; i < G; ++i) {  // every thread folds its column's partials in the same order
    const float4 p = part[i * ROI4 + c];
    mean.x += p.x; mean.y += p.y; mean.z += p.z; mean.w += p.w;
  }
  const float inv = 1.0f / (float)TR;
  mean.x *= inv; mean.y *= inv; mean.z *= inv; mean.w *= inv;
  __syncthreads();
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
  if (active) {
#pragma unroll 4
    for (int t = g; t < TR; t += G) {
      const float4 v = clean(base[(long long)t * ROI4]);
      const float dx = v.x - mean.x, dy = v.y - mean.y, dz = v.z - mean.z, dw = v.w - mean.w;
      q.x += dx * dx; q.y += dy * dy; q.z += dz * dz; q.w += dw * dw;
    }
    part[g * ROI4 + c] = q;
  }
  __syncthreads();
  if (g == 0) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = 0; i < G; ++i) {
      const float4 p = part[i * ROI4 + c];
      t.x += p.x; t.y += p.y; t.z += p.z; t.w += p.w;
    }
    float4* o = reinterpret_cast<float4*>(out + b * 8ll * ROI4);
    o[c] = mean;
    o[ROI4 + c] = make_float4(sqrtf(t.x * inv), sqrtf(t.y * inv), sqrtf(t.z * inv), sqrtf(t.w * inv));
  }
}

// ------------------------------------------------------------------ per-item z-score
// EEG_CODE/run_training_lite.py:48-51
__global__ void zscore_kernel(const float* __restrict__ x, long long len, float eps, float* __restrict__ out) {
  __shared__ double smd[32];
  const float* xi = x + blockIdx.x * len;
  float* oi = out + blockIdx.x * len;
  double s = 0.0;
  for (long long i = threadIdx.x; i < len; i += blockDim.x) s += (double)xi[i];
  const double mean = block_sum(s, smd) / (double)len;
  double q = 0.0;
  for (long long i = threadIdx.x; i < len; i += blockDim.x) {
    const double d = (double)xi[i] - mean;
    q += d * d;
  }
  const double var = block_sum(q, smd) / (double)len;
  const float fmean = (float)mean;
  const float inv = 1.0f / ((float)sqrt(var) + eps);
  for (long long i = threadIdx.x; i < len; i += blockDim.x) oi[i] = (xi[i] - fmean) * inv;
}

// ------------------------------------------------------------------ BatchNorm (channels-last)
// Activations are (R, C) row-major with pitch ld: R = B*T rows of one time step (or one sample for
// nn.Linear outputs, T = 1), C channels contiguous.  Per-channel statistics are column
// reductions; MaxPool1d(2) pairs rows (b, 2t') and (b, 2t'+1).
//
// grid (ceil(C/32), nsplit), block 32 x 8: block (cx, s) reduces rows [s*rps, (s+1)*rps) of 32 columns.
__global__ void bn_partial_stats_kernel(const float* __restrict__ y, long long R, int C, long long ld,
                                        long long rows_per_split, double* __restrict__ partials) {
  __shared__ double sm[2][8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const long long r0 = blockIdx.y * rows_per_split;
  const long long r1 = min(R, r0 + rows_per_split);
  double sum = 0.0, sq = 0.0;
  if (c < C) {
    float ps = 0.f, pq = 0.f;
    int n = 0;
    for (long long r = r0 + ty; r < r1; r += 8) {
      const float v = y[r * ld + c];
      ps += v;
      pq += v * v;
      if (++n == 64) {
        sum += (double)ps; sq += (double)pq;
        ps = 0.f; pq = 0.f; n = 0;
      }
    }
    sum += (double)ps;
    sq += (double)pq;
  }
  sm[0][ty][tx] = sum;
  sm[1][ty][tx] = sq;
  __syncthreads();
  if (ty == 0 && c < C) {
    double s = 0.0, q = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s += sm[0][i][tx];
      q += sm[1][i][tx];
    }
    partials[((long long)blockIdx.y * C + c) * 2 + 0] = s;
    partials[((long long)blockIdx.y * C + c) * 2 + 1] = q;
  }
}

// one warp per channel: lanes stride over the splits, fp64 shuffle reduction
__global__ void bn_finalize_stats_kernel(const double* __restrict__ partials, int nsplit, int C, double n, float eps,
                                         float* __restrict__ mean, float* __restrict__ invstd, float* running_mean,
                                         float* running_var, float momentum) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= C) return;
  const int lane = threadIdx.x & 31;
  double s = 0.0, q = 0.0;
  for (int i = lane; i < nsplit; i += 32) {
    s += partials[((long long)i * C + c) * 2 + 0];
    q += partials[((long long)i * C + c) * 2 + 1];
  }
  s = warp_sum(s);
  q = warp_sum(q);
  if (lane != 0) return;
  const double mu = s / n;
  double var = q / n - mu * mu;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)mu;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * (float)mu;
  if (running_var) {
    const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
    running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

struct BnActArgs {
  const float* y;
  const float* mean;
  const float* invstd;
  const float* gamma;
  const float* beta;
  long long B, T, ldy, ldo;  // T rows per sample; R = B*T
  int C, act, pool, drop_before_pool, round_out;
  float drop_scale;      // 1/(1-p), 1 when p == 0
  uint32_t drop_thresh;  // p * 2^32, 0 when p == 0
  uint64_t seed;
};

// Dropout mask of the BatchNorm blocks: ONE hash per group of four consecutive elements (a float4 of channels), then an
// LCG step per element (h' = h * 747796405 + 2891336453, the stream the fused attention / FFN kernels use): element
// idx is kept iff state (idx & 3) + 1 of the stream seeded by hash_u32(idx >> 2, seed) is >= p * 2^32.  The streaming
// kernels around GELU are instruction-issue bound and a full hash per element was a quarter of their instructions.
XM_DEVICE uint32_t lcg_next(uint32_t h) { return h * 747796405u + 2891336453u; }
XM_DEVICE float drop_mul(const BnActArgs& a, long long idx) {  // scalar kernels (any C)
  if (a.drop_thresh == 0u) return 1.0f;
  uint32_t h = hash_u32((uint64_t)(idx >> 2), a.seed);
  const int j = (int)(idx & 3);
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i <= j) h = lcg_next(h);
  return h >= a.drop_thresh ? a.drop_scale : 0.0f;
}
struct Mul4 {
  float v[4];
};
XM_DEVICE Mul4 drop_mul4(const BnActArgs& a, long long idx0) {  // idx0 % 4 == 0: the four elements of one float4
  Mul4 m;
  uint32_t h = hash_u32((uint64_t)(idx0 >> 2), a.seed);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h = lcg_next(h);
    m.v[j] = h >= a.drop_thresh ? a.drop_scale : 0.0f;
  }
  return m;
}

// forward value and dz for one channel of one output position.
//   pool == 0: input row r  -> out row r
//   pool == 2: input rows (b*T + 2tp, +1) -> out row b*To + tp
XM_DEVICE float bn_act_fwd_elem(const BnActArgs& a, long long ro, int c, float sc, float sh) {
  if (a.pool == 2) {
    const long long To = a.T / 2, b = (long long)((unsigned)ro / (unsigned)To), tp = ro - b * To;
    const long long r0 = b * a.T + 2 * tp;
    float a0 = apply_act(a.y[r0 * a.ldy + c] * sc + sh, a.act);
    float a1 = apply_act(a.y[(r0 + 1) * a.ldy + c] * sc + sh, a.act);
    if (a.drop_before_pool) {
      a0 *= drop_mul(a, r0 * a.C + c);
      a1 *= drop_mul(a, (r0 + 1) * a.C + c);
      return fmaxf(a0, a1);
    }
    return fmaxf(a0, a1) * drop_mul(a, ro * a.C + c);
  }
  return apply_act(a.y[ro * a.ldy + c] * sc + sh, a.act) * drop_mul(a, ro * a.C + c);
}

// dz (gradient wrt z = gamma*xhat + beta) of the (up to) two input rows behind output row ro.
XM_DEVICE void bn_act_dz(const BnActArgs& a, const float* __restrict__ dout, long long ro, int c, float sc, float sh,
                         float& x0, float& x1, float& dz0, float& dz1) {
  const float g = dout[ro * a.ldo + c];
  if (a.pool == 2) {
    const long long To = a.T / 2, b = (long long)((unsigned)ro / (unsigned)To), tp = ro - b * To;
    const long long r0 = b * a.T + 2 * tp;
    x0 = a.y[r0 * a.ldy + c];
    x1 = a.y[(r0 + 1) * a.ldy + c];
    const float z0 = x0 * sc + sh, z1 = x1 * sc + sh;
    float a0 = apply_act(z0, a.act), a1 = apply_act(z1, a.act);
    float m0 = 1.f, m1 = 1.f, gg = g;
    if (a.drop_before_pool) {
      m0 = drop_mul(a, r0 * a.C + c);
      m1 = drop_mul(a, (r0 + 1) * a.C + c);
      a0 *= m0;
      a1 *= m1;
    } else {
      gg *= drop_mul(a, ro * a.C + c);
    }
    const bool first = a0 >= a1;  // ties -> first element, as torch max_pool1d
    dz0 = first ? gg * m0 * act_grad(z0, a.act) : 0.f;
    dz1 = first ? 0.f : gg * m1 * act_grad(z1, a.act);
  } else {
    x0 = a.y[ro * a.ldy + c];
    x1 = 0.f;
    dz0 = g * drop_mul(a, ro * a.C + c) * act_grad(x0 * sc + sh, a.act);
    dz1 = 0.f;
  }
}

// one thread per (output row, channel); channels fastest => coalesced
__global__ void bn_act_fwd_kernel(const BnActArgs a, long long R_out, float* __restrict__ out) {
  const long long total = R_out * a.C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long ro = i / a.C;
    const int c = (int)(i - ro * a.C);
    const float sc = a.gamma[c] * a.invstd[c], sh = a.beta[c] - a.mean[c] * sc;
    const float o = bn_act_fwd_elem(a, ro, c, sc, sh);
    if (a.round_out == 2) {  // tf32 split along the channel axis [hi | lo] (operand of a 3-pass conv)
      const float hi = round_tf32(o);
      out[ro * a.ldo + c] = hi;
      out[ro * a.ldo + a.C + c] = round_tf32(o - hi);
    } else {
      out[ro * a.ldo + c] = a.round_out ? round_tf32(o) : o;
    }
  }
}

// grid (ceil(C/32), nsplit), block 32 x 8 over OUTPUT rows: partial sums of dz and dz*xhat
__global__ void bn_act_bwd_reduce_kernel(const BnActArgs a, const float* __restrict__ dout, long long R_out,
                                         long long rows_per_split, double* __restrict__ partials) {
  __shared__ double sm[2][8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const long long r0 = blockIdx.y * rows_per_split;
  const long long r1 = min(R_out, r0 + rows_per_split);
  double sdz = 0.0, sdzx = 0.0;
  if (c < a.C) {
    const float mu = a.mean[c], is = a.invstd[c];
    const float sc = a.gamma[c] * is, sh = a.beta[c] - mu * sc;
    float p0 = 0.f, p1 = 0.f;
    int n = 0;
    for (long long ro = r0 + ty; ro < r1; ro += 8) {
      float x0, x1, dz0, dz1;
      bn_act_dz(a, dout, ro, c, sc, sh, x0, x1, dz0, dz1);
      p0 += dz0 + dz1;
      p1 += dz0 * ((x0 - mu) * is) + dz1 * ((x1 - mu) * is);
      if (++n == 64) {
        sdz += (double)p0; sdzx += (double)p1;
        p0 = 0.f; p1 = 0.f; n = 0;
      }
    }
    sdz += (double)p0;
    sdzx += (double)p1;
  }
  sm[0][ty][tx] = sdz;
  sm[1][ty][tx] = sdzx;
  __syncthreads();
  if (ty == 0 && c < a.C) {
    double s = 0.0, q = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s += sm[0][i][tx];
      q += sm[1][i][tx];
    }
    partials[((long long)blockIdx.y * a.C + c) * 2 + 0] = s;
    partials[((long long)blockIdx.y * a.C + c) * 2 + 1] = q;
  }
}

__global__ void bn_bwd_finalize_kernel(const double* __restrict__ partials, int nsplit, int C, float* __restrict__ dbeta,
                                       float* __restrict__ dgamma) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= C) return;
  const int lane = threadIdx.x & 31;
  double s = 0.0, q = 0.0;
  for (int i = lane; i < nsplit; i += 32) {
    s += partials[((long long)i * C + c) * 2 + 0];
    q += partials[((long long)i * C + c) * 2 + 1];
  }
  s = warp_sum(s);
  q = warp_sum(q);
  if (lane == 0) {
    dbeta[c] = (float)s;
    dgamma[c] = (float)q;
  }
}

// dy = gamma*invstd*(dz - dbeta/n - xhat*dgamma/n); one thread per (output row, channel)
__global__ void bn_act_bwd_apply_kernel(const BnActArgs a, const float* __restrict__ dout,
                                        const float* __restrict__ dbeta, const float* __restrict__ dgamma, float inv_n,
                                        long long R_out, float* __restrict__ dy) {
  const long long total = R_out * a.C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long ro = i / a.C;
    const int c = (int)(i - ro * a.C);
    const float mu = a.mean[c], is = a.invstd[c];
    const float sc = a.gamma[c] * is, sh = a.beta[c] - mu * sc;
    const float k1 = dbeta[c] * inv_n, k2 = dgamma[c] * inv_n;
    float x0, x1, dz0, dz1;
    bn_act_dz(a, dout, ro, c, sc, sh, x0, x1, dz0, dz1);
    if (a.pool == 2) {
      const long long To = a.T / 2, b = (long long)((unsigned)ro / (unsigned)To), tp = ro - b * To;
      const long long r0 = b * a.T + 2 * tp;
      float o0 = sc * (dz0 - k1 - (x0 - mu) * is * k2);
      float o1 = sc * (dz1 - k1 - (x1 - mu) * is * k2);
      if (a.round_out) { o0 = round_tf32(o0); o1 = round_tf32(o1); }
      dy[r0 * a.ldy + c] = o0;
      dy[(r0 + 1) * a.ldy + c] = o1;
      if ((a.T & 1) && tp == To - 1) {  // odd tail row is dropped by the pool: dz = 0
        const float xt = a.y[(r0 + 2) * a.ldy + c];
        const float ot = sc * (0.f - k1 - (xt - mu) * is * k2);
        dy[(r0 + 2) * a.ldy + c] = a.round_out ? round_tf32(ot) : ot;
      }
    } else {
      const float o = sc * (dz0 - k1 - (x0 - mu) * is * k2);
      dy[ro * a.ldy + c] = a.round_out ? round_tf32(o) : o;
    }
  }
}

// mean over the T rows of each sample: x (B, T, C) pitch ld -> out (B, C); grid (ceil(C/32), B), block 32 x 8
__global__ void seqmean_kernel(const float* __restrict__ x, long long T, int C, long long ld, float* __restrict__ out) {
  __shared__ float sm[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const long long b = blockIdx.y;
  float acc = 0.f;
  if (c < C)
    for (long long t = ty; t < T; t += 8) acc += x[(b * T + t) * ld + c];
  sm[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += sm[i][tx];
    out[b * C + c] = s / (float)T;
  }
}
__global__ void seqmean_bwd_kernel(const float* __restrict__ dout, long long T, int C, long long ld, long long total,
                                   float* __restrict__ dx) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    dx[r * ld + c] = dout[(r / T) * C + c] / (float)T;
  }
}

// ------------------------------------------------------------------ LayerNorm + act + dropout (warp per row)
__global__ void ln_act_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, float* __restrict__ out, float* __restrict__ mean,
                                  float* __restrict__ rstd, long long M, int D, float eps, int act, float drop_scale,
                                  uint32_t drop_thresh, uint64_t seed) {
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + row * D;
  float s = 0.f;
  for (int d = lane; d < D; d += 32) s += xr[d];
  const float mu = warp_sum(s) / (float)D;
  float q = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float v = xr[d] - mu;
    q += v * v;
  }
  const float rs = rsqrtf(warp_sum(q) / (float)D + eps);
  if (lane == 0) {
    mean[row] = mu;
    rstd[row] = rs;
  }
  for (int d = lane; d < D; d += 32) {
    float o = apply_act((xr[d] - mu) * rs * gamma[d] + beta[d], act);
    if (drop_thresh) o = dropout_keep((uint64_t)(row * D + d), seed, drop_thresh) ? o * drop_scale : 0.f;
    out[row * D + d] = o;
  }
}

// grid = nblk blocks of 8 warps; block accumulates dgamma/dbeta partials for its rows in smem.
__global__ void ln_act_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ x,
                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                  const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ dx,
                                  float* __restrict__ dgamma_part, float* __restrict__ dbeta_part, long long M, int D,
                                  int act, float drop_scale, uint32_t drop_thresh, uint64_t seed, int per_warp) {
  // per_warp: every warp owns a private (2, D) accumulator (element d is only ever touched by lane d % 32 of that warp,
  // in row order) and the block adds the warps in a fixed order -> bit-reproducible partials.  Wide rows (D > 768: the
  // private copies would not fit 48 KB) share one accumulator through shared-memory atomics.
  extern __shared__ float sm[];  // per_warp ? warps * 2 * D : 2 * D
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  float* sg = sm + (per_warp ? (threadIdx.x >> 5) * 2 * D : 0);
  float* sb = sg + D;
  for (int d = threadIdx.x; d < (per_warp ? wpb : 1) * 2 * D; d += blockDim.x) sm[d] = 0.f;
  __syncthreads();
  for (long long row = blockIdx.x * (long long)wpb + (threadIdx.x >> 5); row < M; row += (long long)gridDim.x * wpb) {
    const float mu = mean[row], rs = rstd[row];
    const float* xr = x + row * D;
    const float* dr = dout + row * D;
    float s1 = 0.f, s2 = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float xh = (xr[d] - mu) * rs;
      const float z = xh * gamma[d] + beta[d];
      float g = dr[d];
      if (drop_thresh) g = dropout_keep((uint64_t)(row * D + d), seed, drop_thresh) ? g * drop_scale : 0.f;
      const float dz = g * act_grad(z, act);
      if (per_warp) {
        sg[d] += dz * xh;
        sb[d] += dz;
      } else {
        atomicAdd(&sg[d], dz * xh);
        atomicAdd(&sb[d], dz);
      }
      const float dzg = dz * gamma[d];
      s1 += dzg;
      s2 += dzg * xh;
    }
    s1 = warp_sum(s1) / (float)D;
    s2 = warp_sum(s2) / (float)D;
    for (int d = lane; d < D; d += 32) {
      const float xh = (xr[d] - mu) * rs;
      const float z = xh * gamma[d] + beta[d];
      float g = dr[d];
      if (drop_thresh) g = dropout_keep((uint64_t)(row * D + d), seed, drop_thresh) ? g * drop_scale : 0.f;
      const float dzg = g * act_grad(z, act) * gamma[d];
      dx[row * D + d] = rs * (dzg - s1 - xh * s2);
    }
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < (per_warp ? wpb : 1); ++w) {
      a += sm[w * 2 * D + d];
      b += sm[w * 2 * D + D + d];
    }
    dgamma_part[(long long)blockIdx.x * D + d] = a;
    dbeta_part[(long long)blockIdx.x * D + d] = b;
  }
}

// ------------------------------------------------------------------ activation + dropout
__global__ void act_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, long long n, int act,
                               float drop_scale, uint32_t drop_thresh, uint64_t seed) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float o = apply_act(x[i], act);
    if (drop_thresh) o = dropout_keep((uint64_t)i, seed, drop_thresh) ? o * drop_scale : 0.f;
    out[i] = o;
  }
}
__global__ void act_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ x, float* __restrict__ dx,
                               long long n, int act, float drop_scale, uint32_t drop_thresh, uint64_t seed) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float g = dout[i];
    if (drop_thresh) g = dropout_keep((uint64_t)i, seed, drop_thresh) ? g * drop_scale : 0.f;
    dx[i] = g * act_grad(x[i], act);
  }
}

// All-gather over NVLink peer mappings: block row r copies rank r's shard (read through its mapped pointer,
// 128-bit loads) into slot r of the local buffer.  One launch, no NCCL.
struct PeerPtrs {
  const float* p[8];
};
__global__ void peer_gather_kernel(const PeerPtrs pp, long long n4, float* __restrict__ dst) {
  const float4* src = reinterpret_cast<const float4*>(pp.p[blockIdx.y]);
  float4* out = reinterpret_cast<float4*>(dst) + (long long)blockIdx.y * n4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
    out[i] = src[i];
}

__global__ void round_tf32_kernel(const float* __restrict__ x, float* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = round_tf32(x[i]);
}

// 3-way tf32 split of a (rows, cols) matrix, concatenated along `axis`:
//   which == 0: [hi | lo | hi]     which == 1: [hi | hi | lo]     hi = tf32(x), lo = tf32(x - hi)
// Contracting a which-0 operand with a which-1 operand over the tripled axis gives hi*hi + lo*hi + hi*lo,
// an fp32-accurate product on the tf32 tensor cores (used for the small, error-sensitive projections).
__global__ void split3_kernel(const float* __restrict__ x, float* __restrict__ out, long long rows, long long cols,
                              int which, int axis) {
  const long long n = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    const float hi = round_tf32(v);
    const float lo = round_tf32(v - hi);
    const float p1 = which == 0 ? lo : hi, p2 = which == 0 ? hi : lo;
    if (axis == 1) {
      const long long r = i / cols, c = i - r * cols;
      float* o = out + r * 3 * cols + c;
      o[0] = hi; o[cols] = p1; o[2 * cols] = p2;
    } else {
      out[i] = hi; out[n + i] = p1; out[2 * n + i] = p2;
    }
  }
}

// float4 variants (n % 4 == 0, 16-B aligned)
__global__ void act_fwd_v4_kernel(const float* __restrict__ x, float* __restrict__ out, long long n4, int act,
                                  float drop_scale, uint32_t drop_thresh, uint64_t seed, int round_out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    float o[4] = {v.x, v.y, v.z, v.w};
    if (act == XM_ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = gelu_erf(o[j]);
    } else if (act != XM_ACT_NONE) {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = apply_act(o[j], act);
    }
    if (drop_thresh) {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = dropout_keep((uint64_t)(4 * i + j), seed, drop_thresh) ? o[j] * drop_scale : 0.f;
    }
    if (round_out) {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = round_tf32(o[j]);
    }
    reinterpret_cast<float4*>(out)[i] = make_float4(o[0], o[1], o[2], o[3]);
  }
}
__global__ void act_bwd_v4_kernel(const float* __restrict__ dout, const float* __restrict__ x, float* __restrict__ dx,
                                  long long n4, int act, float drop_scale, uint32_t drop_thresh, uint64_t seed,
                                  int round_out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 gv = reinterpret_cast<const float4*>(dout)[i];
    const float4 xv = reinterpret_cast<const float4*>(x)[i];
    float g[4] = {gv.x, gv.y, gv.z, gv.w};
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
    if (drop_thresh) {
#pragma unroll
      for (int j = 0; j < 4; ++j) g[j] = dropout_keep((uint64_t)(4 * i + j), seed, drop_thresh) ? g[j] * drop_scale : 0.f;
    }
    if (act == XM_ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 4; ++j) g[j] *= gelu_erf_grad(xs[j]);
    } else if (act != XM_ACT_NONE) {
#pragma unroll
      for (int j = 0; j < 4; ++j) g[j] *= act_grad(xs[j], act);
    }
    if (round_out) {
#pragma unroll
      for (int j = 0; j < 4; ++j) g[j] = round_tf32(g[j]);
    }
    reinterpret_cast<float4*>(dx)[i] = make_float4(g[0], g[1], g[2], g[3]);
  }
}

// act backward over (M, C) rows that also accumulates the column sums of dx (the bias gradient of the Linear
// whose output was activated): thread tx owns 4 columns, rows are strided over blockIdx / ty, per-block
// partial sums -> (gridDim.x, C).  Saves a separate pass over dx.
__global__ void act_bwd_colsum_v4_kernel(const float* __restrict__ dout, const float* __restrict__ x, float* __restrict__ dx,
                                         long long M, int C, int act, float drop_scale, uint32_t drop_thresh, uint64_t seed,
                                         int round_out, float* __restrict__ colsum_part) {
  extern __shared__ float smf[];  // [RY][C]
  const int tx = threadIdx.x, ty = threadIdx.y, RY = blockDim.y, c0 = 4 * tx;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long r = blockIdx.x * (long long)RY + ty; r < M; r += (long long)gridDim.x * RY) {
    const long long i = r * C + c0;
    const float4 gv = *reinterpret_cast<const float4*>(dout + i);
    const float4 xv = *reinterpret_cast<const float4*>(x + i);
    float g[4] = {gv.x, gv.y, gv.z, gv.w};
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
    if (drop_thresh) {
#pragma unroll
      for (int j = 0; j < 4; ++j) g[j] = dropout_keep((uint64_t)(i + j), seed, drop_thresh) ? g[j] * drop_scale : 0.f;
    }
    if (act == XM_ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 4; ++j) g[j] *= gelu_erf_grad(xs[j]);
    } else if (act != XM_ACT_NONE) {
#pragma unroll
      for (int j = 0; j < 4; ++j) g[j] *= act_grad(xs[j], act);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[j] += g[j];
      if (round_out) g[j] = round_tf32(g[j]);
    }
    *reinterpret_cast<float4*>(dx + i) = make_float4(g[0], g[1], g[2], g[3]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) smf[ty * C + c0 + j] = acc[j];
  __syncthreads();
  for (int c = ty * blockDim.x + tx; c < C; c += blockDim.x * RY) {
    float t = 0.f;
    for (int w = 0; w < RY; ++w) t += smf[w * C + c];
    colsum_part[(long long)blockIdx.x * C + c] = t;
  }
}

// ------------------------------------------------------------------ small reductions
// out[n] = sum_m x[m, n] in two deterministic stages: block (bx, by) sums rows [by*chunk, (by+1)*chunk)
// of columns [32 bx, 32 bx + 32) into part[by][n] (or straight into out when gridDim.y == 1); a second
// launch of the same kernel reduces the gridDim.y partial rows.  block = 32 columns x 8 row lanes.
__global__ void colsum_kernel(const float* __restrict__ x, long long M, long long N, long long ld, long long chunk,
                              float* __restrict__ out) {
  __shared__ float sm[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long n = blockIdx.x * 32ll + tx;
  const long long m0 = blockIdx.y * chunk;
  const long long m1 = min(M, m0 + chunk);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (n < N) {
    long long m = m0 + ty;
    for (; m + 24 < m1; m += 32) {  // 4 independent loads in flight per thread
      a0 += x[m * ld + n];
      a1 += x[(m + 8) * ld + n];
      a2 += x[(m + 16) * ld + n];
      a3 += x[(m + 24) * ld + n];
    }
    for (; m < m1; m += 8) a0 += x[m * ld + n];
  }
  sm[ty][tx] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (ty == 0 && n < N) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += sm[i][tx];
    out[blockIdx.y * N + n] = s;
  }
}

// float4 variant (N % 4 == 0, 16-B aligned rows): block = 32 column quads x 8 row lanes = 128 columns
__global__ void colsum_v4_kernel(const float* __restrict__ x, long long M, long long N, long long ld, long long chunk,
                                 float* __restrict__ out) {
  __shared__ float4 sm[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long n = blockIdx.x * 128ll + 4 * tx;
  const long long m0 = blockIdx.y * chunk;
  const long long m1 = min(M, m0 + chunk);
  float4 acc[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n < N) {
    const float* base = x + n;
    long long m = m0 + ty;
    for (; m + 24 < m1; m += 32) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 v = *reinterpret_cast<const float4*>(base + (m + 8 * u) * ld);
        acc[u].x += v.x; acc[u].y += v.y; acc[u].z += v.z; acc[u].w += v.w;
      }
    }
    for (; m < m1; m += 8) {
      const float4 v = *reinterpret_cast<const float4*>(base + m * ld);
      acc[0].x += v.x; acc[0].y += v.y; acc[0].z += v.z; acc[0].w += v.w;
    }
  }
  sm[ty][tx] = make_float4((acc[0].x + acc[1].x) + (acc[2].x + acc[3].x), (acc[0].y + acc[1].y) + (acc[2].y + acc[3].y),
                           (acc[0].z + acc[1].z) + (acc[2].z + acc[3].z), (acc[0].w + acc[1].w) + (acc[2].w + acc[3].w));
  __syncthreads();
  if (ty == 0 && n < N) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 v = sm[i][tx];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    *reinterpret_cast<float4*>(out + blockIdx.y * N + n) = t;
  }
}

// ------------------------------------------------------------------ L2 row normalisation (warp per row)
__global__ void l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ xn, float* __restrict__ inv_norm,
                                  long long M, int D, float eps) {
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  float q = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float v = x[row * D + d];
    q += v * v;
  }
  const float inv = 1.0f / fmaxf(sqrtf(warp_sum(q)), eps);
  if (lane == 0) inv_norm[row] = inv;
  for (int d = lane; d < D; d += 32) xn[row * D + d] = round_tf32(x[row * D + d] * inv);
}
// xn = x / max(||x||, eps) in full fp32 plus the 3-way tf32 split xs (M, 3D) of xn:
//   which == 0: [hi | lo | hi]      which == 1: [hi | hi | lo]      hi = tf32(xn), lo = tf32(xn - hi)
// so that  xs_a . xs_b = hi_a hi_b + lo_a hi_b + hi_a lo_b  ~ fp32-accurate dot product on the tf32 tensor cores
// (the similarity is divided by tau = 0.07 before the softmax: a plain tf32 product would put ~4e-3 of
// relative noise into every softmax probability and hence into every contrastive gradient).
__global__ void l2norm_split_fwd_kernel(const float* __restrict__ x, float* __restrict__ xn, float* __restrict__ xs,
                                        float* __restrict__ inv_norm, long long M, int D, float eps, int which) {
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  float q = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float v = x[row * D + d];
    q += v * v;
  }
  const float inv = 1.0f / fmaxf(sqrtf(warp_sum(q)), eps);
  if (lane == 0) inv_norm[row] = inv;
  for (int d = lane; d < D; d += 32) {
    const float v = x[row * D + d] * inv;
    const float hi = round_tf32(v);
    const float lo = round_tf32(v - hi);
    xn[row * D + d] = v;
    float* o = xs + row * 3ll * D;
    o[d] = hi;
    o[D + d] = which == 0 ? lo : hi;
    o[2 * D + d] = which == 0 ? hi : lo;
  }
}
__global__ void l2norm_bwd_kernel(const float* __restrict__ dxn, const float* __restrict__ xn,
                                  const float* __restrict__ inv_norm, float* __restrict__ dx, long long M, int D) {
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  float dot = 0.f;
  for (int d = lane; d < D; d += 32) dot += dxn[row * D + d] * xn[row * D + d];
  dot = warp_sum(dot);
  const float inv = inv_norm[row];
  for (int d = lane; d < D; d += 32) dx[row * D + d] = (dxn[row * D + d] - xn[row * D + d] * dot) * inv;
}

static BnActArgs make_bn_args(const float* y, const float* mean, const float* invstd, const float* gamma,
                              const float* beta, int64_t B, int64_t T, int64_t C, int64_t ldy, int64_t ldo, int act,
                              int pool, float drop_p, uint64_t seed, int drop_before_pool, int round_out) {
  BnActArgs a;
  a.y = y; a.mean = mean; a.invstd = invstd; a.gamma = gamma; a.beta = beta;
  a.B = B; a.T = T; a.ldy = ldy; a.ldo = ldo; a.C = (int)C; a.act = act; a.pool = pool;
  a.drop_before_pool = drop_before_pool; a.round_out = round_out; a.seed = seed;
  if (drop_p > 0.f) {
    a.drop_scale = 1.0f / (1.0f - drop_p);
    double th = (double)drop_p * 4294967296.0;
    a.drop_thresh = th >= 4294967295.0 ? 4294967295u : (uint32_t)th;
    if (a.drop_thresh == 0u) a.drop_thresh = 1u;
  } else {
    a.drop_scale = 1.0f;
    a.drop_thresh = 0u;
  }
  return a;
}

static bool bn_args_ok(const void* y, const void* m, const void* is, const void* g, const void* b, int64_t B, int64_t T,
                       int64_t C, int64_t ldy, int pool, float p) {
  if (!y || !m || !is || !g || !b || B <= 0 || C <= 0 || T <= 0 || ldy < C) return false;
  if (B * T >= 2147483647ll) return false;  // 32-bit row arithmetic in the pooled paths
  if (pool != 0 && pool != 2) return false;
  if (pool == 2 && T < 2) return false;
  if (!(p >= 0.f && p < 1.f)) return false;
  return true;
}

// ------------------------------------------------------------------ vectorised BatchNorm path (C % 4 == 0)
// Thread (tx, ty): tx owns 4 consecutive channels (one float4) of every row it visits, ty strides over
// rows, so the per-channel constants live in registers and there is no index division per element.
// A warp covers whole 16-B aligned row segments: 128-bit coalesced loads/stores.  Rows are unrolled
// x4 for memory-level parallelism (HBM latency x bandwidth needs ~32 KB in flight per SM).
struct F4 {
  float v[4];
};
XM_DEVICE F4 ld4(const float* p) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  return F4{{t.x, t.y, t.z, t.w}};
}
XM_DEVICE void st4(float* p, const F4& a) { *reinterpret_cast<float4*>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]); }

template <int ACT>
XM_DEVICE float act_c(float x, int act) {
  if (ACT == XM_ACT_NONE) return x;
  if (ACT == XM_ACT_RELU) return fmaxf(x, 0.f);
  if (ACT == XM_ACT_GELU) return gelu_erf(x);
  return apply_act(x, act);
}
template <int ACT>
XM_DEVICE float actg_c(float x, int act) {
  if (ACT == XM_ACT_NONE) return 1.f;
  if (ACT == XM_ACT_RELU) return x > 0.f ? 1.f : 0.f;
  if (ACT == XM_ACT_GELU) return gelu_erf_grad(x);
  return act_grad(x, act);
}

// grid (nsplit), block (C/4, RY)
__global__ void bn_partial_stats_v4_kernel(const float* __restrict__ y, long long R, int C, long long ld,
                                           long long rows_per_split, double* __restrict__ partials) {
  extern __shared__ double smd[];  // [RY][C][2]
  const int tx = threadIdx.x, ty = threadIdx.y, RY = blockDim.y;
  const long long r0 = blockIdx.x * rows_per_split;
  const long long r1 = min(R, r0 + rows_per_split);
  double sum[4] = {0, 0, 0, 0}, sq[4] = {0, 0, 0, 0};
  float ps[4] = {0, 0, 0, 0}, pq[4] = {0, 0, 0, 0};
  int n = 0;
  const float* base = y + 4 * tx;
  long long r = r0 + ty;
  for (; r + 3ll * RY < r1; r += 4ll * RY) {
    F4 a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = ld4(base + (r + (long long)u * RY) * ld);
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        ps[j] += a[u].v[j];
        pq[j] += a[u].v[j] * a[u].v[j];
      }
    if (++n == 16) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sum[j] += (double)ps[j]; sq[j] += (double)pq[j];
        ps[j] = 0.f; pq[j] = 0.f;
      }
      n = 0;
    }
  }
  for (; r < r1; r += RY) {
    const F4 a = ld4(base + r * ld);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      ps[j] += a.v[j];
      pq[j] += a.v[j] * a.v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sum[j] += (double)ps[j];
    sq[j] += (double)pq[j];
    smd[((long long)ty * C + 4 * tx + j) * 2 + 0] = sum[j];
    smd[((long long)ty * C + 4 * tx + j) * 2 + 1] = sq[j];
  }
  __syncthreads();
  const int tid = ty * blockDim.x + tx;
  for (int c = tid; c < C; c += blockDim.x * RY) {
    double s = 0.0, q = 0.0;
    for (int i = 0; i < RY; ++i) {
      s += smd[((long long)i * C + c) * 2 + 0];
      q += smd[((long long)i * C + c) * 2 + 1];
    }
    partials[((long long)blockIdx.x * C + c) * 2 + 0] = s;
    partials[((long long)blockIdx.x * C + c) * 2 + 1] = q;
  }
}

struct ChanConst {
  float sc[4], sh[4], mu[4], is[4];
};
XM_DEVICE ChanConst chan_consts(const BnActArgs& a, int c0) {
  ChanConst k;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    k.mu[j] = a.mean[c0 + j];
    k.is[j] = a.invstd[c0 + j];
    k.sc[j] = a.gamma[c0 + j] * k.is[j];
    k.sh[j] = a.beta[c0 + j] - k.mu[j] * k.sc[j];
  }
  return k;
}

// forward: grid-stride over OUTPUT rows; block (C/4, RY)
template <int ACT, int POOL>
__global__ void bn_act_fwd_v4_kernel(const BnActArgs a, long long R_out, float* __restrict__ out) {
  const int tx = threadIdx.x, c0 = 4 * tx;
  const ChanConst k = chan_consts(a, c0);
  const long long To = a.T / 2;
  const long long stride = (long long)gridDim.x * blockDim.y;
  for (long long ro = blockIdx.x * (long long)blockDim.y + threadIdx.y; ro < R_out; ro += stride) {
    F4 o;
    if (POOL == 2) {
      const long long b = (long long)((unsigned)ro / (unsigned)To), tp = ro - b * To;  // rows < 2^31 (host-checked)
      const long long r0 = b * a.T + 2 * tp;
      const F4 x0 = ld4(a.y + r0 * a.ldy + c0), x1 = ld4(a.y + (r0 + 1) * a.ldy + c0);
      Mul4 m0, m1;  // m0: row r0 (drop before pool) or the pooled row (drop after pool); m1: row r0 + 1
      if (a.drop_thresh) {
        m0 = drop_mul4(a, (a.drop_before_pool ? r0 : ro) * a.C + c0);
        if (a.drop_before_pool) m1 = drop_mul4(a, (r0 + 1) * a.C + c0);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a0 = act_c<ACT>(x0.v[j] * k.sc[j] + k.sh[j], a.act);
        float a1 = act_c<ACT>(x1.v[j] * k.sc[j] + k.sh[j], a.act);
        if (a.drop_thresh) {
          if (a.drop_before_pool) {
            a0 *= m0.v[j];
            a1 *= m1.v[j];
            o.v[j] = fmaxf(a0, a1);
          } else {
            o.v[j] = fmaxf(a0, a1) * m0.v[j];
          }
        } else {
          o.v[j] = fmaxf(a0, a1);
        }
      }
    } else {
      const F4 x0 = ld4(a.y + ro * a.ldy + c0);
      Mul4 m0;
      if (a.drop_thresh) m0 = drop_mul4(a, ro * a.C + c0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = act_c<ACT>(x0.v[j] * k.sc[j] + k.sh[j], a.act);
        if (a.drop_thresh) v *= m0.v[j];
        o.v[j] = v;
      }
    }
    if (a.round_out == 2) {  // tf32 split along the channel axis [hi | lo]: operand of a 3-pass conv
      F4 lo;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float hi = round_tf32(o.v[j]);
        lo.v[j] = round_tf32(o.v[j] - hi);
        o.v[j] = hi;
      }
      st4(out + ro * a.ldo + c0, o);
      st4(out + ro * a.ldo + a.C + c0, lo);
      continue;
    }
    if (a.round_out) {
#pragma unroll
      for (int j = 0; j < 4; ++j) o.v[j] = round_tf32(o.v[j]);
    }
    st4(out + ro * a.ldo + c0, o);
  }
}

// dz of the (up to) two input rows behind output row ro, 4 channels at once.  Loads and arithmetic are separate steps so
// that a caller can issue the loads of several rows before the first (branchy) computation: with load + compute per row the
// compiler kept one row's two 128-bit loads in flight per thread, and the backward-reduce kernel ran at 3.3 TB/s.
struct BnRow {
  F4 g, x0, x1;
  long long r0;
};
template <int POOL>
XM_DEVICE BnRow bn_load_row(const BnActArgs& a, const float* __restrict__ dout, long long ro, int c0) {
  BnRow w;
  w.g = ld4(dout + ro * a.ldo + c0);
  if (POOL == 2) {
    const long long To = a.T / 2, b = (long long)((unsigned)ro / (unsigned)To), tp = ro - b * To;
    w.r0 = b * a.T + 2 * tp;
    w.x0 = ld4(a.y + w.r0 * a.ldy + c0);
    w.x1 = ld4(a.y + (w.r0 + 1) * a.ldy + c0);
  } else {
    w.r0 = ro;
    w.x0 = ld4(a.y + ro * a.ldy + c0);
#pragma unroll
    for (int j = 0; j < 4; ++j) w.x1.v[j] = 0.f;
  }
  return w;
}
template <int ACT, int POOL>
XM_DEVICE void bn_dz4_of(const BnActArgs& a, const ChanConst& k, const BnRow& w, long long ro, int c0, F4& dz0, F4& dz1) {
  const F4& g = w.g;
  const F4& x0 = w.x0;
  const F4& x1 = w.x1;
  const long long r0 = w.r0;
  if (POOL == 2) {
    Mul4 d0, d1;
    if (a.drop_thresh) {
      d0 = drop_mul4(a, (a.drop_before_pool ? r0 : ro) * a.C + c0);
      if (a.drop_before_pool) d1 = drop_mul4(a, (r0 + 1) * a.C + c0);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float z0 = x0.v[j] * k.sc[j] + k.sh[j], z1 = x1.v[j] * k.sc[j] + k.sh[j];
      float a0 = act_c<ACT>(z0, a.act), a1 = act_c<ACT>(z1, a.act);
      float m0 = 1.f, m1 = 1.f, gg = g.v[j];
      if (a.drop_thresh) {
        if (a.drop_before_pool) {
          m0 = d0.v[j];
          m1 = d1.v[j];
          a0 *= m0;
          a1 *= m1;
        } else {
          gg *= d0.v[j];
        }
      }
      const bool first = a0 >= a1;  // ties -> first element, as torch max_pool1d
      dz0.v[j] = first ? gg * m0 * actg_c<ACT>(z0, a.act) : 0.f;
      dz1.v[j] = first ? 0.f : gg * m1 * actg_c<ACT>(z1, a.act);
    }
  } else {
    Mul4 d0;
    if (a.drop_thresh) d0 = drop_mul4(a, ro * a.C + c0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float gg = g.v[j];
      if (a.drop_thresh) gg *= d0.v[j];
      dz0.v[j] = gg * actg_c<ACT>(x0.v[j] * k.sc[j] + k.sh[j], a.act);
      dz1.v[j] = 0.f;
    }
  }
}
template <int ACT, int POOL>
XM_DEVICE void bn_dz4(const BnActArgs& a, const ChanConst& k, const float* __restrict__ dout, long long ro, int c0,
                      long long& r0, F4& x0, F4& x1, F4& dz0, F4& dz1) {
  const BnRow w = bn_load_row<POOL>(a, dout, ro, c0);
  bn_dz4_of<ACT, POOL>(a, k, w, ro, c0, dz0, dz1);
  r0 = w.r0;
  x0 = w.x0;
  x1 = w.x1;
}

// grid (nsplit), block (C/4, RY) over OUTPUT rows: partial sums of dz and dz*xhat
template <int ACT, int POOL>
__global__ void __launch_bounds__(256, POOL == 2 ? 2 : 3) bn_act_bwd_reduce_v4_kernel(const BnActArgs a, const float* __restrict__ dout, long long R_out,
                                            long long rows_per_split, double* __restrict__ partials) {
  extern __shared__ double smd[];  // [RY][C][2]
  const int tx = threadIdx.x, ty = threadIdx.y, RY = blockDim.y, c0 = 4 * tx;
  const ChanConst k = chan_consts(a, c0);
  const long long r_lo = blockIdx.x * rows_per_split;
  const long long r_hi = min(R_out, r_lo + rows_per_split);
  double s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
  float p0[4] = {0, 0, 0, 0}, p1[4] = {0, 0, 0, 0};
  int n = 0;
  long long ro = r_lo + ty;
  for (; ro + RY < r_hi; ro += 2ll * RY) {  // two output rows per iteration: twice the loads in flight
    const BnRow wa = bn_load_row<POOL>(a, dout, ro, c0), wb = bn_load_row<POOL>(a, dout, ro + RY, c0);  // all loads first
    F4 da0, da1, db0, db1;
    bn_dz4_of<ACT, POOL>(a, k, wa, ro, c0, da0, da1);
    bn_dz4_of<ACT, POOL>(a, k, wb, ro + RY, c0, db0, db1);
    const F4 &xa0 = wa.x0, &xa1 = wa.x1, &xb0 = wb.x0, &xb1 = wb.x1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      p0[j] += (da0.v[j] + da1.v[j]) + (db0.v[j] + db1.v[j]);
      p1[j] += da0.v[j] * ((xa0.v[j] - k.mu[j]) * k.is[j]) + da1.v[j] * ((xa1.v[j] - k.mu[j]) * k.is[j]) +
               db0.v[j] * ((xb0.v[j] - k.mu[j]) * k.is[j]) + db1.v[j] * ((xb1.v[j] - k.mu[j]) * k.is[j]);
    }
    if (++n == 32) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s0[j] += (double)p0[j]; s1[j] += (double)p1[j];
        p0[j] = 0.f; p1[j] = 0.f;
      }
      n = 0;
    }
  }
  for (; ro < r_hi; ro += RY) {
    long long r0;
    F4 x0, x1, dz0, dz1;
    bn_dz4<ACT, POOL>(a, k, dout, ro, c0, r0, x0, x1, dz0, dz1);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      p0[j] += dz0.v[j] + dz1.v[j];
      p1[j] += dz0.v[j] * ((x0.v[j] - k.mu[j]) * k.is[j]) + dz1.v[j] * ((x1.v[j] - k.mu[j]) * k.is[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    smd[((long long)ty * a.C + c0 + j) * 2 + 0] = s0[j] + (double)p0[j];
    smd[((long long)ty * a.C + c0 + j) * 2 + 1] = s1[j] + (double)p1[j];
  }
  __syncthreads();
  const int tid = ty * blockDim.x + tx;
  for (int c = tid; c < a.C; c += blockDim.x * RY) {
    double s = 0.0, q = 0.0;
    for (int i = 0; i < RY; ++i) {
      s += smd[((long long)i * a.C + c) * 2 + 0];
      q += smd[((long long)i * a.C + c) * 2 + 1];
    }
    partials[((long long)blockIdx.x * a.C + c) * 2 + 0] = s;
    partials[((long long)blockIdx.x * a.C + c) * 2 + 1] = q;
  }
}

template <int ACT, int POOL>
__global__ void bn_act_bwd_apply_v4_kernel(const BnActArgs a, const float* __restrict__ dout,
                                           const float* __restrict__ dbeta, const float* __restrict__ dgamma, float inv_n,
                                           long long R_out, float* __restrict__ dy) {
  const int tx = threadIdx.x, c0 = 4 * tx;
  const ChanConst k = chan_consts(a, c0);
  float k1[4], k2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    k1[j] = dbeta[c0 + j] * inv_n;
    k2[j] = dgamma[c0 + j] * inv_n;
  }
  const long long To = a.T / 2;
  const long long stride = (long long)gridDim.x * blockDim.y;
  for (long long ro = blockIdx.x * (long long)blockDim.y + threadIdx.y; ro < R_out; ro += stride) {
    long long r0;
    F4 x0, x1, dz0, dz1, o0, o1;
    bn_dz4<ACT, POOL>(a, k, dout, ro, c0, r0, x0, x1, dz0, dz1);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      o0.v[j] = k.sc[j] * (dz0.v[j] - k1[j] - (x0.v[j] - k.mu[j]) * k.is[j] * k2[j]);
      o1.v[j] = k.sc[j] * (dz1.v[j] - k1[j] - (x1.v[j] - k.mu[j]) * k.is[j] * k2[j]);
      if (a.round_out) { o0.v[j] = round_tf32(o0.v[j]); o1.v[j] = round_tf32(o1.v[j]); }
    }
    st4(dy + r0 * a.ldy + c0, o0);
    if (POOL == 2) {
      st4(dy + (r0 + 1) * a.ldy + c0, o1);
      const long long b = (long long)((unsigned)ro / (unsigned)To), tp = ro - b * To;  // rows < 2^31 (host-checked)
      if ((a.T & 1) && tp == To - 1) {  // odd tail row is dropped by the pool: dz = 0
        const F4 xt = ld4(a.y + (r0 + 2) * a.ldy + c0);
        F4 ot;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ot.v[j] = k.sc[j] * (0.f - k1[j] - (xt.v[j] - k.mu[j]) * k.is[j] * k2[j]);
          if (a.round_out) ot.v[j] = round_tf32(ot.v[j]);
        }
        st4(dy + (r0 + 2) * a.ldy + c0, ot);
      }
    }
  }
}

static bool v4_ok(const void* p0, const void* p1, const void* p2, int64_t C, int64_t ld0, int64_t ld1,
                  int64_t rows = 0) {
  if (rows >= 2147483647ll) return false;  // the float4 kernels use 32-bit row arithmetic
  auto al = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return (C % 4 == 0) && C <= 1024 && (ld0 % 4 == 0) && (ld1 % 4 == 0) && al(p0) && al(p1) && al(p2);
}
static dim3 v4_block(int64_t C) {
  const int tx = (int)(C / 4);
  int ry = 256 / tx;
  if (ry < 1) ry = 1;
  return dim3(tx, ry);
}
static int v4_grid(long long rows, int ry) {
  long long b = (rows + ry - 1) / ry;
  const long long cap = (long long)kNumSMs * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

#define XM_BN_DISPATCH(KERNEL, act, pool, ...)                                                              \
  do {                                                                                                      \
    if (pool == 2) {                                                                                        \
      if (act == XM_ACT_GELU) KERNEL<XM_ACT_GELU, 2> __VA_ARGS__;                                           \
      else if (act == XM_ACT_RELU) KERNEL<XM_ACT_RELU, 2> __VA_ARGS__;                                      \
      else if (act == XM_ACT_NONE) KERNEL<XM_ACT_NONE, 2> __VA_ARGS__;                                      \
      else KERNEL<99, 2> __VA_ARGS__;                                                                       \
    } else {                                                                                                \
      if (act == XM_ACT_GELU) KERNEL<XM_ACT_GELU, 0> __VA_ARGS__;                                           \
      else if (act == XM_ACT_RELU) KERNEL<XM_ACT_RELU, 0> __VA_ARGS__;                                      \
      else if (act == XM_ACT_NONE) KERNEL<XM_ACT_NONE, 0> __VA_ARGS__;                                      \
      else KERNEL<99, 0> __VA_ARGS__;                                                                       \
    }                                                                                                       \
  } while (0)

static int ew_grid(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = (long long)kNumSMs * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace xm

XM_DEFINE_SEED_EPOCH_SLOT(elementwise)

using namespace xm;

extern "C" {

int xm_roi_meanstd_f32(const float* x, int64_t B, int64_t TR, int64_t ROI, float* out, void* stream) {
  if (!x || !out || B <= 0 || TR <= 0 || ROI <= 0 || B > 65535) return XM_ERR_INVALID;
  if (ROI % 4 == 0 && ROI / 4 <= kRoiThreads && TR < (1 << 30) &&
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    const int roi4 = (int)(ROI / 4);
    int G = kRoiThreads / roi4;
    if (G > TR) G = (int)TR;
    roi_meanstd_v4_kernel<<<(unsigned)B, kRoiThreads, (size_t)G * roi4 * sizeof(float4), (cudaStream_t)stream>>>(x, (int)TR, roi4, G,
                                                                                                                 out);
    return check_launch();
  }
  dim3 grid(ceil_div(ROI, 128), (unsigned)B);
  roi_meanstd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(x, TR, ROI, out);
  return check_launch();
}

int xm_zscore_f32(const float* x, int64_t n_items, int64_t item_len, float eps, float* out, void* stream) {
  if (!x || !out || n_items <= 0 || item_len <= 0) return XM_ERR_INVALID;
  zscore_kernel<<<grid_rows(n_items), 256, 0, (cudaStream_t)stream>>>(x, item_len, eps, out);
  return check_launch();
}

int xm_bn_nsplit(int64_t R, int64_t C) {
  const int64_t cblk = (C + 31) / 32;
  int64_t ns = (kNumSMs * 16 + cblk - 1) / cblk;  // the float4 kernels use ONE block per split (all channels)
  const int64_t max_by_rows = (R + 63) / 64;  // at least 64 rows per split
  if (ns > max_by_rows) ns = max_by_rows;
  if (ns < 1) ns = 1;
  if (ns > 65535) ns = 65535;
  return (int)ns;
}

int xm_bn_partial_stats_f32(const float* y, int64_t R, int64_t C, int64_t ldy, double* partials, void* stream) {
  if (!y || !partials || R <= 0 || C <= 0 || ldy < C) return XM_ERR_INVALID;
  const int ns = xm_bn_nsplit(R, C);
  const long long rps = (R + ns - 1) / ns;
  if (v4_ok(y, nullptr, nullptr, C, ldy, 4)) {
    const dim3 blk = v4_block(C);
    bn_partial_stats_v4_kernel<<<ns, blk, blk.y * C * 2 * sizeof(double), (cudaStream_t)stream>>>(y, R, (int)C, ldy, rps,
                                                                                                partials);
    return check_launch();
  }
  dim3 grid(ceil_div(C, 32), (unsigned)ns);
  bn_partial_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(y, R, (int)C, ldy, rps, partials);
  return check_launch();
}

int xm_bn_finalize_stats(const double* partials, int nsplit, int64_t C, double total_count, float eps, float* mean,
                         float* invstd, float* running_mean, float* running_var, float momentum, void* stream) {
  if (!partials || !mean || !invstd || nsplit <= 0 || C <= 0 || total_count <= 0) return XM_ERR_INVALID;
  bn_finalize_stats_kernel<<<ceil_div(C, 8), 256, 0, (cudaStream_t)stream>>>(partials, nsplit, (int)C, total_count, eps,
                                                                               mean, invstd, running_mean, running_var,
                                                                               momentum);
  return check_launch();
}

int xm_bn_act_fwd_f32(const float* y, const float* mean, const float* invstd, const float* gamma, const float* beta,
                      float* out, int64_t B, int64_t T, int64_t C, int64_t ldy, int64_t ldo, int act, int pool,
                      float drop_p, uint64_t seed, int drop_before_pool, int round_out, void* stream) {
  if (!bn_args_ok(y, mean, invstd, gamma, beta, B, T, C, ldy, pool, drop_p) || !out || ldo < (round_out == 2 ? 2 * C : C))
    return XM_ERR_INVALID;
  BnActArgs a = make_bn_args(y, mean, invstd, gamma, beta, B, T, C, ldy, ldo, act, pool, drop_p, seed, drop_before_pool,
                             round_out);
  const long long R_out = B * (pool == 2 ? T / 2 : T);
  if (v4_ok(y, out, nullptr, C, ldy, ldo, B * T)) {
    const dim3 blk = v4_block(C);
    XM_BN_DISPATCH(bn_act_fwd_v4_kernel, act, pool, <<<v4_grid(R_out, blk.y), blk, 0, (cudaStream_t)stream>>>(a, R_out, out));
    return check_launch();
  }
  bn_act_fwd_kernel<<<ew_grid(R_out * C), 256, 0, (cudaStream_t)stream>>>(a, R_out, out);
  return check_launch();
}

int xm_bn_act_bwd_reduce_f32(const float* dout, const float* y, const float* mean, const float* invstd,
                             const float* gamma, const float* beta, int64_t B, int64_t T, int64_t C, int64_t ldy,
                             int64_t ldo, int act, int pool, float drop_p, uint64_t seed, int drop_before_pool,
                             double* partials, void* stream) {
  if (!bn_args_ok(y, mean, invstd, gamma, beta, B, T, C, ldy, pool, drop_p) || !dout || !partials) return XM_ERR_INVALID;
  BnActArgs a = make_bn_args(y, mean, invstd, gamma, beta, B, T, C, ldy, ldo, act, pool, drop_p, seed, drop_before_pool, 0);
  const long long R_out = B * (pool == 2 ? T / 2 : T);
  const int ns = xm_bn_nsplit(B * T, C);  // same split count as the forward statistics
  const long long rps = (R_out + ns - 1) / ns;
  if (v4_ok(y, dout, nullptr, C, ldy, ldo, B * T)) {
    const dim3 blk = v4_block(C);
    XM_BN_DISPATCH(bn_act_bwd_reduce_v4_kernel, act, pool,
                   <<<ns, blk, blk.y * C * 2 * sizeof(double), (cudaStream_t)stream>>>(a, dout, R_out, rps, partials));
    return check_launch();
  }
  dim3 grid(ceil_div(C, 32), (unsigned)ns);
  bn_act_bwd_reduce_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, dout, R_out, rps, partials);
  return check_launch();
}

int xm_bn_bwd_finalize(const double* partials, int nsplit, int64_t C, float* dbeta, float* dgamma, void* stream) {
  if (!partials || !dbeta || !dgamma || nsplit <= 0 || C <= 0) return XM_ERR_INVALID;
  bn_bwd_finalize_kernel<<<ceil_div(C, 8), 256, 0, (cudaStream_t)stream>>>(partials, nsplit, (int)C, dbeta, dgamma);
  return check_launch();
}

int xm_bn_act_bwd_apply_f32(const float* dout, const float* y, const float* mean, const float* invstd,
                            const float* gamma, const float* beta, const float* dbeta, const float* dgamma,
                            double total_count, float* dy, int64_t B, int64_t T, int64_t C, int64_t ldy, int64_t ldo,
                            int act, int pool, float drop_p, uint64_t seed, int drop_before_pool, int round_out,
                            void* stream) {
  if (!bn_args_ok(y, mean, invstd, gamma, beta, B, T, C, ldy, pool, drop_p) || !dout || !dbeta || !dgamma || !dy ||
      total_count <= 0)
    return XM_ERR_INVALID;
  BnActArgs a = make_bn_args(y, mean, invstd, gamma, beta, B, T, C, ldy, ldo, act, pool, drop_p, seed, drop_before_pool,
                             round_out);
  const long long R_out = B * (pool == 2 ? T / 2 : T);
  if (v4_ok(y, dout, dy, C, ldy, ldo, B * T)) {
    const dim3 blk = v4_block(C);
    XM_BN_DISPATCH(bn_act_bwd_apply_v4_kernel, act, pool,
                   <<<v4_grid(R_out, blk.y), blk, 0, (cudaStream_t)stream>>>(a, dout, dbeta, dgamma,
                                                                             (float)(1.0 / total_count), R_out, dy));
    return check_launch();
  }
  bn_act_bwd_apply_kernel<<<ew_grid(R_out * C), 256, 0, (cudaStream_t)stream>>>(a, dout, dbeta, dgamma,
                                                                              (float)(1.0 / total_count), R_out, dy);
  return check_launch();
}

int xm_seqmean_f32(const float* x, int64_t B, int64_t T, int64_t C, int64_t ldx, float* out, void* stream) {
  if (!x || !out || B <= 0 || T <= 0 || C <= 0 || B > 65535) return XM_ERR_INVALID;
  dim3 grid(ceil_div(C, 32), (unsigned)B);
  seqmean_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, T, (int)C, ldx, out);
  return check_launch();
}
int xm_seqmean_bwd_f32(const float* dout, int64_t B, int64_t T, int64_t C, int64_t lddx, float* dx, void* stream) {
  if (!dout || !dx || B <= 0 || T <= 0 || C <= 0) return XM_ERR_INVALID;
  seqmean_bwd_kernel<<<ew_grid(B * T * C), 256, 0, (cudaStream_t)stream>>>(dout, T, (int)C, lddx, B * T * C, dx);
  return check_launch();
}

static void drop_consts(float p, float& scale, uint32_t& thresh) {
  if (p > 0.f) {
    scale = 1.0f / (1.0f - p);
    double th = (double)p * 4294967296.0;
    thresh = th >= 4294967295.0 ? 4294967295u : (uint32_t)th;
    if (thresh == 0u) thresh = 1u;
  } else {
    scale = 1.0f;
    thresh = 0u;
  }
}

int xm_ln_act_fwd_f32(const float* x, const float* gamma, const float* beta, float* out, float* mean, float* rstd,
                      int64_t M, int64_t D, float eps, int act, float drop_p, uint64_t seed, void* stream) {
  if (!x || !gamma || !beta || !out || !mean || !rstd || M <= 0 || D <= 0 || !(drop_p >= 0.f && drop_p < 1.f))
    return XM_ERR_INVALID;
  float sc;
  uint32_t th;
  drop_consts(drop_p, sc, th);
  ln_act_fwd_kernel<<<ceil_div(M, 8), 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, out, mean, rstd, M, (int)D, eps, act,
                                                                      sc, th, seed);
  return check_launch();
}

int xm_ln_nblk(int64_t M) {
  int64_t n = (M + 7) / 8;
  if (n > kNumSMs * 2) n = kNumSMs * 2;
  return (int)(n < 1 ? 1 : n);
}

int xm_ln_act_bwd_f32(const float* dout, const float* x, const float* gamma, const float* beta, const float* mean,
                      const float* rstd, float* dx, float* dgamma_part, float* dbeta_part, int64_t M, int64_t D,
                      int act, float drop_p, uint64_t seed, void* stream) {
  if (!dout || !x || !gamma || !beta || !mean || !rstd || !dx || !dgamma_part || !dbeta_part || M <= 0 || D <= 0 ||
      D > 4096 || !(drop_p >= 0.f && drop_p < 1.f))
    return XM_ERR_INVALID;
  float sc;
  uint32_t th;
  drop_consts(drop_p, sc, th);
  const int per_warp = D <= 768;
  ln_act_bwd_kernel<<<xm_ln_nblk(M), 256, (per_warp ? 8 : 1) * 2 * D * sizeof(float), (cudaStream_t)stream>>>(
      dout, x, gamma, beta, mean, rstd, dx, dgamma_part, dbeta_part, M, (int)D, act, sc, th, seed, per_warp);
  return check_launch();
}

int xm_act_fwd_f32(const float* x, float* out, int64_t n, int act, float drop_p, uint64_t seed, int round_out,
                   void* stream) {
  if (!x || !out || n <= 0 || !(drop_p >= 0.f && drop_p < 1.f)) return XM_ERR_INVALID;
  float sc;
  uint32_t th;
  drop_consts(drop_p, sc, th);
  if ((n % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    act_fwd_v4_kernel<<<ew_grid(n / 4), 256, 0, (cudaStream_t)stream>>>(x, out, n / 4, act, sc, th, seed, round_out);
    return check_launch();
  }
  act_fwd_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(x, out, n, act, sc, th, seed);
  int rc = check_launch();
  if (rc == XM_OK && round_out) rc = xm_round_tf32_f32(out, out, n, stream);
  return rc;
}
int xm_act_bwd_f32(const float* dout, const float* x, float* dx, int64_t n, int act, float drop_p, uint64_t seed,
                   int round_out, void* stream) {
  if (!dout || !x || !dx || n <= 0 || !(drop_p >= 0.f && drop_p < 1.f)) return XM_ERR_INVALID;
  float sc;
  uint32_t th;
  drop_consts(drop_p, sc, th);
  if ((n % 4 == 0) &&
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0) {
    act_bwd_v4_kernel<<<ew_grid(n / 4), 256, 0, (cudaStream_t)stream>>>(dout, x, dx, n / 4, act, sc, th, seed, round_out);
    return check_launch();
  }
  act_bwd_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(dout, x, dx, n, act, sc, th, seed);
  int rc = check_launch();
  if (rc == XM_OK && round_out) rc = xm_round_tf32_f32(dx, dx, n, stream);
  return rc;
}

int xm_act_bwd_colsum_nblk(int64_t M, int64_t C) {
  if (C <= 0 || (C & 3) || C > 4096) return 0;
  const int ry = (int)(C / 4 >= 256 ? 1 : 256 / (C / 4));
  int64_t n = (M + ry - 1) / ry;
  if (n > kNumSMs * 8) n = kNumSMs * 8;
  return (int)(n < 1 ? 1 : n);
}

int xm_act_bwd_colsum_f32(const float* dout, const float* x, float* dx, int64_t M, int64_t C, int act, float drop_p,
                          uint64_t seed, int round_out, float* colsum_part, void* stream) {
  if (!dout || !x || !dx || !colsum_part || M <= 0 || !(drop_p >= 0.f && drop_p < 1.f)) return XM_ERR_INVALID;
  const int nblk = xm_act_bwd_colsum_nblk(M, C);
  if (nblk == 0 || C / 4 > 1024) return XM_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(dx)) & 15) return XM_ERR_INVALID;
  float sc;
  uint32_t th;
  drop_consts(drop_p, sc, th);
  const int tx = (int)(C / 4);
  const int ry = tx >= 256 ? 1 : 256 / tx;
  act_bwd_colsum_v4_kernel<<<nblk, dim3(tx, ry), (size_t)ry * C * sizeof(float), (cudaStream_t)stream>>>(
      dout, x, dx, M, (int)C, act, sc, th, seed, round_out, colsum_part);
  return check_launch();
}

int xm_round_tf32_f32(const float* x, float* out, int64_t n, void* stream) {
  if (!x || !out || n <= 0) return XM_ERR_INVALID;
  if ((n % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    act_fwd_v4_kernel<<<ew_grid(n / 4), 256, 0, (cudaStream_t)stream>>>(x, out, n / 4, XM_ACT_NONE, 1.f, 0u, 0ull, 1);
    return check_launch();
  }
  round_tf32_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(x, out, n);
  return check_launch();
}

int xm_split3_f32(const float* x, float* out, int64_t rows, int64_t cols, int which, int axis, void* stream) {
  if (!x || !out || rows <= 0 || cols <= 0 || (which != 0 && which != 1) || (axis != 0 && axis != 1)) return XM_ERR_INVALID;
  split3_kernel<<<ew_grid(rows * cols), 256, 0, (cudaStream_t)stream>>>(x, out, rows, cols, which, axis);
  return check_launch();
}

int xm_peer_gather_f32(const void* const* src_peers, int n_peers, int64_t elems_per_peer, float* dst, void* stream) {
  if (!src_peers || !dst || n_peers <= 0 || n_peers > 8 || elems_per_peer <= 0 || (elems_per_peer & 3)) return XM_ERR_INVALID;
  PeerPtrs pp;
  for (int r = 0; r < 8; ++r) pp.p[r] = r < n_peers ? (const float*)src_peers[r] : nullptr;
  for (int r = 0; r < n_peers; ++r)
    if (!pp.p[r] || (reinterpret_cast<uintptr_t>(pp.p[r]) & 15)) return XM_ERR_INVALID;
  if (reinterpret_cast<uintptr_t>(dst) & 15) return XM_ERR_INVALID;
  const long long n4 = elems_per_peer / 4;
  peer_gather_kernel<<<dim3((unsigned)ew_grid(n4), (unsigned)n_peers), 256, 0, (cudaStream_t)stream>>>(pp, n4, dst);
  return check_launch();
}

int xm_colsum_nsplit(int64_t M, int64_t N) {
  const long long col_blocks = (N + 127) / 128;
  long long want = (8ll * kNumSMs + col_blocks - 1) / col_blocks;  // ~8 CTAs per SM in total
  const long long max_by_rows = (M + 255) / 256;                   // >= 256 rows per block
  if (want > max_by_rows) want = max_by_rows;
  if (want < 1) want = 1;
  if (want > 65535) want = 65535;
  return (int)want;
}

int xm_colsum_f32(const float* x, int64_t M, int64_t N, int64_t ldx, float* out, float* workspace, void* stream) {
  if (!x || !out || M <= 0 || N <= 0) return XM_ERR_INVALID;
  const int ns = xm_colsum_nsplit(M, N);
  if (ns > 1 && !workspace) return XM_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const long long chunk = (M + ns - 1) / ns;
  float* dst = ns > 1 ? workspace : out;
  if ((N % 4 == 0) && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
    colsum_v4_kernel<<<dim3(ceil_div(N, 128), ns), 256, 0, st>>>(x, M, N, ldx, chunk, dst);
  } else {
    colsum_kernel<<<dim3(ceil_div(N, 32), ns), 256, 0, st>>>(x, M, N, ldx, chunk, dst);
  }
  int rc = check_launch();
  if (rc != XM_OK || ns == 1) return rc;
  colsum_kernel<<<dim3(ceil_div(N, 32), 1), 256, 0, st>>>(workspace, ns, N, N, ns, out);
  return check_launch();
}

int xm_l2norm_fwd_f32(const float* x, float* xn, float* inv_norm, int64_t M, int64_t D, float eps, void* stream) {
  if (!x || !xn || !inv_norm || M <= 0 || D <= 0) return XM_ERR_INVALID;
  l2norm_fwd_kernel<<<ceil_div(M, 8), 256, 0, (cudaStream_t)stream>>>(x, xn, inv_norm, M, (int)D, eps);
  return check_launch();
}
int xm_l2norm_split_fwd_f32(const float* x, float* xn, float* xs, float* inv_norm, int64_t M, int64_t D, float eps,
                            int which, void* stream) {
  if (!x || !xn || !xs || !inv_norm || M <= 0 || D <= 0) return XM_ERR_INVALID;
  l2norm_split_fwd_kernel<<<ceil_div(M, 8), 256, 0, (cudaStream_t)stream>>>(x, xn, xs, inv_norm, M, (int)D, eps, which);
  return check_launch();
}
int xm_l2norm_bwd_f32(const float* dxn, const float* xn, const float* inv_norm, float* dx, int64_t M, int64_t D,
                      void* stream) {
  if (!dxn || !xn || !inv_norm || !dx || M <= 0 || D <= 0) return XM_ERR_INVALID;
  l2norm_bwd_kernel<<<ceil_div(M, 8), 256, 0, (cudaStream_t)stream>>>(dxn, xn, inv_norm, dx, M, (int)D);
  return check_launch();
}

}  // extern "C"
