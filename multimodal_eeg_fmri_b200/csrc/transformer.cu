// Residual-stream kernels of the pre-norm transformer block (EEG_CODE/enhanced_models_v4.py:89-107):
//
//   s = x + Dropout(a)          (a = the previous branch output: attention out-projection or FFN)
//   h = LayerNorm(s) * g + b    -> rounded to tf32, because h is the A operand of the next projection GEMM
//
// fused into ONE pass over the (M, D) token matrix, forward and backward, plus the final
// "residual add + mean over time" that feeds the encoder head.  One warp per token row, float4 per lane,
// row statistics by warp shuffle; dropout masks are regenerated from (seed, element index).
// HBM-bound: forward reads x, a and writes s, h (16*D B per token); backward reads dh, dres, s and writes
// dx, da (20*D B per token).  D % 128 == 0, D <= 512.
#include "xm_common.cuh"

namespace xm {

constexpr int kLnWarps = 8;

template <int NV>  // NV = D / 128 float4 per lane
__global__ void __launch_bounds__(kLnWarps * 32)
resid_ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ a, const float* __restrict__ pe, long long L,
                    const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ s_out,
                    float* __restrict__ h, float* __restrict__ mean, float* __restrict__ rstd, long long M, float eps,
                    float drop_scale, uint32_t drop_thresh, uint64_t seed) {
  constexpr int D = NV * 128;
  const long long row = blockIdx.x * (long long)kLnWarps + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  float v[NV][4];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = i * 128 + lane * 4;
    const float4 xv = *reinterpret_cast<const float4*>(x + row * D + c);
    v[i][0] = xv.x; v[i][1] = xv.y; v[i][2] = xv.z; v[i][3] = xv.w;
    if (pe != nullptr) {  // positional table row (row mod L), added before the dropout (PositionalEncoding.forward)
      const float4 pv = *reinterpret_cast<const float4*>(pe + (row % L) * D + c);
      v[i][0] += pv.x; v[i][1] += pv.y; v[i][2] += pv.z; v[i][3] += pv.w;
      if (drop_thresh) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          v[i][j] = dropout_keep((uint64_t)(row * D + c + j), seed, drop_thresh) ? v[i][j] * drop_scale : 0.f;
      }
    }
    if (a != nullptr) {
      const float4 av = *reinterpret_cast<const float4*>(a + row * D + c);
      float t[4] = {av.x, av.y, av.z, av.w};
      if (drop_thresh) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          t[j] = dropout_keep((uint64_t)(row * D + c + j), seed, drop_thresh) ? t[j] * drop_scale : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) v[i][j] += t[j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) sum += v[i][j];
  }
  const float mu = warp_sum(sum) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float d = v[i][j] - mu;
      q += d * d;
    }
  const float rs = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
  if (lane == 0) {
    mean[row] = mu;
    rstd[row] = rs;
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = i * 128 + lane * 4;
    if (s_out != nullptr)
      *reinterpret_cast<float4*>(s_out + row * D + c) = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
    const float4 g = *reinterpret_cast<const float4*>(gamma + c);
    const float4 b = *reinterpret_cast<const float4*>(beta + c);
    float4 o;
    o.x = round_tf32((v[i][0] - mu) * rs * g.x + b.x);
    o.y = round_tf32((v[i][1] - mu) * rs * g.y + b.y);
    o.z = round_tf32((v[i][2] - mu) * rs * g.z + b.z);
    o.w = round_tf32((v[i][3] - mu) * rs * g.w + b.w);
    *reinterpret_cast<float4*>(h + row * D + c) = o;
  }
}

// Backward: ds = dres + LayerNormBackward(dh);  dx = ds;  da = mask * scale * ds (tf32: operand of the
// previous projection's dgrad / wgrad).  Per-block partial sums of dgamma / dbeta -> (gridDim.x, D).
template <int NV>
__global__ void __launch_bounds__(kLnWarps * 32)
resid_ln_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ dres, const float* __restrict__ s,
                    const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                    float* __restrict__ dx, float* __restrict__ da, float* __restrict__ dgamma_part,
                    float* __restrict__ dbeta_part, float* __restrict__ dabias_part, long long M, float drop_scale,
                    uint32_t drop_thresh, uint64_t seed) {
  constexpr int D = NV * 128;
  __shared__ float red[3][kLnWarps][D];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float ag[NV][4], ab[NV][4], ac[NV][4];  // dgamma, dbeta, column sums of da (bias gradient of the branch's Linear)
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) ag[i][j] = ab[i][j] = ac[i][j] = 0.f;
  float4 g4[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) g4[i] = *reinterpret_cast<const float4*>(gamma + i * 128 + lane * 4);

  for (long long row = blockIdx.x * (long long)kLnWarps + warp; row < M; row += (long long)gridDim.x * kLnWarps) {
    const float mu = mean[row], rs = rstd[row];
    float xh[NV][4], dz[NV][4];
    float s1 = 0.f, s2 = 0.f;
    float4 rv4[NV];  // the residual-stream gradient is needed only after the two reductions, but its load goes out with the others
#pragma unroll
    for (int i = 0; i < NV; ++i)
      rv4[i] = dres != nullptr ? *reinterpret_cast<const float4*>(dres + row * D + i * 128 + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 128 + lane * 4;
      const float4 sv = *reinterpret_cast<const float4*>(s + row * D + c);
      const float4 dv = *reinterpret_cast<const float4*>(dh + row * D + c);
      const float sa[4] = {sv.x, sv.y, sv.z, sv.w}, da4[4] = {dv.x, dv.y, dv.z, dv.w};
      const float ga[4] = {g4[i].x, g4[i].y, g4[i].z, g4[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        xh[i][j] = (sa[j] - mu) * rs;
        ag[i][j] += da4[j] * xh[i][j];
        ab[i][j] += da4[j];
        dz[i][j] = da4[j] * ga[j];
        s1 += dz[i][j];
        s2 += dz[i][j] * xh[i][j];
      }
    }
    s1 = warp_sum(s1) * (1.0f / D);
    s2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 128 + lane * 4;
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = rs * (dz[i][j] - s1 - xh[i][j] * s2);
      o[0] += rv4[i].x; o[1] += rv4[i].y; o[2] += rv4[i].z; o[3] += rv4[i].w;
      *reinterpret_cast<float4*>(dx + row * D + c) = make_float4(o[0], o[1], o[2], o[3]);
      if (da != nullptr) {
        float t[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          t[j] = o[j];
          if (drop_thresh) t[j] = dropout_keep((uint64_t)(row * D + c + j), seed, drop_thresh) ? t[j] * drop_scale : 0.f;
          ac[i][j] += t[j];
          t[j] = round_tf32(t[j]);
        }
        *reinterpret_cast<float4*>(da + row * D + c) = make_float4(t[0], t[1], t[2], t[3]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      red[0][warp][i * 128 + lane * 4 + j] = ag[i][j];
      red[1][warp][i * 128 + lane * 4 + j] = ab[i][j];
      red[2][warp][i * 128 + lane * 4 + j] = ac[i][j];
    }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float sg = 0.f, sb = 0.f, sc = 0.f;
#pragma unroll
    for (int w = 0; w < kLnWarps; ++w) {
      sg += red[0][w][c];
      sb += red[1][w][c];
      sc += red[2][w][c];
    }
    dgamma_part[(long long)blockIdx.x * D + c] = sg;
    dbeta_part[(long long)blockIdx.x * D + c] = sb;
    if (dabias_part != nullptr) dabias_part[(long long)blockIdx.x * D + c] = sc;
  }
}

// out[b, c] = mean_t (x[b, t, c] + Dropout(a)[b, t, c]);  grid (B), block 32 float4-lanes x 8 row lanes, D == 128*NV
template <int NV>
__global__ void __launch_bounds__(256)
resid_seqmean_fwd_kernel(const float* __restrict__ x, const float* __restrict__ a, long long T, float* __restrict__ out,
                         float drop_scale, uint32_t drop_thresh, uint64_t seed) {
  constexpr int D = NV * 128;
  __shared__ float red[8][D];
  const int lane = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long b = blockIdx.x;
  float acc[NV][4];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long long t = ty; t < T; t += 8) {
    const long long row = b * T + t;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 128 + lane * 4;
      const float4 xv = *reinterpret_cast<const float4*>(x + row * D + c);
      acc[i][0] += xv.x; acc[i][1] += xv.y; acc[i][2] += xv.z; acc[i][3] += xv.w;
      if (a != nullptr) {
        const float4 av = *reinterpret_cast<const float4*>(a + row * D + c);
        float tt[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (drop_thresh) tt[j] = dropout_keep((uint64_t)(row * D + c + j), seed, drop_thresh) ? tt[j] * drop_scale : 0.f;
          acc[i][j] += tt[j];
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) red[ty][i * 128 + lane * 4 + j] = acc[i][j];
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += 256) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[w][c];
    out[b * D + c] = sum / (float)T;
  }
}

// dx[b, t, :] = dout[b, :] / T ;  da = mask * scale * dx (tf32-rounded)
__global__ void resid_seqmean_bwd_kernel(const float* __restrict__ dout, long long T, int D, long long total4,
                                         float* __restrict__ dx, float* __restrict__ da, float drop_scale,
                                         uint32_t drop_thresh, uint64_t seed) {
  const int D4 = D / 4;
  const float inv_t = 1.0f / (float)T;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / D4;
    const int c = (int)(i - row * D4) * 4;
    const float4 g = *reinterpret_cast<const float4*>(dout + (row / T) * D + c);
    float o[4] = {g.x * inv_t, g.y * inv_t, g.z * inv_t, g.w * inv_t};
    if (dx != nullptr) *reinterpret_cast<float4*>(dx + row * D + c) = make_float4(o[0], o[1], o[2], o[3]);
    if (da != nullptr) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (drop_thresh) o[j] = dropout_keep((uint64_t)(row * D + c + j), seed, drop_thresh) ? o[j] * drop_scale : 0.f;
        o[j] = round_tf32(o[j]);
      }
      *reinterpret_cast<float4*>(da + row * D + c) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
}

static void drop_consts32(float p, float& scale, uint32_t& thresh) {
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
static bool aligned16(const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace xm

XM_DEFINE_SEED_EPOCH_SLOT(transformer)

using namespace xm;

extern "C" {

int xm_resid_ln_supported(int64_t D) { return (D % 128 == 0 && D >= 128 && D <= 512) ? 1 : 0; }

int xm_resid_ln_nblk(int64_t M) {
  int64_t n = (M + kLnWarps - 1) / kLnWarps;
  if (n > kNumSMs * 8) n = kNumSMs * 8;
  return (int)(n < 1 ? 1 : n);
}

int xm_resid_ln_fwd_f32(const float* x, const float* a, const float* pe, int64_t L, const float* gamma, const float* beta,
                        float* s_out, float* h, float* mean, float* rstd, int64_t M, int64_t D, float eps, float drop_p,
                        uint64_t seed, void* stream) {
  if (!x || !gamma || !beta || !h || !mean || !rstd || M <= 0 || !(drop_p >= 0.f && drop_p < 1.f)) return XM_ERR_INVALID;
  if ((a != nullptr || pe != nullptr) && s_out == nullptr) return XM_ERR_INVALID;
  if (pe != nullptr && L <= 0) return XM_ERR_INVALID;
  if (!xm_resid_ln_supported(D)) return XM_ERR_UNSUPPORTED;
  if (!aligned16(x) || !aligned16(a) || !aligned16(pe) || !aligned16(gamma) || !aligned16(beta) || !aligned16(s_out) ||
      !aligned16(h))
    return XM_ERR_INVALID;
  float sc;
  uint32_t th;
  drop_consts32(drop_p, sc, th);
  const int blocks = ceil_div(M, kLnWarps);
  cudaStream_t st = (cudaStream_t)stream;
  switch (D / 128) {
    case 1: resid_ln_fwd_kernel<1><<<blocks, kLnWarps * 32, 0, st>>>(x, a, pe, L, gamma, beta, s_out, h, mean, rstd, M, eps, sc, th, seed); break;
    case 2: resid_ln_fwd_kernel<2><<<blocks, kLnWarps * 32, 0, st>>>(x, a, pe, L, gamma, beta, s_out, h, mean, rstd, M, eps, sc, th, seed); break;
    case 3: resid_ln_fwd_kernel<3><<<blocks, kLnWarps * 32, 0, st>>>(x, a, pe, L, gamma, beta, s_out, h, mean, rstd, M, eps, sc, th, seed); break;
    default: resid_ln_fwd_kernel<4><<<blocks, kLnWarps * 32, 0, st>>>(x, a, pe, L, gamma, beta, s_out, h, mean, rstd, M, eps, sc, th, seed); break;
  }
  return check_launch();
}

int xm_resid_ln_bwd_f32(const float* dh, const float* dres, const float* s, const float* gamma, const float* mean,
                        const float* rstd, float* dx, float* da, float* dgamma_part, float* dbeta_part, float* dabias_part,
                        int64_t M, int64_t D, float drop_p, uint64_t seed, void* stream) {
  if (!dh || !s || !gamma || !mean || !rstd || !dx || !dgamma_part || !dbeta_part || M <= 0 ||
      !(drop_p >= 0.f && drop_p < 1.f))
    return XM_ERR_INVALID;
  if (!xm_resid_ln_supported(D)) return XM_ERR_UNSUPPORTED;
  if (!aligned16(dh) || !aligned16(dres) || !aligned16(s) || !aligned16(gamma) || !aligned16(dx) || !aligned16(da))
    return XM_ERR_INVALID;
  float sc;
  uint32_t th;
  drop_consts32(drop_p, sc, th);
  const int blocks = xm_resid_ln_nblk(M);
  cudaStream_t st = (cudaStream_t)stream;
  switch (D / 128) {
    case 1: resid_ln_bwd_kernel<1><<<blocks, kLnWarps * 32, 0, st>>>(dh, dres, s, gamma, mean, rstd, dx, da, dgamma_part, dbeta_part, dabias_part, M, sc, th, seed); break;
    case 2: resid_ln_bwd_kernel<2><<<blocks, kLnWarps * 32, 0, st>>>(dh, dres, s, gamma, mean, rstd, dx, da, dgamma_part, dbeta_part, dabias_part, M, sc, th, seed); break;
    case 3: resid_ln_bwd_kernel<3><<<blocks, kLnWarps * 32, 0, st>>>(dh, dres, s, gamma, mean, rstd, dx, da, dgamma_part, dbeta_part, dabias_part, M, sc, th, seed); break;
    default: resid_ln_bwd_kernel<4><<<blocks, kLnWarps * 32, 0, st>>>(dh, dres, s, gamma, mean, rstd, dx, da, dgamma_part, dbeta_part, dabias_part, M, sc, th, seed); break;
  }
  return check_launch();
}

int xm_resid_seqmean_fwd_f32(const float* x, const float* a, int64_t B, int64_t T, int64_t D, float* out, float drop_p,
                             uint64_t seed, void* stream) {
  if (!x || !out || B <= 0 || T <= 0 || !(drop_p >= 0.f && drop_p < 1.f)) return XM_ERR_INVALID;
  if (!xm_resid_ln_supported(D)) return XM_ERR_UNSUPPORTED;
  if (!aligned16(x) || !aligned16(a)) return XM_ERR_INVALID;
  float sc;
  uint32_t th;
  drop_consts32(drop_p, sc, th);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)B;
  switch (D / 128) {
    case 1: resid_seqmean_fwd_kernel<1><<<blocks, 256, 0, st>>>(x, a, T, out, sc, th, seed); break;
    case 2: resid_seqmean_fwd_kernel<2><<<blocks, 256, 0, st>>>(x, a, T, out, sc, th, seed); break;
    case 3: resid_seqmean_fwd_kernel<3><<<blocks, 256, 0, st>>>(x, a, T, out, sc, th, seed); break;
    default: resid_seqmean_fwd_kernel<4><<<blocks, 256, 0, st>>>(x, a, T, out, sc, th, seed); break;
  }
  return check_launch();
}

int xm_resid_seqmean_bwd_f32(const float* dout, int64_t B, int64_t T, int64_t D, float* dx, float* da, float drop_p,
                             uint64_t seed, void* stream) {
  if (!dout || (!dx && !da) || B <= 0 || T <= 0 || D <= 0 || (D & 3) || !(drop_p >= 0.f && drop_p < 1.f)) return XM_ERR_INVALID;
  if (!aligned16(dout) || !aligned16(dx) || !aligned16(da)) return XM_ERR_INVALID;
  float sc;
  uint32_t th;
  drop_consts32(drop_p, sc, th);
  const long long total4 = B * T * (D / 4);
  long long blocks = (total4 + 255) / 256;
  if (blocks > (long long)kNumSMs * 16) blocks = (long long)kNumSMs * 16;
  resid_seqmean_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dout, T, (int)D, total4, dx, da, sc, th, seed);
  return check_launch();
}

}  // extern "C"
