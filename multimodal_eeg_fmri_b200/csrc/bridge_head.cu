// Cross-attention core of the bridge's supervised head (bridge_utils.py:74-83): ONE EEG query per sample attends over
// the two-token sequence [eeg, fmri] through nn.MultiheadAttention -- per (sample, head) two scores, a two-way softmax,
// dropout on the weights and a weighted sum of two value vectors.  The reference runs it as a batched matmul + softmax +
// dropout + matmul over (B*H, 1, 2) tensors; here one warp per (sample, head) does all of it, forward and backward
// (the probabilities are recomputed from q and k; the dropout mask is a pure function of (seed, sample, head, token)).
//
//   q (B, d);  kv (2B, 2d): row t*B + b = token t of sample b, columns [0, d) keys, [d, 2d) values;  d = H * dh
//   forward : s_t = q . k_t / sqrt(dh),  p = softmax_t(s),  w_t = mask_t p_t / (1 - drop),  o = sum_t w_t v_t
//   backward: dv_t = w_t do,  dw_t = do . v_t,  dp_t = mask_t dw_t / (1 - drop),  ds_t = p_t (dp_t - sum_u p_u dp_u),
//             dq = sum_t ds_t k_t / sqrt(dh),  dk_t = ds_t q / sqrt(dh)
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/xmodal_b200.h"
#include "xm_common.cuh"

namespace xm {
namespace head {

XM_DEVICE float keep_scale(unsigned long long idx, unsigned long long seed, uint32_t thr, float scale) {
  return thr == 0u ? 1.0f : (dropout_keep(idx, seed, thr) ? scale : 0.0f);
}

// one warp per (sample, head); BWD also writes dq / dkv
template <bool BWD>
__global__ void __launch_bounds__(256)
cross2_kernel(const float* __restrict__ q, const float* __restrict__ kv, const float* __restrict__ dout, float* __restrict__ out,
              float* __restrict__ att, float* __restrict__ dq, float* __restrict__ dkv, long long B, int H, int dh, float inv_sqrt,
              float drop_scale, uint32_t thr, unsigned long long seed) {
  const long long w = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= B * H) return;
  const int lane = threadIdx.x & 31;
  const long long b = w / H;
  const int h = (int)(w - b * H), d = H * dh;
  const float* qr = q + b * d + h * dh;
  const float* k0 = kv + b * 2 * d + h * dh;
  const float* k1 = kv + (B + b) * 2 * d + h * dh;
  const float* v0 = k0 + d;
  const float* v1 = k1 + d;
  float s0 = 0.f, s1 = 0.f;
  for (int j = lane; j < dh; j += 32) {
    s0 = fmaf(qr[j], k0[j], s0);
    s1 = fmaf(qr[j], k1[j], s1);
  }
  s0 = warp_sum(s0) * inv_sqrt;
  s1 = warp_sum(s1) * inv_sqrt;
  const float m = fmaxf(s0, s1);
  const float e0 = __expf(s0 - m), e1 = __expf(s1 - m);
  const float inv = 1.0f / (e0 + e1);
  const float p0 = e0 * inv, p1 = e1 * inv;
  const float m0 = keep_scale((unsigned long long)w * 2, seed, thr, drop_scale);
  const float m1 = keep_scale((unsigned long long)w * 2 + 1, seed, thr, drop_scale);
  const float w0 = p0 * m0, w1 = p1 * m1;
  if (!BWD) {
    for (int j = lane; j < dh; j += 32) out[b * d + h * dh + j] = fmaf(w0, v0[j], w1 * v1[j]);
    if (lane == 0) {
      att[w * 2] = w0;
      att[w * 2 + 1] = w1;
    }
    return;
  }
  const float* dor = dout + b * d + h * dh;
  float a0 = 0.f, a1 = 0.f;  // dw_t = do . v_t
  for (int j = lane; j < dh; j += 32) {
    a0 = fmaf(dor[j], v0[j], a0);
    a1 = fmaf(dor[j], v1[j], a1);
  }
  const float dp0 = warp_sum(a0) * m0, dp1 = warp_sum(a1) * m1;
  const float dot = p0 * dp0 + p1 * dp1;
  const float ds0 = p0 * (dp0 - dot) * inv_sqrt, ds1 = p1 * (dp1 - dot) * inv_sqrt;
  float* dk0 = dkv + b * 2 * d + h * dh;
  float* dk1 = dkv + (B + b) * 2 * d + h * dh;
  for (int j = lane; j < dh; j += 32) {
    const float qj = qr[j], dj = dor[j];
    dq[b * d + h * dh + j] = fmaf(ds0, k0[j], ds1 * k1[j]);
    dk0[j] = ds0 * qj;
    dk1[j] = ds1 * qj;
    dk0[d + j] = w0 * dj;
    dk1[d + j] = w1 * dj;
  }
}

}  // namespace head
}  // namespace xm

XM_DEFINE_SEED_EPOCH_SLOT(bridge_head)

using namespace xm;

extern "C" {

int xm_cross2_attn_fwd_f32(const float* q, const float* kv, float* out, float* att, int64_t B, int64_t H, int64_t dh, float drop_p,
                           uint64_t seed, void* stream) {
  if (!q || !kv || !out || !att || B <= 0 || H <= 0 || dh <= 0 || !(drop_p >= 0.f && drop_p < 1.f)) return XM_ERR_INVALID;
  const uint32_t thr = drop_p > 0.f ? (uint32_t)((double)drop_p * 4294967296.0) : 0u;
  const long long warps = B * H;
  head::cross2_kernel<false><<<(unsigned)((warps + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      q, kv, nullptr, out, att, nullptr, nullptr, B, (int)H, (int)dh, 1.0f / sqrtf((float)dh), 1.0f / (1.0f - drop_p), thr, seed);
  return check_launch();
}

int xm_cross2_attn_bwd_f32(const float* dout, const float* q, const float* kv, float* dq, float* dkv, int64_t B, int64_t H, int64_t dh,
                           float drop_p, uint64_t seed, void* stream) {
  if (!dout || !q || !kv || !dq || !dkv || B <= 0 || H <= 0 || dh <= 0 || !(drop_p >= 0.f && drop_p < 1.f)) return XM_ERR_INVALID;
  const uint32_t thr = drop_p > 0.f ? (uint32_t)((double)drop_p * 4294967296.0) : 0u;
  const long long warps = B * H;
  head::cross2_kernel<true><<<(unsigned)((warps + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      q, kv, dout, nullptr, nullptr, dq, dkv, B, (int)H, (int)dh, 1.0f / sqrtf((float)dh), 1.0f / (1.0f - drop_p), thr, seed);
  return check_launch();
}

}  // extern "C"
