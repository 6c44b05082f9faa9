// Shape-general multi-head self-attention core (sm_100a, SIMT fp32): every shape the fused tcgen05 kernel
// (attention_fused.cu: head dim 32, L <= 512, no mask) does not cover -- any head dim <= 256, any sequence length
// that fits the per-warp score rows in shared memory (L <= ~13 000), optional additive mask, dropout on the
// probabilities.  nn.MultiheadAttention inside TemporalTransformerBlock (EEG_CODE/enhanced_models_v4.py:71-73,98)
// for the small / odd configurations (hidden 32 with 4 heads in the reference's own smoke test,
// enhanced_models_v4.py:844-890) and for explicit attn_mask arguments.  Not a throughput kernel: one warp per
// query row (forward, dq) or per key row (dk / dv), scores recomputed in the backward from the saved logsumexp,
// nothing of size L x L stored, no atomics (deterministic).
//
//   S = scale * q k^T + mask,  P = softmax_j(S),  P~ = keep * P / (1 - p_drop),  O = P~ v
//   dP~ = dO v^T,  dP = keep * dP~ / (1 - p_drop),  delta_i = sum_j P_ij dP_ij,  dS = P * (dP - delta)
//   dq = scale * dS k,  dk = scale * dS^T q,  dv = P~^T dO
#include "xm_common.cuh"

namespace xm {
namespace ga {

constexpr int kWarps = 4;

struct Args {
  const float* qkv;    // (B, L, 3E) packed in_proj output, E = H * dh
  const float* mask;   // additive, (L, L) [mask_stride 0] or (B*H, L, L) [mask_stride L*L]; may be null
  long long mask_stride;
  int B, L, H, dh;
  float scale, dscale;
  uint32_t thresh;     // keep <=> hash >= thresh (0: no dropout)
  uint64_t seed;
};

XM_DEVICE float keep_scale(const Args& a, long long bh, int i, int j) {
  if (a.thresh == 0u) return 1.0f;
  return dropout_keep((uint64_t)((bh * a.L + i) * (long long)a.L + j), a.seed, a.thresh) ? a.dscale : 0.0f;
}

XM_DEVICE float dot_row(const float* __restrict__ s_vec, const float* __restrict__ g_row, int dh) {
  float acc = 0.f;
  for (int d = 0; d < dh; ++d) acc = fmaf(s_vec[d], __ldg(g_row + d), acc);
  return acc;
}

// ---- forward: one warp per (sample, head, query)
__global__ void __launch_bounds__(kWarps * 32)
attn_general_fwd_kernel(const Args a, float* __restrict__ out, float* __restrict__ lse, int round_out) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sc = smem + (size_t)warp * (a.L + a.dh);  // scores / probabilities of this row
  float* sq = sc + a.L;                            // the query vector
  const int E = a.H * a.dh;
  const long long rows = (long long)a.B * a.H * a.L;
  for (long long r = (long long)blockIdx.x * kWarps + warp; r < rows; r += (long long)gridDim.x * kWarps) {
    const int i = (int)(r % a.L);
    const long long bh = r / a.L;
    const int h = (int)(bh % a.H);
    const long long b = bh / a.H;
    const float* base = a.qkv + b * a.L * 3ll * E + (long long)h * a.dh;
    for (int d = lane; d < a.dh; d += 32) sq[d] = base[(long long)i * 3 * E + d];
    __syncwarp();
    const float* mrow = a.mask ? a.mask + bh * a.mask_stride + (long long)i * a.L : nullptr;
    float mx = -INFINITY;
    for (int j = lane; j < a.L; j += 32) {
      float s = a.scale * dot_row(sq, base + (long long)j * 3 * E + E, a.dh);
      if (mrow) s += mrow[j];
      sc[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < a.L; j += 32) {
      const float e = expf(sc[j] - mx);  // a fully masked row gives exp(-inf + inf) = NaN, as torch's softmax does
      sc[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int j = lane; j < a.L; j += 32) sc[j] = sc[j] * inv * keep_scale(a, bh, i, j);
    if (lane == 0) lse[r] = mx + logf(sum);
    __syncwarp();
    for (int d = lane; d < a.dh; d += 32) {
      const float* v = base + 2 * E + d;
      float acc = 0.f;
      for (int j = 0; j < a.L; ++j) acc = fmaf(sc[j], __ldg(v + (long long)j * 3 * E), acc);
      out[(b * a.L + i) * (long long)E + h * a.dh + d] = round_out ? round_tf32(acc) : acc;
    }
    __syncwarp();
  }
}

// ---- backward 1: one warp per query row -> dq, delta
__global__ void __launch_bounds__(kWarps * 32)
attn_general_bwd_q_kernel(const Args a, const float* __restrict__ dout, const float* __restrict__ lse,
                          float* __restrict__ dqkv, float* __restrict__ delta, int round_out) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sc = smem + (size_t)warp * (a.L + 2 * a.dh);
  float* sq = sc + a.L;
  float* sdo = sq + a.dh;
  const int E = a.H * a.dh;
  const long long rows = (long long)a.B * a.H * a.L;
  for (long long r = (long long)blockIdx.x * kWarps + warp; r < rows; r += (long long)gridDim.x * kWarps) {
    const int i = (int)(r % a.L);
    const long long bh = r / a.L;
    const int h = (int)(bh % a.H);
    const long long b = bh / a.H;
    const float* base = a.qkv + b * a.L * 3ll * E + (long long)h * a.dh;
    for (int d = lane; d < a.dh; d += 32) {
      sq[d] = base[(long long)i * 3 * E + d];
      sdo[d] = dout[(b * a.L + i) * (long long)E + h * a.dh + d];
    }
    __syncwarp();
    const float* mrow = a.mask ? a.mask + bh * a.mask_stride + (long long)i * a.L : nullptr;
    const float l = lse[r];
    float dl = 0.f;
    for (int j = lane; j < a.L; j += 32) {
      float s = a.scale * dot_row(sq, base + (long long)j * 3 * E + E, a.dh);
      if (mrow) s += mrow[j];
      const float p = expf(s - l);
      const float dp = dot_row(sdo, base + (long long)j * 3 * E + 2 * E, a.dh) * keep_scale(a, bh, i, j);
      dl = fmaf(p, dp, dl);
      sc[j] = dp;           // dP_ij for now
      // p is recomputed below (one expf more per element; keeps the row buffer single)
    }
    dl = warp_sum(dl);
    if (lane == 0) delta[r] = dl;
    for (int j = lane; j < a.L; j += 32) {
      float s = a.scale * dot_row(sq, base + (long long)j * 3 * E + E, a.dh);
      if (mrow) s += mrow[j];
      sc[j] = a.scale * expf(s - l) * (sc[j] - dl);  // scale * dS_ij
    }
    __syncwarp();
    for (int d = lane; d < a.dh; d += 32) {
      const float* k = base + E + d;
      float acc = 0.f;
      for (int j = 0; j < a.L; ++j) acc = fmaf(sc[j], __ldg(k + (long long)j * 3 * E), acc);
      dqkv[(b * a.L + i) * 3ll * E + h * a.dh + d] = round_out ? round_tf32(acc) : acc;
    }
    __syncwarp();
  }
}

// ---- backward 2: one warp per key row -> dk, dv
__global__ void __launch_bounds__(kWarps * 32)
attn_general_bwd_kv_kernel(const Args a, const float* __restrict__ dout, const float* __restrict__ lse,
                           const float* __restrict__ delta, float* __restrict__ dqkv, int round_out) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sp = smem + (size_t)warp * (2 * a.L + 2 * a.dh);  // P~_ij over i (dv weights)
  float* sd = sp + a.L;                                    // scale * dS_ij over i (dk weights)
  float* sk = sd + a.L;
  float* sv = sk + a.dh;
  const int E = a.H * a.dh;
  const long long rows = (long long)a.B * a.H * a.L;
  for (long long r = (long long)blockIdx.x * kWarps + warp; r < rows; r += (long long)gridDim.x * kWarps) {
    const int j = (int)(r % a.L);
    const long long bh = r / a.L;
    const int h = (int)(bh % a.H);
    const long long b = bh / a.H;
    const float* base = a.qkv + b * a.L * 3ll * E + (long long)h * a.dh;
    const float* dob = dout + b * a.L * (long long)E + (long long)h * a.dh;
    for (int d = lane; d < a.dh; d += 32) {
      sk[d] = base[(long long)j * 3 * E + E + d];
      sv[d] = base[(long long)j * 3 * E + 2 * E + d];
    }
    __syncwarp();
    const float* mcol = a.mask ? a.mask + bh * a.mask_stride + j : nullptr;
    for (int i = lane; i < a.L; i += 32) {
      float s = a.scale * dot_row(sk, base + (long long)i * 3 * E, a.dh);
      if (mcol) s += mcol[(long long)i * a.L];
      const float p = expf(s - lse[bh * a.L + i]);
      const float ks = keep_scale(a, bh, i, j);
      const float dp = dot_row(sv, dob + (long long)i * E, a.dh) * ks;
      sp[i] = p * ks;
      sd[i] = a.scale * p * (dp - delta[bh * a.L + i]);
    }
    __syncwarp();
    for (int d = lane; d < a.dh; d += 32) {
      float dk = 0.f, dv = 0.f;
      for (int i = 0; i < a.L; ++i) {
        dk = fmaf(sd[i], __ldg(base + (long long)i * 3 * E + d), dk);
        dv = fmaf(sp[i], __ldg(dob + (long long)i * E + d), dv);
      }
      float* o = dqkv + (b * a.L + j) * 3ll * E + h * a.dh + d;
      o[E] = round_out ? round_tf32(dk) : dk;
      o[2 * E] = round_out ? round_tf32(dv) : dv;
    }
    __syncwarp();
  }
}

static int make_args(Args& a, const float* qkv, const float* mask, int mask_per_head, int64_t B, int64_t L, int64_t H,
                     int64_t dh, float scale, float drop_p, uint64_t seed) {
  if (!qkv || B <= 0 || L <= 0 || H <= 0 || dh <= 0 || !(drop_p >= 0.f && drop_p < 1.f)) return XM_ERR_INVALID;
  if (dh > 256 || B * H * L > (int64_t)1 << 40) return XM_ERR_UNSUPPORTED;
  a.qkv = qkv;
  a.mask = mask;
  a.mask_stride = mask_per_head ? L * L : 0;
  a.B = (int)B; a.L = (int)L; a.H = (int)H; a.dh = (int)dh;
  a.scale = scale;
  a.dscale = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
  a.thresh = drop_p > 0.f ? (uint32_t)((double)drop_p * 4294967296.0) : 0u;
  a.seed = seed;
  return XM_OK;
}

template <typename K>
static int prepare(K kernel, size_t smem) {
  if (smem > 227 * 1024) return XM_ERR_UNSUPPORTED;
  if (smem > 48 * 1024 && cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    g_last_cuda_error = (int)cudaGetLastError();
    return XM_ERR_LAUNCH;
  }
  return XM_OK;
}

static int grid_for(int64_t rows) {
  const int64_t blocks = (rows + kWarps - 1) / kWarps;
  return (int)(blocks < (int64_t)kNumSMs * 64 ? blocks : (int64_t)kNumSMs * 64);
}

}  // namespace ga
}  // namespace xm

XM_DEFINE_SEED_EPOCH_SLOT(attention_general)

using namespace xm;

extern "C" int xm_attn_general_supported(int64_t L, int64_t dh) {
  return dh > 0 && dh <= 256 && L > 0 && (size_t)ga::kWarps * (2 * L + 2 * dh) * sizeof(float) <= 227 * 1024;
}

extern "C" int xm_attn_general_fwd_f32(const float* qkv, const float* mask, int mask_per_head, float* out, float* lse, int64_t B,
                                       int64_t L, int64_t H, int64_t dh, float scale, float drop_p, uint64_t seed, int round_out,
                                       void* stream) {
  ga::Args a;
  int rc = ga::make_args(a, qkv, mask, mask_per_head, B, L, H, dh, scale, drop_p, seed);
  if (rc != XM_OK) return rc;
  if (!out || !lse) return XM_ERR_INVALID;
  const size_t smem = (size_t)ga::kWarps * (L + dh) * sizeof(float);
  if ((rc = ga::prepare(ga::attn_general_fwd_kernel, smem)) != XM_OK) return rc;
  ga::attn_general_fwd_kernel<<<ga::grid_for(B * H * L), ga::kWarps * 32, smem, (cudaStream_t)stream>>>(a, out, lse, round_out);
  return check_launch();
}

extern "C" int xm_attn_general_bwd_f32(const float* dout, const float* qkv, const float* mask, int mask_per_head, const float* lse,
                                       float* dqkv, float* delta, int64_t B, int64_t L, int64_t H, int64_t dh, float scale,
                                       float drop_p, uint64_t seed, int round_out, void* stream) {
  ga::Args a;
  int rc = ga::make_args(a, qkv, mask, mask_per_head, B, L, H, dh, scale, drop_p, seed);
  if (rc != XM_OK) return rc;
  if (!dout || !lse || !dqkv || !delta) return XM_ERR_INVALID;
  const size_t smem_q = (size_t)ga::kWarps * (L + 2 * dh) * sizeof(float);
  const size_t smem_kv = (size_t)ga::kWarps * (2 * L + 2 * dh) * sizeof(float);
  if ((rc = ga::prepare(ga::attn_general_bwd_q_kernel, smem_q)) != XM_OK) return rc;
  if ((rc = ga::prepare(ga::attn_general_bwd_kv_kernel, smem_kv)) != XM_OK) return rc;
  const int grid = ga::grid_for(B * H * L);
  ga::attn_general_bwd_q_kernel<<<grid, ga::kWarps * 32, smem_q, (cudaStream_t)stream>>>(a, dout, lse, dqkv, delta, round_out);
  if ((rc = check_launch()) != XM_OK) return rc;
  ga::attn_general_bwd_kv_kernel<<<grid, ga::kWarps * 32, smem_kv, (cudaStream_t)stream>>>(a, dout, lse, delta, dqkv, round_out);
  return check_launch();
}
