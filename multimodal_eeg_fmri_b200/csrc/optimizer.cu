// clip_grad_norm_ + AdamW over ONE flat bucket in two launches (the step recipe of _test_bridge.py:775-788,869 /
// run_fmri_v11.py:430-450 / run_training_lite.py:478-489: clip_grad_norm_(max_norm) then AdamW.step).  The
// reference's torch path issues a foreach norm, a clamp, a foreach multiply and the optimizer's foreach kernels per
// parameter group; here every parameter, gradient and moment lives in one flat fp32 buffer:
//   1. xm_sumsq_partials_f32: per-block sum of squares of the gradient bucket in fp64 (deterministic two-stage sum)
//   2. xm_clip_adamw_f32: every block folds the partials into the total norm, derives the clip coefficient
//      min(1, max_norm / (norm + 1e-6)) (torch.nn.utils.clip_grad_norm_) and applies torch.optim.AdamW's update
//        p *= 1 - lr * wd;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;
//        p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
//      with g already scaled by the clip coefficient; the clipped gradient is written back (the bucket then holds
//      what clip_grad_norm_ leaves in .grad).
#include "xm_common.cuh"

namespace xm {
namespace opt {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) sumsq_partials_kernel(const float* __restrict__ g, long long n, double* __restrict__ partials) {
  double acc = 0.0;
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < n4; i += (long long)gridDim.x * kThreads) {
    const float4 v = g4[i];
    acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += kThreads) acc += (double)g[i] * g[i];
  __shared__ double red[kThreads / 32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < kThreads / 32 ? red[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0) partials[blockIdx.x] = v;
  }
}

struct AdamArgs {
  float max_norm, lr, beta1, beta2, eps, weight_decay, bias_c1, bias_c2_sqrt;
};

__global__ void __launch_bounds__(kThreads)
clip_adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                  const double* __restrict__ partials, int nblk, AdamArgs a, float* __restrict__ norm_out,
                  const float* __restrict__ lr_dev, const long long* __restrict__ step_dev) {
  __shared__ double red[kThreads / 32];
  __shared__ float s_coef, s_lr, s_c1, s_c2;
  if (threadIdx.x == 32) {  // step count / learning rate from device memory (a captured step replays with fresh values)
    s_lr = lr_dev != nullptr ? *lr_dev : a.lr;
    s_c1 = a.bias_c1;
    s_c2 = a.bias_c2_sqrt;
    if (step_dev != nullptr) {
      const double t = (double)*step_dev;
      s_c1 = (float)(1.0 - pow((double)a.beta1, t));
      s_c2 = (float)sqrt(1.0 - pow((double)a.beta2, t));
    }
  }
  double acc = 0.0;
  for (int i = threadIdx.x; i < nblk; i += kThreads) acc += partials[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = threadIdx.x < kThreads / 32 ? red[threadIdx.x] : 0.0;
    t = warp_sum(t);
    if (threadIdx.x == 0) {
      const float norm = (float)sqrt(t);
      s_coef = a.max_norm > 0.f ? fminf(a.max_norm / (norm + 1e-6f), 1.0f) : 1.0f;
      if (blockIdx.x == 0 && norm_out != nullptr) *norm_out = norm;
    }
  }
  __syncthreads();
  const float coef = s_coef;
  const float decay = 1.0f - s_lr * a.weight_decay, step = s_lr / s_c1, inv_c2 = 1.0f / s_c2;
  for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
    const float gi = g[i] * coef;
    const float mi = a.beta1 * m[i] + (1.0f - a.beta1) * gi;
    const float vi = a.beta2 * v[i] + (1.0f - a.beta2) * gi * gi;
    g[i] = gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = p[i] * decay - step * mi / (sqrtf(vi) * inv_c2 + a.eps);
  }
}


// Gather of per-parameter gradient tensors into the flat bucket in ONE launch (autograd hands every parameter its own
// gradient tensor; accumulating each into a preset view of the bucket costs one tiny add kernel per parameter and
// step).  blockIdx.y = tensor, blockIdx.x strides over its elements.
constexpr int kMaxGather = 96;
struct GatherArgs {
  const float* src[kMaxGather];
  long long dst_off[kMaxGather];
  long long numel[kMaxGather];
};
__global__ void __launch_bounds__(kThreads) gather_flat_kernel(const __grid_constant__ GatherArgs a, float* __restrict__ dst) {
  const int t = blockIdx.y;
  const float* __restrict__ s = a.src[t];
  float* __restrict__ d = dst + a.dst_off[t];
  const long long n = a.numel[t];
  if ((((uintptr_t)s | (uintptr_t)d) & 15) == 0) {
    const long long n4 = n >> 2;
    for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < n4; i += (long long)gridDim.x * kThreads)
      reinterpret_cast<float4*>(d)[i] = reinterpret_cast<const float4*>(s)[i];
    if (blockIdx.x == 0)
      for (long long i = (n4 << 2) + threadIdx.x; i < n; i += kThreads) d[i] = s[i];
  } else {
    for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) d[i] = s[i];
  }
}

}  // namespace opt
}  // namespace xm

using namespace xm;

extern "C" int xm_sumsq_nblk(int64_t n) {
  int64_t b = (n / 4 + opt::kThreads - 1) / opt::kThreads;
  if (b > kNumSMs * 4) b = kNumSMs * 4;
  return (int)(b < 1 ? 1 : b);
}

extern "C" int xm_sumsq_partials_f32(const float* g, int64_t n, double* partials, void* stream) {
  if (!g || !partials || n <= 0 || (reinterpret_cast<uintptr_t>(g) & 15)) return XM_ERR_INVALID;
  opt::sumsq_partials_kernel<<<xm_sumsq_nblk(n), opt::kThreads, 0, (cudaStream_t)stream>>>(g, n, partials);
  return check_launch();
}

extern "C" int xm_clip_adamw_f32(float* p, float* g, float* m, float* v, int64_t n, const double* partials, int nblk, float max_norm,
                                 float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step, float* norm_out,
                                 void* stream) {
  if (!p || !g || !m || !v || !partials || n <= 0 || nblk <= 0 || step <= 0) return XM_ERR_INVALID;
  opt::AdamArgs a;
  a.max_norm = max_norm; a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
  a.bias_c1 = (float)(1.0 - pow((double)beta1, (double)step));
  a.bias_c2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  int64_t blocks = (n + opt::kThreads - 1) / opt::kThreads;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  opt::clip_adamw_kernel<<<(int)blocks, opt::kThreads, 0, (cudaStream_t)stream>>>(p, g, m, v, n, partials, nblk, a, norm_out,
                                                                                  nullptr, nullptr);
  return check_launch();
}

extern "C" int xm_clip_adamw_dev_f32(float* p, float* g, float* m, float* v, int64_t n, const double* partials, int nblk,
                                     float max_norm, const float* lr_dev, float beta1, float beta2, float eps, float weight_decay,
                                     const int64_t* step_dev, float* norm_out, void* stream) {
  if (!p || !g || !m || !v || !partials || !lr_dev || !step_dev || n <= 0 || nblk <= 0) return XM_ERR_INVALID;
  opt::AdamArgs a;
  a.max_norm = max_norm; a.lr = 0.f; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
  a.bias_c1 = 1.f; a.bias_c2_sqrt = 1.f;
  int64_t blocks = (n + opt::kThreads - 1) / opt::kThreads;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  opt::clip_adamw_kernel<<<(int)blocks, opt::kThreads, 0, (cudaStream_t)stream>>>(
      p, g, m, v, n, partials, nblk, a, norm_out, lr_dev, reinterpret_cast<const long long*>(step_dev));
  return check_launch();
}

// ---- seed epoch: one device counter, mirrored into the __constant__ slot of every translation unit that hashes ----
extern "C" {
void* xm_seed_epoch_slot_elementwise();
void* xm_seed_epoch_slot_transformer();
void* xm_seed_epoch_slot_bridge_head();
void* xm_seed_epoch_slot_attention_general();
void* xm_seed_epoch_slot_attention_fused();
void* xm_seed_epoch_slot_ffn_fused();
void* xm_seed_epoch_slot_peer_exchange();
}

namespace xm {
namespace opt {
constexpr int kEpochSlots = 7;
struct EpochState {
  unsigned long long* counter = nullptr;
  void* slots[kEpochSlots] = {};
  int device = -1;
};
static EpochState g_epoch;

__global__ void seed_epoch_kernel(unsigned long long* counter, unsigned long long value, int add) {
  *counter = add ? *counter + value : value;
}

static int epoch_ready() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) return XM_ERR_LAUNCH;
  if (g_epoch.counter != nullptr) return dev == g_epoch.device ? XM_OK : XM_ERR_INVALID;  // one GPU per process
  void* (*const get[kEpochSlots])() = {xm_seed_epoch_slot_elementwise,        xm_seed_epoch_slot_transformer,
                                       xm_seed_epoch_slot_bridge_head,        xm_seed_epoch_slot_attention_general,
                                       xm_seed_epoch_slot_attention_fused,    xm_seed_epoch_slot_ffn_fused,
                                       xm_seed_epoch_slot_peer_exchange};
  for (int i = 0; i < kEpochSlots; ++i)
    if ((g_epoch.slots[i] = get[i]()) == nullptr) return XM_ERR_LAUNCH;
  if (cudaMalloc(&g_epoch.counter, sizeof(unsigned long long)) != cudaSuccess) return XM_ERR_LAUNCH;
  if (cudaMemset(g_epoch.counter, 0, sizeof(unsigned long long)) != cudaSuccess) return XM_ERR_LAUNCH;
  g_epoch.device = dev;
  return XM_OK;
}

static int epoch_update(unsigned long long value, int add, cudaStream_t st) {
  if (g_epoch.counter == nullptr) return XM_ERR_INVALID;  // xm_seed_epoch_init first (it allocates: not capturable)
  seed_epoch_kernel<<<1, 1, 0, st>>>(g_epoch.counter, value, add);
  for (int i = 0; i < kEpochSlots; ++i)
    if (cudaMemcpyAsync(g_epoch.slots[i], g_epoch.counter, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
      return XM_ERR_LAUNCH;
  return check_launch();
}
}  // namespace opt
}  // namespace xm

extern "C" int xm_seed_epoch_init(void) { return opt::epoch_ready(); }
extern "C" int xm_seed_epoch_advance(void* stream) { return opt::epoch_update(1ull, 1, (cudaStream_t)stream); }
extern "C" int xm_seed_epoch_set(uint64_t value, void* stream) { return opt::epoch_update(value, 0, (cudaStream_t)stream); }
extern "C" int xm_seed_epoch_get(uint64_t* value_out) {
  if (!value_out || opt::g_epoch.counter == nullptr) return XM_ERR_INVALID;
  return cudaMemcpy(value_out, opt::g_epoch.counter, sizeof(uint64_t), cudaMemcpyDeviceToHost) == cudaSuccess ? XM_OK : XM_ERR_LAUNCH;
}

extern "C" int xm_gather_flat_f32(const void* const* src, const int64_t* dst_off, const int64_t* numel, int n_tensors, float* dst,
                                  void* stream) {
  if (!src || !dst_off || !numel || !dst || n_tensors <= 0) return XM_ERR_INVALID;
  for (int t0 = 0; t0 < n_tensors; t0 += opt::kMaxGather) {
    const int nt = n_tensors - t0 < opt::kMaxGather ? n_tensors - t0 : opt::kMaxGather;
    opt::GatherArgs a{};
    long long biggest = 0;
    for (int i = 0; i < nt; ++i) {
      if (!src[t0 + i] || numel[t0 + i] < 0 || dst_off[t0 + i] < 0) return XM_ERR_INVALID;
      a.src[i] = (const float*)src[t0 + i];
      a.dst_off[i] = dst_off[t0 + i];
      a.numel[i] = numel[t0 + i];
      if (numel[t0 + i] > biggest) biggest = numel[t0 + i];
    }
    long long bx = (biggest / 4 + opt::kThreads - 1) / opt::kThreads;
    if (bx < 1) bx = 1;
    if (bx > 2 * kNumSMs) bx = 2 * kNumSMs;
    opt::gather_flat_kernel<<<dim3((unsigned)bx, (unsigned)nt), opt::kThreads, 0, (cudaStream_t)stream>>>(a, dst);
    int rc = check_launch();
    if (rc != XM_OK) return rc;
  }
  return XM_OK;
}
