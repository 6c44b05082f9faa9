// Functional connectivity of an fMRI ROI time series on the device: per sample the Pearson correlation matrix of the
// ROI columns over the TR axis, flattened -- the `connectivity` input of fMRIFusionNet (fMRI_CODE/fmri_utils.py:90-103;
// the reference loads such matrices from CSV files, fmri_utils.py:161-198, and SURVEY.md section 8d defines the
// synthetic connectivity input as the flattened per-sample corrcoef of the ROI series).  Deriving it next to the ROI
// mean/std aggregation removes the largest host->device transfer of the paired step (655 of 1507 MB per 4096 samples).
//
//   xc = x - mean_t(x);  C = xc^T xc;  conn[i][j] = clip(C_ij / sqrt(C_ii C_jj), -1, 1)        (numpy.corrcoef)
//
// fp32 SIMT (1e-5 against fp64): one CTA per sample stages the (TR, ROI) series in shared memory, centres it in
// place, and every thread accumulates 4 x 4 output blocks over TR from two 128-bit shared-memory reads per 16 FMAs.
// A constant column has C_ii = 0 and yields NaN in its row / column, as numpy.corrcoef does.
#include "xm_common.cuh"

namespace xm {
namespace conn {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
corrcoef_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int TR, int ROI, int ROIp) {
  extern __shared__ float sm[];
  float* xs = sm;                  // [TR][ROIp], columns >= ROI are zero
  float* inv = sm + TR * ROIp;     // [ROIp] 1 / sqrt(C_ii)
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const float* xb = x + (long long)b * TR * ROI;
    for (int i = threadIdx.x; i < TR * ROIp; i += kThreads) {
      const int t = i / ROIp, r = i - t * ROIp;
      float v = r < ROI ? xb[(long long)t * ROI + r] : 0.f;
      if (!(fabsf(v) <= 3.4028235e38f)) v = isnan(v) ? 0.f : copysignf(3.4028235e38f, v);  // nan_to_num (fmri_utils.py:140)
      xs[i] = v;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < ROIp; r += kThreads) {  // centre each column, 1 / norm
      float s = 0.f;
      for (int t = 0; t < TR; ++t) s += xs[t * ROIp + r];
      const float mu = s / (float)TR;
      float q = 0.f;
      for (int t = 0; t < TR; ++t) {
        const float v = xs[t * ROIp + r] - mu;
        xs[t * ROIp + r] = v;
        q = fmaf(v, v, q);
      }
      inv[r] = 1.0f / sqrtf(q);  // q == 0 -> inf -> 0 * inf = NaN below (numpy: 0 / 0)
    }
    __syncthreads();
    const int nt = ROIp >> 2;  // 4 x 4 blocks per side
    float* ob = out + (long long)b * ROI * ROI;
    for (int tile = threadIdx.x; tile < nt * nt; tile += kThreads) {
      const int ti = tile / nt, tj = tile - ti * nt;
      const int i0 = ti * 4, j0 = tj * 4;
      float acc[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
      for (int t = 0; t < TR; ++t) {
        const float4 va = *reinterpret_cast<const float4*>(xs + t * ROIp + i0);
        const float4 vb = *reinterpret_cast<const float4*>(xs + t * ROIp + j0);
        const float av[4] = {va.x, va.y, va.z, va.w}, bv[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(av[a], bv[c], acc[a][c]);
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int i = i0 + a;
        if (i >= ROI) break;
        const float ii = inv[i];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int j = j0 + c;
          if (j < ROI) ob[(long long)i * ROI + j] = fminf(fmaxf(acc[a][c] * ii * inv[j], -1.0f), 1.0f);  // NaN stays NaN
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace conn
}  // namespace xm

using namespace xm;

extern "C" int xm_roi_corrcoef_supported(int64_t TR, int64_t ROI) {
  const int64_t roip = (ROI + 3) / 4 * 4;
  return TR >= 2 && ROI >= 1 && (TR * roip + roip) * 4 <= 220 * 1024;
}

extern "C" int xm_roi_corrcoef_f32(const float* x, int64_t B, int64_t TR, int64_t ROI, float* out, void* stream) {
  if (!x || !out || B <= 0 || TR <= 0 || ROI <= 0) return XM_ERR_INVALID;
  if (!xm_roi_corrcoef_supported(TR, ROI)) return XM_ERR_UNSUPPORTED;
  const int roip = (int)((ROI + 3) / 4 * 4);
  const size_t smem = (size_t)(TR * roip + roip) * sizeof(float);
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(conn::corrcoef_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    g_last_cuda_error = (int)cudaGetLastError();
    return XM_ERR_LAUNCH;
  }
  const int per_sm = (int)(220 * 1024 / smem) < 1 ? 1 : (int)(220 * 1024 / smem);
  const int64_t grid = B < (int64_t)kNumSMs * per_sm ? B : (int64_t)kNumSMs * per_sm;
  conn::corrcoef_kernel<<<(int)grid, conn::kThreads, smem, (cudaStream_t)stream>>>(x, out, (int)B, (int)TR, (int)ROI, roip);
  return check_launch();
}
