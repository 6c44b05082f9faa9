// Functional connectivity of an fMRI ROI time series on the device: per sample the Pearson correlation matrix of the
// ROI columns over the TR axis, flattened -- the `connectivity` input of fMRIFusionNet (fMRI_CODE/fmri_utils.py:90-103;
// the reference loads such matrices from CSV files, fmri_utils.py:161-198, and SURVEY.md section 8d defines the
// synthetic connectivity input as the flattened per-sample corrcoef of the ROI series).  Deriving it next to the ROI
// mean/std aggregation removes the largest host->device transfer of the paired step (655 of 1507 MB per 4096 samples).
//
//   xc = x - mean_t(x);  C = xc^T xc;  conn[i][j] = clip(C_ij / sqrt(C_ii C_jj), -1, 1)        (numpy.corrcoef)
//
// fp32 SIMT (1e-5 against fp64): one CTA per sample stages the (TR, ROI) series in shared memory, centres it in
// place, and every thread accumulates 8 x 4 output blocks of the upper triangle over TR from three 128-bit
// shared-memory reads per 32 FMAs; the lower triangle is the mirror.
// A constant column has C_ii = 0 and yields NaN in its row / column, as numpy.corrcoef does.
#include "xm_common.cuh"

namespace xm {
namespace conn {

constexpr int kThreads = 352;  // 650 upper-triangle blocks at ROI = 200: two rounds of 325

XM_DEVICE float nan_to_num(float v) {  // fmri_utils.py:140 applies np.nan_to_num before aggregating
  if (!(fabsf(v) <= 3.4028235e38f)) v = isnan(v) ? 0.f : copysignf(3.4028235e38f, v);
  return v;
}

// The matrix is symmetric: only the 8 x 4 blocks that touch the upper triangle are accumulated; a block writes its
// elements (i, j >= i) row-wise and mirrors the strictly upper ones to (j, i).
__global__ void __launch_bounds__(kThreads, 2)
corrcoef_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int TR, int ROI, int ROIp, int split3) {
  extern __shared__ float sm[];
  float* xs = sm;                  // [TR][ROIp], columns >= ROI are zero
  float* inv = sm + TR * ROIp;     // [ROIp] 1 / sqrt(C_ii)
  unsigned short* tiles = reinterpret_cast<unsigned short*>(inv + ROIp);  // (ti, tj) of every block touching the upper triangle
  const int nti = (ROIp + 7) >> 3, ntj = ROIp >> 2;  // 8-row x 4-column blocks
  for (int ti = threadIdx.x; ti < nti; ti += kThreads) {
    int k = ti * ntj - ti * (ti - 1);  // blocks of the rows above: sum over t < ti of (ntj - 2 t)
    for (int tj = 2 * ti; tj < ntj; ++tj, ++k) {
      tiles[2 * k] = (unsigned short)ti;
      tiles[2 * k + 1] = (unsigned short)tj;
    }
  }
  const int n_tiles = nti * ntj - nti * (nti - 1);
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const float* xb = x + (long long)b * TR * ROI;
    if (ROI == ROIp && (reinterpret_cast<uintptr_t>(xb) & 15) == 0) {  // the sample is one contiguous, aligned block
      const float4* x4 = reinterpret_cast<const float4*>(xb);
      for (int i = threadIdx.x; i < (TR * ROI) >> 2; i += kThreads) {
        const float4 v = __ldg(x4 + i);
        reinterpret_cast<float4*>(xs)[i] = make_float4(nan_to_num(v.x), nan_to_num(v.y), nan_to_num(v.z), nan_to_num(v.w));
      }
    } else {
      for (int t = threadIdx.x / 32; t < TR; t += kThreads / 32)
        for (int r = threadIdx.x & 31; r < ROIp; r += 32) xs[t * ROIp + r] = r < ROI ? nan_to_num(xb[(long long)t * ROI + r]) : 0.f;
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
    float* ob = out + (long long)b * ROI * ROI;
    for (int tile = threadIdx.x; tile < n_tiles; tile += kThreads) {
      const int i0 = tiles[2 * tile] * 8, j0 = tiles[2 * tile + 1] * 4;
      float acc[8][4];
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
      const bool hi_ok = i0 + 4 < ROIp;  // ROIp is a multiple of 4, not of 8: the last row block may be half
      for (int t = 0; t < TR; ++t) {
        const float* row = xs + t * ROIp;
        const float4 a0 = *reinterpret_cast<const float4*>(row + i0);
        const float4 a1 = hi_ok ? *reinterpret_cast<const float4*>(row + i0 + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 vb = *reinterpret_cast<const float4*>(row + j0);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, bv[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(av[a], bv[c], acc[a][c]);
      }
      // scale, clip, store.  Planes: 1 (plain) or 3 (row-stacked tf32 split [hi; hi; lo], see xm_roi_corrcoef_f32).
      const long long plane = (long long)B * ROI * ROI;
      const bool interior = j0 >= i0 + 8 && i0 + 8 <= ROI && j0 + 4 <= ROI && (ROI & 3) == 0;
      float iv[8], jv[4];
#pragma unroll
      for (int a = 0; a < 8; ++a) iv[a] = i0 + a < ROIp ? inv[i0 + a] : 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) jv[c] = inv[j0 + c];
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float v = acc[a][c] * iv[a] * jv[c];
          acc[a][c] = v != v ? v : fminf(fmaxf(v, -1.0f), 1.0f);  // fminf / fmaxf alone would drop a NaN
        }
      if (interior) {
        // strictly above the diagonal and inside the matrix: the block itself as 8 rows of 16 B, its mirror as 4 rows
        // of 32 B (whole sectors), per plane
        for (int pl = 0; pl < (split3 ? 3 : 1); ++pl) {
          float* o = ob + pl * plane;
          auto part = [&](float v) {  // this plane's share of a value
            if (!split3) return v;
            const float hi = round_tf32(v);
            return pl < 2 ? hi : round_tf32(v - hi);
          };
#pragma unroll
          for (int a = 0; a < 8; ++a)
            *reinterpret_cast<float4*>(o + (long long)(i0 + a) * ROI + j0) =
                make_float4(part(acc[a][0]), part(acc[a][1]), part(acc[a][2]), part(acc[a][3]));
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float4* m = reinterpret_cast<float4*>(o + (long long)(j0 + c) * ROI + i0);
            m[0] = make_float4(part(acc[0][c]), part(acc[1][c]), part(acc[2][c]), part(acc[3][c]));
            m[1] = make_float4(part(acc[4][c]), part(acc[5][c]), part(acc[6][c]), part(acc[7][c]));
          }
        }
      } else {
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          const int i = i0 + a;
          if (i >= ROI) break;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int j = j0 + c;
            if (j >= ROI || j < i) continue;
            const float v = acc[a][c];
            const float hi = round_tf32(v), lo = round_tf32(v - hi);
            for (int pl = 0; pl < (split3 ? 3 : 1); ++pl) {
              const float w = !split3 ? v : (pl < 2 ? hi : lo);
              ob[pl * plane + (long long)i * ROI + j] = w;
              if (j > i) ob[pl * plane + (long long)j * ROI + i] = w;
            }
          }
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
  return TR >= 2 && ROI >= 1 && ROI <= 4096 && (TR * roip + roip) * 4 + roip * roip / 8 <= 220 * 1024;
}

extern "C" int xm_roi_corrcoef_f32(const float* x, int64_t B, int64_t TR, int64_t ROI, float* out, int split3, void* stream) {
  if (!x || !out || B <= 0 || TR <= 0 || ROI <= 0) return XM_ERR_INVALID;
  if (!xm_roi_corrcoef_supported(TR, ROI)) return XM_ERR_UNSUPPORTED;
  const int roip = (int)((ROI + 3) / 4 * 4);
  const size_t smem = (size_t)(TR * roip + roip) * sizeof(float) + (size_t)((roip + 7) / 8) * (roip / 4) * 4;  // + block table
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(conn::corrcoef_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    g_last_cuda_error = (int)cudaGetLastError();
    return XM_ERR_LAUNCH;
  }
  const int per_sm = (int)(220 * 1024 / smem) < 1 ? 1 : (int)(220 * 1024 / smem);
  const int64_t grid = B < (int64_t)kNumSMs * per_sm ? B : (int64_t)kNumSMs * per_sm;
  conn::corrcoef_kernel<<<(int)grid, conn::kThreads, smem, (cudaStream_t)stream>>>(x, out, (int)B, (int)TR, (int)ROI, roip, split3);
  return check_launch();
}
