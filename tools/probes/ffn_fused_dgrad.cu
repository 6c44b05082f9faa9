// PROBE (not part of the library, not on the product path): fused transformer FFN backward, data-gradient half, for
// d_model = 128, hidden = 512 -- DESIGN.md section 7, "first cut".  Written at the end of round 1 WITHOUT a GPU at
// hand: it compiles for sm_100a and carries its own checker, but has never run.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lineinfo \
//        -o tools/probes/ffn_fused_dgrad tools/probes/ffn_fused_dgrad.cu && tools/probes/ffn_fused_dgrad [rows] [drop_p]
//
// Given x (M, 128), dY (M, 128), W1 (512, 128), b1, and the transposed copies W2t = W2^T (512, 128), W1t = W1^T
// (128, 512) -- so that every B operand is K-major -- one pass over 128-row tiles recomputes the hidden activations
// and produces   A  = round(dropout(gelu(x W1^T + b1)))                (M, 512)   operand of dW2 = A^T dY
//                dH = round(dY W2 * gelu'(x W1^T + b1) * dropout mask)  (M, 512)   operand of dW1 = dH^T x, db1
//                dX = dH W1                                             (M, 128)
// for the existing weight-gradient kernels.  Nothing is read that the unfused path had to save (h and its activation,
// 2 x 2.1 GB per block at the bench shape).
// Per tile and hidden chunk c of 128:  M1(c): H_c = x W1_c^T -> TMEM H[c & 1];  M3(c): G_c = dY W2t_c^T -> TMEM G;
//   T(c): 8 transform warps read H_c and G_c, store A_c and dH_c to global memory and write dH_c over G_c;
//   M4(c): Xacc += dH_c W1t[:, chunk c]^T with dH_c as the TMEM-resident A operand.
// MMA issue order per tile (tcgen05.mma executes in issue order, so TMEM reuse between MMAs needs no barrier):
//   M1(0) M3(0) M1(1) | M4(0) M3(1) M1(2) | M4(1) M3(2) M1(3) | M4(2) M3(3) | M4(3)      (each M4 waits for its T)
// TMEM: H0 [0,128) H1 [128,256) G [256,384) Xacc [384,512).  Shared memory: x tile 64 KB + dY tile 64 KB + 6 x 16 KB
// weight ring = 224 KB.  192 MMAs per tile (~10.8 us at the tf32 rate) against 704 KB of HBM per tile (~16 us): HBM-bound,
// ~0.9 ms per block at the bench shape (today: dgrad 0.55 + act-backward 1.0 + dgrad 0.55 ms).
// Known first tuning point: A_c / dH_c are stored from registers, one row per lane (32 lines per instruction); stage
// 32-column slabs through shared memory + TMA stores once the numbers say the LSU is the limit (ring of 4 frees 32 KB).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../multimodal_eeg_fmri_b200/csrc/xm_common.cuh"
#include "../../multimodal_eeg_fmri_b200/csrc/xm_ptx.cuh"

namespace xm {
int g_last_cuda_error = 0;

namespace ffn {

constexpr int kD = 128, kH = 512, kChunk = 128, kNC = kH / kChunk;
constexpr int kRing = 6;
constexpr int kTile = 16384;  // 128 rows x 32 fp32, SWIZZLE_128B
constexpr int kSmem = 2 * 4 * kTile + kRing * kTile + 1024;
constexpr int kThreads = 64 + 8 * 32;
constexpr int kOps = 12;
__device__ __constant__ int kOpKind[kOps] = {0, 1, 0, 2, 1, 0, 2, 1, 0, 2, 1, 2};   // 0: M1, 1: M3, 2: M4
__device__ __constant__ int kOpChunk[kOps] = {0, 0, 1, 0, 1, 2, 1, 2, 3, 2, 3, 3};

struct Params {
  long long M;
  int tiles;
  const float* b1;
  float* a;    // (M, 512)
  float* dh;   // (M, 512)
  float* dx;   // (M, 128)
  float drop_scale;
  uint32_t drop_thresh;
  uint64_t seed;
};

struct Bars {
  uint64_t in_full, in_empty;
  uint64_t w_full[kRing], w_empty[kRing];
  uint64_t g_full, d_ready;
  uint64_t x_done, xacc_free;
};

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

__global__ void __launch_bounds__(kThreads, 1)
ffn_dgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                 const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2t,
                 const __grid_constant__ CUtensorMap tmW1t, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ Bars bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* xs = smem;                    // [4] x k-block tiles (K = input features)
  uint8_t* ys = smem + 4 * kTile;        // [4] dY k-block tiles (K = output features)
  uint8_t* ring = smem + 8 * kTile;      // [kRing] weight k-block tiles
  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmX);
    ptx::prefetch_tensormap(&tmDY);
    ptx::prefetch_tensormap(&tmW1);
    ptx::prefetch_tensormap(&tmW2t);
    ptx::prefetch_tensormap(&tmW1t);
    ptx::mbar_init(&bar.in_full, 1);
    ptx::mbar_init(&bar.in_empty, 1);
    for (int i = 0; i < kRing; ++i) {
      ptx::mbar_init(&bar.w_full[i], 1);
      ptx::mbar_init(&bar.w_empty[i], 1);
    }
    ptx::mbar_init(&bar.g_full, 1);
    ptx::mbar_init(&bar.d_ready, 8);
    ptx::mbar_init(&bar.x_done, 1);
    ptx::mbar_init(&bar.xacc_free, 8);
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
  const uint32_t tG = tmem + 256u, tX = tmem + 384u;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t ws = 0;  // weight k-blocks requested so far
      int it = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        ptx::mbar_wait(&bar.in_empty, ((uint32_t)it & 1u) ^ 1u);
        ptx::mbar_arrive_expect_tx(&bar.in_full, 8 * kTile);
        for (int kb = 0; kb < 4; ++kb) {
          ptx::tma_load_3d(&tmX, &bar.in_full, xs + kb * kTile, kb * 32, tile * 128, 0);
          ptx::tma_load_3d(&tmDY, &bar.in_full, ys + kb * kTile, kb * 32, tile * 128, 0);
        }
        for (int op = 0; op < kOps; ++op) {
          const int c = kOpChunk[op], kind = kOpKind[op];
          for (int kb = 0; kb < 4; ++kb, ++ws) {
            const uint32_t st = ws % kRing;
            ptx::mbar_wait(&bar.w_empty[st], ((ws / kRing) & 1u) ^ 1u);
            ptx::mbar_arrive_expect_tx(&bar.w_full[st], kTile);
            if (kind == 0)
              ptx::tma_load_3d(&tmW1, &bar.w_full[st], ring + st * kTile, kb * 32, c * kChunk, 0);   // W1[128c.., in 32kb..]
            else if (kind == 1)
              ptx::tma_load_3d(&tmW2t, &bar.w_full[st], ring + st * kTile, kb * 32, c * kChunk, 0);  // W2t[128c.., out 32kb..]
            else
              ptx::tma_load_3d(&tmW1t, &bar.w_full[st], ring + st * kTile, c * kChunk + kb * 32, 0, 0);  // W1t[:, 128c + 32kb..]
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc_tf32(128, 128, 0, 0);
      uint32_t ws = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        for (int op = 0; op < kOps; ++op) {
          const int c = kOpChunk[op], kind = kOpKind[op];
          const uint32_t tH = tmem + (uint32_t)((c & 1) * 128);
          const uint32_t cu = 4u * (uint32_t)it + (uint32_t)c;  // chunks processed before this one
          if (kind == 0 && c == 0) {
            ptx::mbar_wait(&bar.in_full, (uint32_t)it & 1u);
            ptx::tc_fence_after_sync();
          }
          if (kind == 2) {
            ptx::mbar_wait(&bar.d_ready, cu & 1u);  // the transform warps have read H_c, G_c and written dH_c over G_c
            ptx::tc_fence_after_sync();
            if (c == 0) {
              ptx::mbar_wait(&bar.xacc_free, ((uint32_t)it & 1u) ^ 1u);  // the epilogue has read the previous tile's dX
              ptx::tc_fence_after_sync();
            }
          }
          for (int kb = 0; kb < 4; ++kb, ++ws) {
            const uint32_t st = ws % kRing;
            ptx::mbar_wait(&bar.w_full[st], (ws / kRing) & 1u);
            ptx::tc_fence_after_sync();
            const uint64_t db = ptx::make_smem_desc(ptx::smem_u32(ring + st * kTile), 16, 1024, 2);
            if (kind == 2) {
#pragma unroll
              for (int k8 = 0; k8 < 4; ++k8)
                mma_tf32_ts(tX, tG + (uint32_t)(kb * 32 + k8 * 8), db + (uint64_t)(k8 * 2), idesc, (c | kb | k8) ? 1u : 0u);
            } else {
              const uint64_t da = ptx::make_smem_desc(ptx::smem_u32((kind == 0 ? xs : ys) + kb * kTile), 16, 1024, 2);
              const uint32_t acc = kind == 0 ? tH : tG;
#pragma unroll
              for (int k8 = 0; k8 < 4; ++k8)
                ptx::mma_tf32_ss(acc, da + (uint64_t)(k8 * 2), db + (uint64_t)(k8 * 2), idesc, (kb | k8) ? 1u : 0u);
            }
            ptx::mma_commit(&bar.w_empty[st]);
          }
          if (kind == 1) {
            ptx::mma_commit(&bar.g_full);  // H_c (issued earlier) and G_c are complete
            if (c == kNC - 1) ptx::mma_commit(&bar.in_empty);
          }
          if (kind == 2 && c == kNC - 1) ptx::mma_commit(&bar.x_done);
        }
      }
    }
  } else {
    const int q = warp & 3;            // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;  // which 64 of the chunk's 128 columns
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
      const long long row = (long long)tile * 128 + q * 32 + lane;
      for (int c = 0; c < kNC; ++c) {
        const uint32_t cu = 4u * (uint32_t)it + (uint32_t)c;
        ptx::mbar_wait(&bar.g_full, cu & 1u);
        ptx::tc_fence_after_sync();
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int col0 = half * 64 + j * 32;
          uint32_t rh[32], rg[32];
          ptx::tmem_ld_32x32(tmem + (uint32_t)((c & 1) * 128 + col0) + lane_base, rh);
          ptx::tmem_ld_32x32(tG + (uint32_t)col0 + lane_base, rg);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int hcol = c * kChunk + col0 + e;
            const float h = __uint_as_float(rh[e]) + __ldg(&p.b1[hcol]);
            float cdf, pdf;
            normal_cdf_pdf(h, cdf, pdf);
            float m = 1.0f;
            if (p.drop_thresh) m = dropout_keep((uint64_t)(row * kH + hcol), p.seed, p.drop_thresh) ? p.drop_scale : 0.f;
            rh[e] = __float_as_uint(round_tf32(h * cdf * m));                                  // A
            rg[e] = __float_as_uint(round_tf32(__uint_as_float(rg[e]) * fmaf(h, pdf, cdf) * m));  // dH
          }
          ptx::tmem_st_32x32(tG + (uint32_t)col0 + lane_base, rg);
          if (row < p.M) {
            float4* da = reinterpret_cast<float4*>(p.a + row * kH + c * kChunk + col0);
            float4* dd = reinterpret_cast<float4*>(p.dh + row * kH + c * kChunk + col0);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              da[e] = make_float4(__uint_as_float(rh[4 * e]), __uint_as_float(rh[4 * e + 1]), __uint_as_float(rh[4 * e + 2]),
                                  __uint_as_float(rh[4 * e + 3]));
              dd[e] = make_float4(__uint_as_float(rg[4 * e]), __uint_as_float(rg[4 * e + 1]), __uint_as_float(rg[4 * e + 2]),
                                  __uint_as_float(rg[4 * e + 3]));
            }
          }
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bar.d_ready);
      }
      // ---- tile done: dX
      ptx::mbar_wait(&bar.x_done, (uint32_t)it & 1u);
      ptx::tc_fence_after_sync();
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int col0 = half * 64 + j * 32;
        uint32_t r[32];
        ptx::tmem_ld_32x32(tX + (uint32_t)col0 + lane_base, r);
        ptx::tmem_ld_wait();
        if (row < p.M) {
          float4* dst = reinterpret_cast<float4*>(p.dx + row * kD + col0);
#pragma unroll
          for (int e = 0; e < 8; ++e)
            dst[e] = make_float4(__uint_as_float(r[4 * e]), __uint_as_float(r[4 * e + 1]), __uint_as_float(r[4 * e + 2]),
                                 __uint_as_float(r[4 * e + 3]));
        }
      }
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bar.xacc_free);
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem, 512);
  }
}

// fp64 reference of a few rows (one block per row): out = [A (512) | dH (512) | dX (128)]
__global__ void ffn_dgrad_reference_kernel(const float* x, const float* dy, const float* w1, const float* b1, const float* w2t,
                                           const long long* rows, double* out, float drop_scale, uint32_t drop_thresh,
                                           uint64_t seed) {
  __shared__ float dh[kH];
  const long long row = rows[blockIdx.x];
  double* o = out + (long long)blockIdx.x * (2 * kH + kD);
  for (int h = threadIdx.x; h < kH; h += blockDim.x) {
    double s = 0.0, g = 0.0;
    for (int k = 0; k < kD; ++k) {
      s += (double)x[row * kD + k] * (double)w1[h * kD + k];
      g += (double)dy[row * kD + k] * (double)w2t[h * kD + k];
    }
    const float hv = (float)s + b1[h];
    float cdf, pdf;
    normal_cdf_pdf(hv, cdf, pdf);
    float m = 1.0f;
    if (drop_thresh) m = dropout_keep((uint64_t)(row * kH + h), seed, drop_thresh) ? drop_scale : 0.f;
    o[h] = (double)round_tf32(hv * cdf * m);
    dh[h] = round_tf32((float)g * fmaf(hv, pdf, cdf) * m);
    o[kH + h] = (double)dh[h];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kD; i += blockDim.x) {
    double s = 0.0;
    for (int h = 0; h < kH; ++h) s += (double)dh[h] * (double)w1[h * kD + i];
    o[2 * kH + i] = s;
  }
}

__global__ void transpose_kernel(const float* in, float* out, int rows, int cols) {  // out (cols, rows) = in (rows, cols)^T
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows * cols; i += gridDim.x * blockDim.x)
    out[(i % cols) * rows + i / cols] = in[i];
}

__global__ void fill_kernel(float* p, long long n, uint64_t seed, float scale, int round) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float u1 = (hash_u32((uint64_t)i, seed) + 0.5f) * (1.0f / 4294967296.0f);
    const float u2 = (hash_u32((uint64_t)i, seed ^ 0x9E3779B97F4A7C15ull) + 0.5f) * (1.0f / 4294967296.0f);
    const float v = scale * sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
    p[i] = round ? round_tf32(v) : v;
  }
}

}  // namespace ffn
}  // namespace xm

using namespace xm;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode2d(EncodeTiledFn fn, CUtensorMap* out, const float* ptr, unsigned long long cols, unsigned long long rows) {
  cuuint64_t dims[3] = {cols, rows, 1};
  cuuint64_t strides[2] = {cols * 4, cols * rows * 4};
  cuuint32_t box[3] = {32, 128, 1}, estr[3] = {1, 1, 1};
  return (int)fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));   \
      return 1;                                                                            \
    }                                                                                      \
  } while (0)

int main(int argc, char** argv) {
  const long long M = argc > 1 ? atoll(argv[1]) : 4096ll * 250;  // bench shape: B 4096 x L 250
  const float drop_p = argc > 2 ? (float)atof(argv[2]) : 0.1f;
  const int reps = 20;
  CK(cudaFree(nullptr));
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fp, 12000, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess) return fprintf(stderr, "no cuTensorMapEncodeTiled\n"), 1;
  EncodeTiledFn enc = (EncodeTiledFn)fp;

  float *x, *dy, *w1, *b1, *w2t, *w1t, *a, *dh, *dx;
  CK(cudaMalloc(&x, M * ffn::kD * 4));
  CK(cudaMalloc(&dy, M * ffn::kD * 4));
  CK(cudaMalloc(&dx, M * ffn::kD * 4));
  CK(cudaMalloc(&a, M * ffn::kH * 4));
  CK(cudaMalloc(&dh, M * ffn::kH * 4));
  CK(cudaMalloc(&w1, ffn::kH * ffn::kD * 4));
  CK(cudaMalloc(&w1t, ffn::kH * ffn::kD * 4));
  CK(cudaMalloc(&w2t, ffn::kH * ffn::kD * 4));
  CK(cudaMalloc(&b1, ffn::kH * 4));
  ffn::fill_kernel<<<1024, 256>>>(x, M * ffn::kD, 1, 1.0f, 1);
  ffn::fill_kernel<<<1024, 256>>>(dy, M * ffn::kD, 6, 1.0f, 1);
  ffn::fill_kernel<<<64, 256>>>(w1, ffn::kH * ffn::kD, 2, 0.088f, 1);   // ~ 1 / sqrt(128)
  ffn::fill_kernel<<<64, 256>>>(w2t, ffn::kH * ffn::kD, 3, 0.044f, 1);  // ~ 1 / sqrt(512)
  ffn::fill_kernel<<<1, 256>>>(b1, ffn::kH, 4, 0.1f, 0);
  ffn::transpose_kernel<<<64, 256>>>(w1, w1t, ffn::kH, ffn::kD);
  CK(cudaMemset(a, 0xff, M * ffn::kH * 4));  // NaN canaries
  CK(cudaMemset(dh, 0xff, M * ffn::kH * 4));
  CK(cudaMemset(dx, 0xff, M * ffn::kD * 4));
  CK(cudaDeviceSynchronize());

  CUtensorMap mx, my, m1, m2t, m1t;
  if (encode2d(enc, &mx, x, ffn::kD, (unsigned long long)M) || encode2d(enc, &my, dy, ffn::kD, (unsigned long long)M) ||
      encode2d(enc, &m1, w1, ffn::kD, ffn::kH) || encode2d(enc, &m2t, w2t, ffn::kD, ffn::kH) ||
      encode2d(enc, &m1t, w1t, ffn::kH, ffn::kD))
    return fprintf(stderr, "tensor map encode failed\n"), 1;
  ffn::Params p{};
  p.M = M;
  p.tiles = (int)((M + 127) / 128);
  p.b1 = b1;
  p.a = a;
  p.dh = dh;
  p.dx = dx;
  p.drop_scale = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
  p.drop_thresh = drop_p > 0.f ? (uint32_t)((double)drop_p * 4294967296.0) : 0u;
  p.seed = 0x1234567887654321ull;
  CK(cudaFuncSetAttribute(ffn::ffn_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ffn::kSmem));
  const int ctas = p.tiles < kNumSMs ? p.tiles : kNumSMs;
  ffn::ffn_dgrad_kernel<<<ctas, ffn::kThreads, ffn::kSmem>>>(mx, my, m1, m2t, m1t, p);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());

  // ---- check: first two tiles, a middle tile and the (possibly ragged) last tile against the fp64 reference
  std::vector<long long> rows;
  for (long long r = 0; r < 256 && r < M; ++r) rows.push_back(r);
  for (long long r = (M / 2 / 128) * 128; r < (M / 2 / 128) * 128 + 128 && r < M; ++r) rows.push_back(r);
  for (long long r = ((M - 1) / 128) * 128; r < M; ++r) rows.push_back(r);
  const int W = 2 * ffn::kH + ffn::kD;
  long long* drows;
  double* dref;
  CK(cudaMalloc(&drows, rows.size() * 8));
  CK(cudaMalloc(&dref, rows.size() * W * 8));
  CK(cudaMemcpy(drows, rows.data(), rows.size() * 8, cudaMemcpyHostToDevice));
  ffn::ffn_dgrad_reference_kernel<<<(unsigned)rows.size(), 128>>>(x, dy, w1, b1, w2t, drows, dref, p.drop_scale, p.drop_thresh, p.seed);
  CK(cudaGetLastError());
  std::vector<double> ref(rows.size() * W);
  std::vector<float> got(W);
  CK(cudaMemcpy(ref.data(), dref, ref.size() * 8, cudaMemcpyDeviceToHost));
  double num[3] = {0, 0, 0}, den[3] = {0, 0, 0};
  bool nan = false;
  for (size_t i = 0; i < rows.size(); ++i) {
    CK(cudaMemcpy(got.data(), a + rows[i] * ffn::kH, ffn::kH * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(got.data() + ffn::kH, dh + rows[i] * ffn::kH, ffn::kH * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(got.data() + 2 * ffn::kH, dx + rows[i] * ffn::kD, ffn::kD * 4, cudaMemcpyDeviceToHost));
    for (int o = 0; o < W; ++o) {
      const int part = o < ffn::kH ? 0 : o < 2 * ffn::kH ? 1 : 2;
      const double d = (double)got[o] - ref[i * W + o];
      if (std::isnan(d)) nan = true;
      num[part] += d * d;
      den[part] += ref[i * W + o] * ref[i * W + o];
    }
  }
  double rel[3];
  bool ok = !nan;
  for (int k = 0; k < 3; ++k) {
    rel[k] = sqrt(num[k] / (den[k] + 1e-300));
    ok = ok && rel[k] < 1e-3;
  }
  printf("rows %lld  drop_p %.2f  checked %zu rows: rel L2 error A %.3e  dH %.3e  dX %.3e%s  -> %s\n", M, drop_p, rows.size(), rel[0],
         rel[1], rel[2], nan ? "  (NaN: unwritten output)" : "", ok ? "PARITY OK" : "PARITY FAILED");

  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) ffn::ffn_dgrad_kernel<<<ctas, ffn::kThreads, ffn::kSmem>>>(mx, my, m1, m2t, m1t, p);
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) ffn::ffn_dgrad_kernel<<<ctas, ffn::kThreads, ffn::kSmem>>>(mx, my, m1, m2t, m1t, p);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= reps;
  const double flop = 6.0 * (double)M * ffn::kD * ffn::kH, bytes = 4.0 * (double)M * (3 * ffn::kD + 2 * ffn::kH);
  printf("%.3f ms per launch: %.1f TFLOP/s (tf32), %.1f GB/s algorithmic; unfused today: dgrad 0.55 + act-backward 1.0 + dgrad 0.55 ms\n",
         ms, flop / ms / 1e9, bytes / ms / 1e6);
  return ok ? 0 : 2;
}
