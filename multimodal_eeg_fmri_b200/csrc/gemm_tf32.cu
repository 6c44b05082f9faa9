// Host side of the tcgen05 TF32 GEMM engine + the C-ABI entry points built on it:
// nn.Linear fwd/dgrad/wgrad, dense Conv1d fwd/dgrad/wgrad (implicit GEMM over taps),
// similarity / InfoNCE contractions.  See gemm_engine.cuh for the kernel.
#include "gemm_engine.cuh"

#include <string.h>

namespace xm {

int g_last_cuda_error = 0;
int g_conv_halo = 1;  // conv wgrad: one halo tile per k-block instead of one shifted copy per tap (xm_debug_set_conv_halo)

// ---------------------------------------------------------------- TMA descriptor encode
// cuTensorMapEncodeTiled is resolved through the runtime so the library has no link-time
// dependency on libcuda (it must load on a CPU-only box for the ABI export test).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
    else (void)cudaGetLastError();
  }
  return fn;
}

int encode_tmap(CUtensorMap* out, const TensorView3& t, unsigned box0, unsigned box1, int mn_major, unsigned box2) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return XM_ERR_NO_DRIVER;
  // The driver entry point needs a current context on THIS thread.  The runtime binds the primary
  // context lazily, and a fresh thread (e.g. PyTorch's autograd worker, whose first op may be one of
  // ours) has none yet: a no-op runtime call binds it (CUDA_ERROR_INVALID_CONTEXT otherwise).
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    (void)cudaFree(nullptr);
    ctx_bound = true;
  }
  if ((reinterpret_cast<uintptr_t>(t.ptr) & 15) != 0) return XM_ERR_INVALID;
  if ((t.stride_bytes[0] & 15) != 0 || (t.stride_bytes[1] & 15) != 0) return XM_ERR_INVALID;
  cuuint64_t dims[3] = {t.dim[0], t.dim[1], t.dim[2]};
  cuuint64_t strides[2] = {t.stride_bytes[0], t.stride_bytes[1]};
  cuuint32_t box[3] = {box0, box1, box2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(t.ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    g_last_cuda_error = 100000 + (int)r;
    return XM_ERR_LAUNCH;
  }
  return XM_OK;
}

template <int EPI>
static int launch_epi(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const PeerMaps& pm,
                      const GemmParams& p, int ctas, size_t smem,
                      cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(gemm_tf32_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    g_last_cuda_error = (int)e;
    return XM_ERR_LAUNCH;
  }
  gemm_tf32_kernel<EPI><<<ctas, gemm_threads(EpiWarps<EPI>::value), smem, stream>>>(ma, mb, mc, pm, p);
  return check_launch();
}

int launch_gemm(int epi, const TensorView3& ta, const TensorView3& tb, const TensorView3& tc, GemmParams& p, dim3 grid,
                cudaStream_t stream, const void* const* b_peers) {
  if (p.n_stride < 0) p.n_stride = p.bn;
  if (p.n_peers < 0 || p.n_peers > kMaxPeers || (p.n_peers > 0 && (!b_peers || p.peer_rows <= 0 || p.taps_n != 1)))
    return XM_ERR_INVALID;
  // a tile's row range must not straddle two shards
  if (p.n_peers > 0 && (p.peer_rows % (p.b.mn_major ? 32 : p.bn))) return XM_ERR_UNSUPPORTED;
  if (p.c_col_mul < 0) p.c_col_mul = p.bn;
  if (p.bn < 16 || p.bn > 256 || (p.bn & 15)) return XM_ERR_UNSUPPORTED;
  if (p.b.mn_major && (p.bn & 31)) return XM_ERR_UNSUPPORTED;
  if (p.taps_n * p.bn > 512) return XM_ERR_UNSUPPORTED;
  if (grid.x == 0 || grid.y == 0 || grid.z == 0) return XM_OK;
  const long long ntiles = (long long)grid.x * grid.y * grid.z;
  if (ntiles > 2000000000ll) return XM_ERR_UNSUPPORTED;
  p.nx = (int)grid.x;
  p.ny = (int)grid.y;
  p.nz = (int)grid.z;
  // TMA-store epilogue whenever the output is addressable by a tensor map (16-B aligned base and pitches)
  p.tma_store = 0;
  if (epi != EPI_LSE && tc.ptr != nullptr && (reinterpret_cast<uintptr_t>(tc.ptr) & 15) == 0 &&
      (tc.stride_bytes[0] & 15) == 0 && (tc.stride_bytes[1] & 15) == 0 && !((p.bn & 31) && grid.y > 1))
    p.tma_store = 1;
  if (!p.tma_store && epi != EPI_LSE && p.c == nullptr) return XM_ERR_INVALID;
  if (p.colstat_part != nullptr && (!p.tma_store || epi != EPI_ROWMAJOR || p.taps_n != 1 || (p.bn & 63) ||
                                    (int)grid.y * p.bn > 256))  // each epilogue warp keeps <= 4 (N tile, 32-column chunk) sums
    return XM_ERR_UNSUPPORTED;
  p.b_halo = 0;
  p.b_blk_bytes = 0;
  if (p.taps_n > 1 && p.b.mn_major && p.n_peers == 0 && g_conv_halo) {
    p.b_halo = p.taps_n - 1;
    p.b_blk_bytes = ((32 + p.b_halo) * 128 + 1023) / 1024 * 1024;
  }
  p.a_halo = 0;
  p.a_slab_bytes = 0;
  if (p.taps_k > 1 && !p.a.mn_major && !p.b.mn_major && p.n_peers == 0 && p.taps_n == 1 && g_conv_halo &&
      (p.a.tap_step[1] == 1 || p.a.tap_step[1] == -1) && p.a.tap_step[0] == 0 && p.a.tap_step[2] == 0) {
    // conv fwd / dgrad: tap t reads rows base + t*dir; the slab starts at the lowest of them
    p.a_halo = p.taps_k - 1;
    p.a_slab_bytes = ((128 + p.a_halo) * 128 + 1023) / 1024 * 1024;
    p.a_tap_dir = p.a.tap_step[1];
    p.a_halo_row_shift = p.a_tap_dir > 0 ? 0 : -p.a_halo;  // dir = -1: rows base - (taps-1) ... base
    p.a_tap_row0 = p.a_tap_dir > 0 ? 0 : p.a_halo;
  }
  int stage_bytes = p.b_halo ? kATileBytes + (p.bn >> 5) * p.b_blk_bytes : kATileBytes + p.taps_n * p.bn * 128;
  if (p.a_halo) {
    stage_bytes = p.a_slab_bytes + p.taps_k * p.bn * 128;
    const int stg = p.tma_store ? staging_bytes(8) : 0;
    if ((222 * 1024 - 1024 - stg) / stage_bytes < 2) {  // the weight tiles of all taps do not fit twice: per-tap stages
      p.a_halo = 0;
      p.a_slab_bytes = 0;
      stage_bytes = kATileBytes + p.bn * 128;
    }
  }
  if (p.a_wrap > 0 && !p.a_halo) return XM_ERR_UNSUPPORTED;  // the wrap lives in the halo producer
  const int total_kb = p.kout_count * p.taps_k * p.kin_count;
  if (total_kb <= 0) return XM_ERR_INVALID;
  const int epi_warps = epi == EPI_LSE ? 4 : 8;
  const int staging = p.tma_store ? staging_bytes(epi_warps) : 0;
  int stages = (222 * 1024 - 1024 - staging) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return XM_ERR_UNSUPPORTED;
  p.stages = stages;
  const int acc_cols = p.taps_n * p.bn;
  p.acc_bufs = (2 * acc_cols <= 512 && ntiles > 1) ? 2 : 1;
  p.tmem_cols = tmem_cols_for(p.acc_bufs * acc_cols);
  // chunked fp32 accumulation: plain row-major tiles stored by TMA, two accumulator sets + the sum region
  if (p.acc_chunk > 0 && (epi != EPI_ROWMAJOR || !p.tma_store || p.taps_n != 1 || p.a_halo || p.b_halo ||
                          (p.bn & 31) || 3 * p.bn > 512 || total_kb <= p.acc_chunk))
    p.acc_chunk = 0;
  if (p.acc_chunk > 0) {
    p.acc_bufs = 2;
    p.tmem_cols = tmem_cols_for(3 * p.bn);
  }
  p.a.rows = 128;
  p.b.rows = p.bn;
  const size_t smem = (size_t)stages * stage_bytes + staging + 1024;

  CUtensorMap ma, mb, mc;
  memset(&mc, 0, sizeof(mc));
  int rc = encode_tmap(&ma, ta, 32, p.a.mn_major ? 32 : (unsigned)(128 + p.a_halo), p.a.mn_major);
  if (rc != XM_OK) return rc;
  if (p.b_rows_dim2) {  // MN-major B whose contraction rows run along tensor dimension 2: box {32, 1, 32}
    if (!p.b.mn_major || p.b_halo) return XM_ERR_INVALID;
    rc = encode_tmap(&mb, tb, 32, 1, 1, 32);
  } else {
    rc = encode_tmap(&mb, tb, 32, p.b.mn_major ? (unsigned)(32 + p.b_halo) : (unsigned)p.bn, p.b.mn_major);
  }
  if (rc != XM_OK) return rc;
  if (p.tma_store) {
    rc = encode_tmap(&mc, tc, 32, 32, 0);
    if (rc != XM_OK) return rc;
  }
  static PeerMaps pm_zero;  // zero-initialised
  PeerMaps pm = pm_zero;
  for (int r = 0; r < p.n_peers; ++r) {  // same geometry as tb, one base pointer per rank
    TensorView3 tv = tb;
    tv.ptr = b_peers[r];
    tv.dim[1] = (unsigned long long)p.peer_rows;
    rc = encode_tmap(&pm.m[r], tv, 32, p.b.mn_major ? 32 : (unsigned)p.bn, p.b.mn_major);
    if (rc != XM_OK) return rc;
  }
  const int ctas = (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
  switch (epi) {
    case EPI_ROWMAJOR: return launch_epi<EPI_ROWMAJOR>(ma, mb, mc, pm, p, ctas, smem, stream);
    case EPI_LSE: return launch_epi<EPI_LSE>(ma, mb, mc, pm, p, ctas, smem, stream);
    case EPI_NCE_GRAD: return launch_epi<EPI_NCE_GRAD>(ma, mb, mc, pm, p, ctas, smem, stream);
  }
  return XM_ERR_INVALID;
}

static void zero_params(GemmParams& p) {
  memset(&p, 0, sizeof(p));
  p.taps_k = 1;
  p.taps_n = 1;
  p.kout_count = 1;
  p.kout_total = 1;
  p.alpha = 1.0f;
  p.n_stride = -1;   // launch_gemm: bn
  p.c_col_mul = -1;  // launch_gemm: bn
}

// N tile: as wide as the problem allows (fewer A re-reads), narrowed while the grid
// would leave most of the 148 SMs idle.
static int choose_bn(int64_t N, int64_t other_tiles, int gran) {
  int bn = (int)(((N + gran - 1) / gran) * gran);
  if (bn > 256) bn = 256;
  while (bn >= 2 * 32 && (bn / 2) % gran == 0 && other_tiles * ((N + bn - 1) / bn) < kNumSMs) bn /= 2;
  return bn;
}

// ---------------------------------------------------------------- split-K / wgrad reductions
// out[m, n] = act(sum_s ws[s, m, n] + bias[n])
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int splits, long long M, long long N,
                                     const float* __restrict__ bias, float* __restrict__ out, long long ldo, int act,
                                     int round_out) {
  const long long total = M * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / N, n = i - m * N;
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += ws[(long long)s * total + i];
    if (bias) acc += bias[n];
    acc = apply_act(acc, act);
    if (round_out) acc = round_tf32(acc);
    out[m * ldo + n] = acc;
  }
}

// dw[co, ci, tap] = sum_s ws[s, tap, co, ci]   (ws slabs are (taps, Cout, bn) with pitch bn)
__global__ void conv_wgrad_reduce_kernel(const float* __restrict__ ws, int splits, int taps, int Cout, int Cin, int bn,
                                         float* __restrict__ dw) {
  const long long total = (long long)Cout * Cin * taps;
  const long long slab = (long long)taps * Cout * bn;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % taps);
    const long long r = i / taps;
    const int ci = (int)(r % Cin);
    const int co = (int)(r / Cin);
    const long long src = ((long long)tap * Cout + co) * bn + ci;
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += ws[(long long)s * slab + src];
    dw[i] = acc;
  }
}

// w (Cout, Cin, taps) -> wk (taps, Cout, ldk) and wt (taps, Cin, ldt), tf32-rounded, zero padded
__global__ void conv_pack_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, float* __restrict__ wk,
                                        int ldk, float* __restrict__ wt, int ldt) {
  const long long nk = (long long)taps * Cout * ldk;
  const long long nt = (long long)taps * Cin * ldt;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nk + nt; i += (long long)gridDim.x * blockDim.x) {
    if (i < nk) {
      const int ci = (int)(i % ldk);
      const long long r = i / ldk;
      const int co = (int)(r % Cout);
      const int tap = (int)(r / Cout);
      wk[i] = ci < Cin ? round_tf32(w[((long long)co * Cin + ci) * taps + tap]) : 0.f;
    } else {
      const long long j = i - nk;
      const int co = (int)(j % ldt);
      const long long r = j / ldt;
      const int ci = (int)(r % Cin);
      const int tap = (int)(r / Cin);
      wt[j] = co < Cout ? round_tf32(w[((long long)co * Cin + ci) * taps + tap]) : 0.f;
    }
  }
}

// lse[m] = shift + log(sum_t partial[t, m])
__global__ void lse_finalize_kernel(const float* __restrict__ partial, int ntiles, long long M, float shift,
                                    float* __restrict__ lse) {
  const long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (m >= M) return;
  float acc = 0.f;
  for (int t = 0; t < ntiles; ++t) acc += partial[(long long)t * M + m];
  lse[m] = shift + logf(acc);
}

static int grid_for(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  const long long cap = (long long)kNumSMs * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}


}  // namespace xm

using namespace xm;

// =================================================================== C ABI
extern "C" {

int xm_abi_version(void) { return XM_ABI_VERSION; }
int xm_last_cuda_error(void) { return xm::g_last_cuda_error; }
const char* xm_strerror(int code) {
  switch (code) {
    case XM_OK: return "ok";
    case XM_ERR_INVALID: return "invalid argument (shape / alignment / null pointer)";
    case XM_ERR_UNSUPPORTED: return "unsupported shape";
    case XM_ERR_LAUNCH: return "CUDA launch or driver error (see xm_last_cuda_error)";
    case XM_ERR_NO_DRIVER: return "cuTensorMapEncodeTiled unavailable (no CUDA driver)";
  }
  return "unknown error";
}

// blocks == 3: x is (3M, K) and w is (3N, K), three row-stacked tf32 split blocks each; y = sum_b x[b] w[b]^T
static int linear_fwd_impl(const float* x, const float* w, const float* bias, float* y, int64_t M, int64_t N, int64_t K,
                           int64_t ldx, int64_t ldw, int64_t ldy, int act, int flags, int splits, float* workspace,
                           int blocks, void* stream) {
  const int round_out = (flags & XM_LINEAR_ROUND_TF32) ? 1 : 0;
  const bool fp32_accum = (flags & XM_LINEAR_FP32_ACCUM) != 0;
  if (!x || !w || !y || M <= 0 || N <= 0 || K <= 0) return XM_ERR_INVALID;
  if ((ldx & 3) || (ldw & 3)) return XM_ERR_INVALID;
  if (splits < 1) splits = 1;
  if (splits > 1 && !workspace) return XM_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const int m_tiles = ceil_div(M, 128);
  GemmParams p;
  zero_params(p);
  p.bn = choose_bn(N, (int64_t)m_tiles * splits, 16);
  const int kb = ceil_div(K, 32);
  const int kb_split = ceil_div(kb, splits);
  splits = ceil_div(kb, kb_split);
  p.kin_count = kb_split;
  p.a.mn_major = 0;
  p.a.sx[1] = 128;
  p.a.kin_step[0] = 32;
  p.a.sz[0] = 32 * kb_split;
  p.b.mn_major = 0;
  p.b.sy[1] = p.bn;
  p.b.kin_step[0] = 32;
  p.b.sz[0] = 32 * kb_split;
  p.M = (int)M;
  p.N = (int)N;
  // the GEMM epilogue fuses bias + {none, relu, gelu}; tanh / sigmoid run as a second (in-place) pass
  const bool late_act = (act == XM_ACT_TANH || act == XM_ACT_SIGMOID);
  if (late_act && ldy != N) return XM_ERR_UNSUPPORTED;
  if (fp32_accum) {  // 8 k-blocks = 32 tensor-core accumulation steps per chunk: truncation bias < 1e-6
    p.acc_chunk = 8;
    if (p.bn > 128) p.bn = 128;  // three accumulator-sized TMEM regions must fit 512 columns
    if (p.bn & 31) p.bn = (p.bn + 31) & ~31;
    p.b.sy[1] = p.bn;
  }
  if (splits == 1) {
    p.c = y;
    p.ldc = ldy;
    p.bias = bias;
    p.act = late_act ? XM_ACT_NONE : act;
    p.round_tf32 = late_act ? 0 : round_out;
  } else {
    p.c = workspace;
    p.ldc = N;
    p.c_z_stride = (long long)M * N;
  }
  TensorView3 ta{x, {(unsigned long long)K, (unsigned long long)M, (unsigned long long)blocks}, {(unsigned long long)ldx * 4, (unsigned long long)M * ldx * 4}};
  TensorView3 tb{w, {(unsigned long long)K, (unsigned long long)N, (unsigned long long)blocks}, {(unsigned long long)ldw * 4, (unsigned long long)N * ldw * 4}};
  if (blocks > 1) {  // the block index is the outer contraction loop: tensor-map dimension 2 of both operands
    p.kout_count = blocks;
    p.kout_total = blocks;
    p.a.kout_step[2] = 1;
    p.b.kout_step[2] = 1;
  }
  p.c_z_mul = 1;
  const TensorView3 tc = splits == 1 ? TensorView3{y, {(unsigned long long)(N), (unsigned long long)(M), (unsigned long long)(1)}, {(unsigned long long)(ldy) * 4, (unsigned long long)(M * ldy) * 4}}
                                     : TensorView3{workspace, {(unsigned long long)(N), (unsigned long long)(M), (unsigned long long)(splits)}, {(unsigned long long)(N) * 4, (unsigned long long)(M * N) * 4}};
  int rc = launch_gemm(EPI_ROWMAJOR, ta, tb, tc, p, dim3(m_tiles, ceil_div(N, p.bn), splits), st);
  if (rc != XM_OK) return rc;
  if (splits > 1) {
    splitk_reduce_kernel<<<grid_for(M * N, 256), 256, 0, st>>>(workspace, splits, M, N, bias, y, ldy, act, round_out);
    return check_launch();
  }
  if (late_act) {
    rc = xm_act_fwd_f32(y, y, M * N, act, 0.f, 0, round_out, stream);
  }
  return rc;
}

int xm_linear_fwd_f32(const float* x, const float* w, const float* bias, float* y, int64_t M, int64_t N, int64_t K,
                      int64_t ldx, int64_t ldw, int64_t ldy, int act, int flags, int splits, float* workspace,
                      void* stream) {
  return linear_fwd_impl(x, w, bias, y, M, N, K, ldx, ldw, ldy, act, flags, splits, workspace, 1, stream);
}

int xm_linear_fwd_stacked3_f32(const float* x3, const float* w3, const float* bias, float* y, int64_t M, int64_t N, int64_t K,
                               int64_t ldx, int64_t ldw, int64_t ldy, int act, int flags, int splits, float* workspace,
                               void* stream) {
  return linear_fwd_impl(x3, w3, bias, y, M, N, K, ldx, ldw, ldy, act, flags, splits, workspace, 3, stream);
}

static int linear_dgrad_impl(const float* dy, const float* w, const void* const* w_peers, int n_peers, int64_t peer_rows,
                             float* dx, int64_t M, int64_t N, int64_t K, int64_t lddy, int64_t ldw, int64_t lddx,
                             int round_out, void* stream) {
  if (!dy || (!w && !w_peers) || !dx || M <= 0 || N <= 0 || K <= 0) return XM_ERR_INVALID;
  if ((lddy & 3) || (ldw & 3)) return XM_ERR_INVALID;
  const int m_tiles = ceil_div(M, 128);
  GemmParams p;
  zero_params(p);
  p.bn = choose_bn(K, m_tiles, 32);
  p.kin_count = ceil_div(N, 32);
  p.a.mn_major = 0;  // dy (M, N): contraction index N is contiguous
  p.a.sx[1] = 128;
  p.a.kin_step[0] = 32;
  p.b.mn_major = 1;  // w (N, K): output index K is contiguous
  p.b.sy[0] = p.bn;
  p.b.kin_step[1] = 32;
  p.M = (int)M;
  p.N = (int)K;
  p.c = dx;
  p.ldc = lddx;
  p.round_tf32 = round_out;
  p.n_peers = n_peers;
  p.peer_rows = (int)peer_rows;
  TensorView3 ta{dy, {(unsigned long long)N, (unsigned long long)M, 1}, {(unsigned long long)lddy * 4, (unsigned long long)M * lddy * 4}};
  TensorView3 tb{w ? (const void*)w : w_peers[0], {(unsigned long long)K, (unsigned long long)N, 1}, {(unsigned long long)ldw * 4, (unsigned long long)N * ldw * 4}};
  const TensorView3 tc = TensorView3{dx, {(unsigned long long)(K), (unsigned long long)(M), (unsigned long long)(1)}, {(unsigned long long)(lddx) * 4, (unsigned long long)(M * lddx) * 4}};
  return launch_gemm(EPI_ROWMAJOR, ta, tb, tc, p, dim3(m_tiles, ceil_div(K, p.bn), 1), (cudaStream_t)stream, w_peers);
}

int xm_linear_dgrad_f32(const float* dy, const float* w, float* dx, int64_t M, int64_t N, int64_t K, int64_t lddy,
                        int64_t ldw, int64_t lddx, int round_out, void* stream) {
  return linear_dgrad_impl(dy, w, nullptr, 0, 0, dx, M, N, K, lddy, ldw, lddx, round_out, stream);
}

int xm_linear_dgrad_peers_f32(const float* dy, const void* const* w_peers, int n_peers, int64_t rows_per_peer, float* dx,
                              int64_t M, int64_t K, int64_t lddy, int64_t ldw, int64_t lddx, int round_out, void* stream) {
  if (n_peers <= 0 || rows_per_peer <= 0) return XM_ERR_INVALID;
  return linear_dgrad_impl(dy, nullptr, w_peers, n_peers, rows_per_peer, dx, M, n_peers * rows_per_peer, K, lddy, ldw, lddx,
                           round_out, stream);
}

int xm_linear_wgrad_f32(const float* dy, const float* x, float* dw, float* db, int64_t M, int64_t N, int64_t K,
                        int64_t lddy, int64_t ldx, int64_t lddw, int splits, float* workspace, float* db_workspace,
                        void* stream) {
  if (!dy || !x || !dw || M <= 0 || N <= 0 || K <= 0) return XM_ERR_INVALID;
  if ((lddy & 3) || (ldx & 3)) return XM_ERR_INVALID;
  if (splits < 1) splits = 1;
  if (splits > 1 && !workspace) return XM_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const int m_tiles = ceil_div(N, 128);  // GEMM rows = output features
  GemmParams p;
  zero_params(p);
  p.bn = choose_bn(K, (int64_t)m_tiles * splits, 32);
  const int kb = ceil_div(M, 32);
  const int kb_split = ceil_div(kb, splits);
  splits = ceil_div(kb, kb_split);
  p.kin_count = kb_split;
  p.a.mn_major = 1;  // dy (M, N): GEMM row index N contiguous, contraction index M strided
  p.a.sx[0] = 128;
  p.a.kin_step[1] = 32;
  p.a.sz[1] = 32 * kb_split;
  p.b.mn_major = 1;  // x (M, K): GEMM col index K contiguous
  p.b.sy[0] = p.bn;
  p.b.kin_step[1] = 32;
  p.b.sz[1] = 32 * kb_split;
  p.M = (int)N;
  p.N = (int)K;
  if (splits == 1) {
    p.c = dw;
    p.ldc = lddw;
  } else {
    p.c = workspace;
    p.ldc = K;
    p.c_z_stride = (long long)N * K;
  }
  TensorView3 ta{dy, {(unsigned long long)N, (unsigned long long)M, 1}, {(unsigned long long)lddy * 4, (unsigned long long)M * lddy * 4}};
  TensorView3 tb{x, {(unsigned long long)K, (unsigned long long)M, 1}, {(unsigned long long)ldx * 4, (unsigned long long)M * ldx * 4}};
  p.c_z_mul = 1;
  const TensorView3 tc = splits == 1 ? TensorView3{dw, {(unsigned long long)(K), (unsigned long long)(N), (unsigned long long)(1)}, {(unsigned long long)(lddw) * 4, (unsigned long long)(N * lddw) * 4}}
                                     : TensorView3{workspace, {(unsigned long long)(K), (unsigned long long)(N), (unsigned long long)(splits)}, {(unsigned long long)(K) * 4, (unsigned long long)(N * K) * 4}};
  int rc = launch_gemm(EPI_ROWMAJOR, ta, tb, tc, p, dim3(m_tiles, ceil_div(K, p.bn), splits), st);
  if (rc != XM_OK) return rc;
  if (splits > 1) {
    splitk_reduce_kernel<<<grid_for(N * K, 256), 256, 0, st>>>(workspace, splits, N, K, nullptr, dw, lddw, XM_ACT_NONE, 0);
    rc = check_launch();
    if (rc != XM_OK) return rc;
  }
  if (db) return xm_colsum_f32(dy, M, N, lddy, db, db_workspace, stream);
  return XM_OK;
}

// ------------------------------------------------------------------ conv1d
int xm_conv1d_pack_weight_f32(const float* w, int64_t Cout, int64_t Cin, int64_t taps, float* wk, int64_t ldk,
                              float* wt, int64_t ldt, void* stream) {
  if (!w || !wk || !wt || ldk < Cin || ldt < Cout || (ldk & 3) || (ldt & 3)) return XM_ERR_INVALID;
  const long long n = taps * Cout * ldk + taps * Cin * ldt;
  conv_pack_weight_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(w, (int)Cout, (int)Cin, (int)taps, wk,
                                                                              (int)ldk, wt, (int)ldt);
  return check_launch();
}

// Channels-last activations: x (B, T, C) with row pitch ld (one row = one time step, C contiguous).
// TMA needs 16-byte aligned INNER coordinates, so a conv tap cannot be a shift along a contiguous
// time axis; with channels innermost the tap is a shift of the ROW coordinate, which is free, and
// TMA zero-fills rows outside [0, T) of each sample -- that is the conv zero padding.
//
// Shared by fwd (dir = +1) and dgrad (dir = -1):
//   out[b, t, n] = sum_tap sum_k in[b, t + dir*(tap - pad), k] * wp[tap, n, k]  (+ bias[n])
static int conv_like(const float* in, const float* wp, const float* bias, float* out, int64_t B, int64_t Kc, int64_t Nc,
                     int64_t T, int64_t taps, int64_t ld_in, int64_t ld_w, int64_t ld_out, int dir, int round_out,
                     cudaStream_t st, double* stat_part = nullptr, int64_t in_cols = 0) {
  // in_cols (0 = Kc): physical channel count of `in` when the tensor stores fewer channel blocks than the contraction
  // walks -- [hi | lo] for a 3-pass conv over [hi | lo | hi]: block coordinates >= in_cols wrap back by in_cols
  if (in_cols <= 0) in_cols = Kc;
  if (in_cols != Kc && (in_cols > Kc || (in_cols & 31) || Kc > 2 * in_cols || taps == 1)) return XM_ERR_UNSUPPORTED;
  if (!in || !wp || !out || B <= 0 || Kc <= 0 || Nc <= 0 || T <= 0 || taps <= 0 || !(taps & 1)) return XM_ERR_INVALID;
  if ((ld_in & 3) || (ld_w & 3) || ld_in < in_cols || ld_out < Nc || B > 65535) return XM_ERR_INVALID;
  const int pad = (int)(taps / 2);
  const int t_tiles = ceil_div(T, 128);
  GemmParams p;
  zero_params(p);
  p.bn = choose_bn(Nc, (int64_t)t_tiles * B, 16);
  // halo staging (launch_gemm) needs two stages of {activation slab + taps weight tiles}: narrow the N tile
  // if that is what it takes (the slab is then re-read per N tile, still far less than once per tap)
  while (taps > 1 && p.bn > 64 && (p.bn / 2) % 16 == 0 &&
         2 * ((((128 + (int)taps - 1) * 128 + 1023) / 1024 * 1024) + (int)taps * p.bn * 128) > (222 - 1 - 64) * 1024)
    p.bn /= 2;
  p.taps_k = (int)taps;
  p.kin_count = ceil_div(Kc, 32);
  p.a.mn_major = 0;  // activations: contraction index (channel) contiguous
  p.a.base[1] = -dir * pad;
  p.a.sx[1] = 128;
  p.a.sz[2] = 1;
  p.a.kin_step[0] = 32;
  p.a.tap_step[1] = dir;
  p.b.mn_major = 0;  // packed weights (taps, N, K): contraction index contiguous
  p.b.sy[1] = p.bn;
  p.b.kin_step[0] = 32;
  p.b.tap_step[2] = 1;
  p.M = (int)T;
  p.N = (int)Nc;
  p.c = out;
  p.ldc = ld_out;
  p.c_z_stride = (long long)T * ld_out;
  p.bias = bias;
  p.round_tf32 = round_out;
  p.a_wrap = in_cols != Kc ? (int)in_cols : 0;
  TensorView3 ta{in, {(unsigned long long)in_cols, (unsigned long long)T, (unsigned long long)B},
                 {(unsigned long long)ld_in * 4, (unsigned long long)T * ld_in * 4}};
  TensorView3 tb{wp, {(unsigned long long)Kc, (unsigned long long)Nc, (unsigned long long)taps},
                 {(unsigned long long)ld_w * 4, (unsigned long long)Nc * ld_w * 4}};
  p.c_z_mul = 1;
  const TensorView3 tc = TensorView3{out, {(unsigned long long)(Nc), (unsigned long long)(T), (unsigned long long)(B)}, {(unsigned long long)(ld_out) * 4, (unsigned long long)(T * ld_out) * 4}};
  if (stat_part != nullptr) {
    // the statistics epilogue covers single-N-tile launches whose tiles leave through the TMA store
    if ((reinterpret_cast<uintptr_t>(out) & 15) || (ld_out & 3)) return XM_ERR_UNSUPPORTED;
    if (cudaMemsetAsync(stat_part, 0, (size_t)xm_conv1d_fwd_stat_rows() * Nc * 2 * sizeof(double), st) != cudaSuccess) {
      g_last_cuda_error = (int)cudaGetLastError();
      return XM_ERR_LAUNCH;
    }
    p.colstat_part = stat_part;
  }
  return launch_gemm(EPI_ROWMAJOR, ta, tb, tc, p, dim3(t_tiles, ceil_div(Nc, p.bn), (unsigned)B), st);
}

int xm_conv1d_fwd_stat_rows(void) { return kNumSMs * 4; }

int xm_conv1d_fwd_stats_f32(const float* x, const float* wk, const float* bias, float* y, double* stat_part, int64_t B,
                            int64_t Cin, int64_t Cout, int64_t T, int64_t taps, int64_t ldx, int64_t ldk, int64_t ldy,
                            int round_out, int64_t x_cols, void* stream) {
  return conv_like(x, wk, bias, y, B, Cin, Cout, T, taps, ldx, ldk, ldy, +1, round_out, (cudaStream_t)stream, stat_part, x_cols);
}

int xm_conv1d_fwd_f32(const float* x, const float* wk, const float* bias, float* y, int64_t B, int64_t Cin,
                      int64_t Cout, int64_t T, int64_t taps, int64_t ldx, int64_t ldk, int64_t ldy, int round_out,
                      void* stream) {
  return conv_like(x, wk, bias, y, B, Cin, Cout, T, taps, ldx, ldk, ldy, +1, round_out, (cudaStream_t)stream);
}

int xm_conv1d_dgrad_f32(const float* dy, const float* wt, float* dx, int64_t B, int64_t Cin, int64_t Cout, int64_t T,
                        int64_t taps, int64_t lddy, int64_t ldt, int64_t lddx, int round_out, void* stream) {
  return conv_like(dy, wt, nullptr, dx, B, Cout, Cin, T, taps, lddy, ldt, lddx, -1, round_out, (cudaStream_t)stream);
}

// wgrad:  dw[tap][co][ci] = sum_b sum_t dy[b, t, co] * x[b, t + tap - pad, ci]
// GEMM rows = co, cols = ci, contraction = (b, t): both operands are MN-major (channel contiguous,
// one row per t).  Samples are the outer contraction index, split across gridDim.z; taps are
// separate TMEM accumulators (as many as fit 512 columns), remaining tap groups are extra launches.
struct WgradPlan {
  int bn, taps_n, tap_groups, samples_per_cta, splits;
};
static WgradPlan wgrad_plan(int64_t B, int64_t Cin, int64_t Cout, int64_t taps) {
  WgradPlan pl;
  pl.bn = (int)((Cin + 31) / 32 * 32);
  int per = 512 / pl.bn;
  if (per < 1) per = 1;
  // keep >= 2 pipeline stages in shared memory: 16 KB + taps_n * bn * 128 B per stage
  while (per > 1 && 2 * (kATileBytes + per * pl.bn * 128) > 220 * 1024) --per;
  pl.taps_n = (int)(taps < per ? taps : per);
  pl.tap_groups = (int)((taps + pl.taps_n - 1) / pl.taps_n);
  const int m_tiles = (int)((Cout + 127) / 128);
  int want = kNumSMs / (m_tiles * pl.tap_groups);
  if (want < 1) want = 1;
  if (want > B) want = (int)B;
  pl.samples_per_cta = (int)((B + want - 1) / want);
  pl.splits = (int)((B + pl.samples_per_cta - 1) / pl.samples_per_cta);
  return pl;
}

int64_t xm_conv1d_wgrad_workspace(int64_t B, int64_t Cin, int64_t Cout, int64_t taps) {
  WgradPlan pl = wgrad_plan(B, Cin, Cout, taps);
  return (int64_t)pl.splits * taps * Cout * pl.bn;
}

int xm_conv1d_wgrad_f32(const float* dy, const float* x, float* dw, float* db, int64_t B, int64_t Cin, int64_t Cout,
                        int64_t T, int64_t taps, int64_t lddy, int64_t ldx, float* workspace, float* db_workspace,
                        void* stream) {
  if (!dy || !x || !dw || !workspace || B <= 0 || Cin <= 0 || Cout <= 0 || T <= 0 || taps <= 0 || !(taps & 1))
    return XM_ERR_INVALID;
  if ((lddy & 3) || (ldx & 3) || Cin > 256 || lddy < Cout || ldx < Cin) return XM_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const int pad = (int)(taps / 2);
  WgradPlan pl = wgrad_plan(B, Cin, Cout, taps);
  const int m_tiles = ceil_div(Cout, 128);
  for (int g = 0; g < pl.tap_groups; ++g) {
    const int tap0 = g * pl.taps_n;
    const int ntap = (int)(taps - tap0 < pl.taps_n ? taps - tap0 : pl.taps_n);
    GemmParams p;
    zero_params(p);
    p.bn = pl.bn;
    p.taps_n = ntap;
    p.kin_count = ceil_div(T, 32);
    p.kout_count = pl.samples_per_cta;
    p.kout_total = (int)B;
    p.kout_split = 1;
    p.a.mn_major = 1;  // dy (B, T, Cout): GEMM row index co contiguous, one row per t
    p.a.sx[0] = 128;
    p.a.kin_step[1] = 32;
    p.a.kout_step[2] = 1;
    p.b.mn_major = 1;  // x (B, T, Cin): GEMM col index ci contiguous
    p.b.base[1] = tap0 - pad;
    p.b.kin_step[1] = 32;
    p.b.kout_step[2] = 1;
    p.b.tap_step[1] = 1;
    p.M = (int)Cout;
    p.N = (int)Cin;
    p.c = workspace + (long long)tap0 * Cout * pl.bn;
    p.ldc = pl.bn;
    p.c_tap_stride = (long long)Cout * pl.bn;
    p.c_z_stride = (long long)taps * Cout * pl.bn;
    TensorView3 ta{dy, {(unsigned long long)Cout, (unsigned long long)T, (unsigned long long)B},
                   {(unsigned long long)lddy * 4, (unsigned long long)T * lddy * 4}};
    TensorView3 tb{x, {(unsigned long long)Cin, (unsigned long long)T, (unsigned long long)B},
                   {(unsigned long long)ldx * 4, (unsigned long long)T * ldx * 4}};
    const TensorView3 tc{nullptr, {0, 0, 0}, {0, 0}};  // direct-store epilogue: one tile per CTA, all smem for the ring
    int rc = launch_gemm(EPI_ROWMAJOR, ta, tb, tc, p, dim3(m_tiles, 1, pl.splits), st);
    if (rc != XM_OK) return rc;
  }
  conv_wgrad_reduce_kernel<<<grid_for(Cout * Cin * taps, 256), 256, 0, st>>>(workspace, pl.splits, (int)taps, (int)Cout,
                                                                             (int)Cin, pl.bn, dw);
  int rc = check_launch();
  if (rc != XM_OK) return rc;
  if (db) return xm_colsum_f32(dy, B * T, Cout, lddy, db, db_workspace, stream);
  return rc;
}

// ------------------------------------------------------------------ similarity / InfoNCE
static void sim_operands(GemmParams& p, int bn) {
  p.bn = bn;
  p.a.mn_major = 0;
  p.a.sx[1] = 128;
  p.a.kin_step[0] = 32;
  p.b.mn_major = 0;
  p.b.sy[1] = bn;
  p.b.kin_step[0] = 32;
}

int xm_similarity_f32(const float* a, const float* b, float* S, int64_t Ml, int64_t Ng, int64_t D, float inv_tau,
                      void* stream) {
  if (!a || !b || !S || Ml <= 0 || Ng <= 0 || D <= 0 || (D & 3)) return XM_ERR_INVALID;
  GemmParams p;
  zero_params(p);
  sim_operands(p, choose_bn(Ng, ceil_div(Ml, 128), 16));
  p.kin_count = ceil_div(D, 32);
  p.M = (int)Ml;
  p.N = (int)Ng;
  p.c = S;
  p.ldc = Ng;
  p.alpha = inv_tau;
  TensorView3 ta{a, {(unsigned long long)D, (unsigned long long)Ml, 1}, {(unsigned long long)D * 4, (unsigned long long)Ml * D * 4}};
  TensorView3 tb{b, {(unsigned long long)D, (unsigned long long)Ng, 1}, {(unsigned long long)D * 4, (unsigned long long)Ng * D * 4}};
  const TensorView3 tc = TensorView3{S, {(unsigned long long)(Ng), (unsigned long long)(Ml), (unsigned long long)(1)}, {(unsigned long long)(Ng) * 4, (unsigned long long)(Ml * Ng) * 4}};
  return launch_gemm(EPI_ROWMAJOR, ta, tb, tc, p, dim3(ceil_div(Ml, 128), ceil_div(Ng, p.bn), 1), (cudaStream_t)stream);
}

int xm_infonce_tile_n(void) { return 128; }

// dx (Ml, D) = G (Ml, Ng) @ f_n (Ng, D) with both operands 3-way tf32 split (fp32-accurate):
//   g3 (Ml, 3*Ng) = [G_a | G_b | G_c] (split3 of G along its columns), f3 (Ng, 3*D) = [F_a | F_b | F_c] (the
//   l2norm split of the unit vectors): dx = G_a F_a + G_b F_b + G_c F_c -- one GEMM whose contraction walks
//   the three (column block of g3, column block of f3) pairs: kout = block, kin = 32-row steps inside it.
// The rows of G sum to ~0 (softmax minus one-hot), so G @ f_n cancels against the common component of the
// embeddings: single-pass tf32 left percent-level errors in every encoder gradient at batch 2048+.
int xm_infonce_dgrad_f32(const float* g3, const float* f3, float* dx, int64_t Ml, int64_t Ng, int64_t D, void* stream) {
  if (!g3 || !f3 || !dx || Ml <= 0 || Ng <= 0 || D <= 0 || (D & 3) || (Ng & 3)) return XM_ERR_INVALID;
  const int m_tiles = ceil_div(Ml, 128);
  GemmParams p;
  zero_params(p);
  p.bn = choose_bn(D, m_tiles, 32);
  p.kin_count = ceil_div(Ng, 32);
  p.kout_count = 3;
  p.kout_total = 3;
  p.a.mn_major = 0;  // g3: contraction (columns) contiguous
  p.a.sx[1] = 128;
  p.a.kin_step[0] = 32;
  p.a.kout_step[0] = (int)Ng;
  p.b.mn_major = 1;  // f3 viewed as (D, 3 blocks, Ng rows): output index D contiguous, rows along dimension 2
  p.b.sy[0] = p.bn;
  p.b.kin_step[2] = 32;
  p.b.kout_step[1] = 1;
  p.b_rows_dim2 = 1;
  p.acc_chunk = 8;  // the contraction runs over 3x the GLOBAL batch: keep the accumulation fp32-accurate
  if (p.bn > 128) p.bn = 128, p.b.sy[0] = 128;
  p.M = (int)Ml;
  p.N = (int)D;
  p.c = dx;
  p.ldc = D;
  TensorView3 ta{g3, {(unsigned long long)(3 * Ng), (unsigned long long)Ml, 1}, {(unsigned long long)(3 * Ng) * 4, (unsigned long long)(Ml * 3 * Ng) * 4}};
  TensorView3 tb{f3, {(unsigned long long)D, 3, (unsigned long long)Ng}, {(unsigned long long)D * 4, (unsigned long long)(3 * D) * 4}};
  const TensorView3 tc{dx, {(unsigned long long)D, (unsigned long long)Ml, 1}, {(unsigned long long)D * 4, (unsigned long long)(Ml * D) * 4}};
  return launch_gemm(EPI_ROWMAJOR, ta, tb, tc, p, dim3(m_tiles, ceil_div(D, p.bn), 1), (cudaStream_t)stream);
}

static int infonce_lse_impl(const float* a, const float* b, const void* const* b_peers, int n_peers, int64_t peer_rows,
                            float* lse, float* diag, int64_t Ml, int64_t Ng, int64_t D, float inv_tau, int64_t diag_off,
                            float* workspace, void* stream) {
  if (!a || (!b && !b_peers) || !lse || !diag || !workspace || Ml <= 0 || Ng <= 0 || D <= 0 || (D & 3)) return XM_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  GemmParams p;
  zero_params(p);
  sim_operands(p, 128);
  p.kin_count = ceil_div(D, 32);
  p.M = (int)Ml;
  p.N = (int)Ng;
  p.alpha = inv_tau;
  p.shift = inv_tau;  // |cos| <= 1  =>  S <= inv_tau: a fixed shift replaces the running max
  p.partial = workspace;
  p.diag = diag;
  p.diag_off = (int)diag_off;
  p.n_peers = n_peers;
  p.peer_rows = (int)peer_rows;
  const int ntiles = ceil_div(Ng, 128);
  TensorView3 ta{a, {(unsigned long long)D, (unsigned long long)Ml, 1}, {(unsigned long long)D * 4, (unsigned long long)Ml * D * 4}};
  TensorView3 tb{b ? (const void*)b : b_peers[0], {(unsigned long long)D, (unsigned long long)Ng, 1}, {(unsigned long long)D * 4, (unsigned long long)Ng * D * 4}};
  const TensorView3 tc{nullptr, {0, 0, 0}, {0, 0}};
  int rc = launch_gemm(EPI_LSE, ta, tb, tc, p, dim3(ceil_div(Ml, 128), ntiles, 1), st, b_peers);
  if (rc != XM_OK) return rc;
  lse_finalize_kernel<<<ceil_div(Ml, 256), 256, 0, st>>>(workspace, ntiles, Ml, inv_tau, lse);
  return check_launch();
}

int xm_infonce_lse_f32(const float* a, const float* b, float* lse, float* diag, int64_t Ml, int64_t Ng, int64_t D,
                       float inv_tau, int64_t diag_off, float* workspace, void* stream) {
  return infonce_lse_impl(a, b, nullptr, 0, 0, lse, diag, Ml, Ng, D, inv_tau, diag_off, workspace, stream);
}

int xm_infonce_lse_peers_f32(const float* a, const void* const* b_peers, int n_peers, int64_t rows_per_peer, float* lse,
                             float* diag, int64_t Ml, int64_t D, float inv_tau, int64_t diag_off, float* workspace,
                             void* stream) {
  if (n_peers <= 0 || rows_per_peer <= 0) return XM_ERR_INVALID;
  return infonce_lse_impl(a, nullptr, b_peers, n_peers, rows_per_peer, lse, diag, Ml, n_peers * rows_per_peer, D, inv_tau,
                          diag_off, workspace, stream);
}

static int infonce_grad_impl(const float* a, const float* b, const void* const* b_peers, int n_peers, int64_t peer_rows,
                             const float* lse_row, const float* lse_col, float* G, int64_t Ml, int64_t Ng, int64_t D,
                             float inv_tau, int64_t diag_off, float coef, int round_out, void* stream) {
  if (!a || (!b && !b_peers) || !lse_row || !lse_col || !G || Ml <= 0 || Ng <= 0 || D <= 0 || (D & 3)) return XM_ERR_INVALID;
  GemmParams p;
  zero_params(p);
  sim_operands(p, 128);
  p.kin_count = ceil_div(D, 32);
  p.M = (int)Ml;
  p.N = (int)Ng;
  p.alpha = inv_tau;
  p.round_tf32 = round_out;
  p.c = G;
  p.ldc = Ng;
  p.lse_row = lse_row;
  p.lse_col = lse_col;
  p.diag_off = (int)diag_off;
  p.coef = coef;
  p.n_peers = n_peers;
  p.peer_rows = (int)peer_rows;
  TensorView3 ta{a, {(unsigned long long)D, (unsigned long long)Ml, 1}, {(unsigned long long)D * 4, (unsigned long long)Ml * D * 4}};
  TensorView3 tb{b ? (const void*)b : b_peers[0], {(unsigned long long)D, (unsigned long long)Ng, 1}, {(unsigned long long)D * 4, (unsigned long long)Ng * D * 4}};
  const TensorView3 tc = TensorView3{G, {(unsigned long long)(Ng), (unsigned long long)(Ml), (unsigned long long)(1)}, {(unsigned long long)(Ng) * 4, (unsigned long long)(Ml * Ng) * 4}};
  return launch_gemm(EPI_NCE_GRAD, ta, tb, tc, p, dim3(ceil_div(Ml, 128), ceil_div(Ng, 128), 1), (cudaStream_t)stream, b_peers);
}

int xm_infonce_grad_f32(const float* a, const float* b, const float* lse_row, const float* lse_col, float* G,
                        int64_t Ml, int64_t Ng, int64_t D, float inv_tau, int64_t diag_off, float coef, int round_out,
                        void* stream) {
  return infonce_grad_impl(a, b, nullptr, 0, 0, lse_row, lse_col, G, Ml, Ng, D, inv_tau, diag_off, coef, round_out, stream);
}

int xm_infonce_grad_peers_f32(const float* a, const void* const* b_peers, int n_peers, int64_t rows_per_peer,
                              const float* lse_row, const float* lse_col, float* G, int64_t Ml, int64_t D, float inv_tau,
                              int64_t diag_off, float coef, int round_out, void* stream) {
  if (n_peers <= 0 || rows_per_peer <= 0) return XM_ERR_INVALID;
  return infonce_grad_impl(a, nullptr, b_peers, n_peers, rows_per_peer, lse_row, lse_col, G, Ml, n_peers * rows_per_peer, D,
                           inv_tau, diag_off, coef, round_out, stream);
}

}  // extern "C"

// ------------------------------------------------------------------ bring-up probe
// Loads ONE TMA box {32, 32, 1} of a 2-D fp32 tensor at (c0, c1) into shared memory and dumps the
// raw (still swizzled) 4 KB image: used by tests/gpu_bringup.py to establish which coordinates /
// swizzle modes the TMA unit accepts on this part.  Not on any product path.
namespace xm {
__global__ void tma_probe_kernel(const __grid_constant__ CUtensorMap tm, int c0, int c1, float* out) {
  __shared__ __align__(1024) float tile[1024];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(&bar, 4096);
    ptx::tma_load_3d(&tm, &bar, tile, c0, c1, 0);
  }
  ptx::mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = tile[i];
}
}  // namespace xm

extern "C" int xm_debug_set_conv_halo(int on) {
  const int old = xm::g_conv_halo;
  xm::g_conv_halo = on ? 1 : 0;
  return old;
}

extern "C" int xm_debug_tma_probe(const float* src, int64_t rows, int64_t cols, int64_t ld, int c0, int c1,
                                  int swizzle_atom32, float* out, void* stream) {
  TensorView3 tv{src, {(unsigned long long)cols, (unsigned long long)rows, 1},
                 {(unsigned long long)ld * 4, (unsigned long long)rows * ld * 4}};
  CUtensorMap tm;
  int rc = encode_tmap(&tm, tv, 32, 32, swizzle_atom32);
  if (rc != XM_OK) return rc;
  tma_probe_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(tm, c0, c1, out);
  return check_launch();
}
