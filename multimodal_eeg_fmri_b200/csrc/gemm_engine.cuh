// tcgen05 TF32 GEMM engine (sm_100a).
//
//   D[m, n] (+)= sum over k-blocks of  A_tile[m, 32] * B_tile[n, 32]^T      (fp32 accumulate in TMEM)
//
// Persistent kernel: one CTA per SM walks 128 x BN output tiles (or `taps_n` of them, one TMEM
// accumulator per conv tap, for the conv weight gradient).  Warp roles: warp 0 = TMA producer (one
// lane), warp 1 = TMEM allocator + tcgen05.mma issuer (one lane), warps 2..5 = epilogue
// (TMEM -> registers -> 128B-swizzled smem -> TMA store).  Two accumulator sets in TMEM let the
// epilogue of tile i overlap the loads and MMAs of tile i+1.  Operand tiles are fetched by TMA (cp.async.bulk.tensor.3d)
// into 128B-swizzled shared memory and consumed directly by tcgen05.mma.kind::tf32 through
// shared-memory descriptors; fp32 data is used as-is (the tensor core reads the tf32 bits).
//
// Both operands may be K-major (rows of 32 consecutive k) or MN-major (rows of 32 consecutive
// m/n, one row per k) so that forward, data-gradient and weight-gradient contractions all read
// the SAME row-major fp32 tensors without transposed copies.  Conv1d is the same engine with a
// 3-D tensor map over channels-last activations (C, T, B): a tap is a TMA coordinate shift along
// the row (time) axis, and TMA's out-of-bounds zero fill implements the conv zero padding and the
// ragged last tile.  (TMA inner coordinates must be 16-byte aligned -- measured on B200 -- which is
// why time cannot be the contiguous axis of a conv operand.)
#pragma once
#include "xm_common.cuh"
#include "xm_ptx.cuh"

namespace xm {

enum : int {
  EPI_ROWMAJOR = 0,    // C[z][tap][m][n] = act(alpha*acc + bias[n])
  EPI_LSE = 2,         // partial[ny][m]  = sum_n exp(alpha*acc - shift) ; diag[m] = alpha*acc[m, m+diag_off]
  EPI_NCE_GRAD = 3,    // C[m][n] = coef*(exp(s-lse_row[m]) + exp(s-lse_col[n]) - 2*[n == m+diag_off]),  s = alpha*acc
};

struct OperandCfg {
  int mn_major;  // 0: K-major tile, one TMA box {32 k, rows, 1}; 1: MN-major, rows/32 boxes {32 mn, 32 k, 1}
  int rows;      // MN extent of the tile (A: 128, B: bn)
  // TMA coordinate d = base[d] + bx*sx[d] + by*sy[d] + bz*sz[d] + kin*kin_step[d] + kout*kout_step[d] + tap*tap_step[d]
  int base[3], sx[3], sy[3], sz[3], kin_step[3], kout_step[3], tap_step[3];
};

constexpr int kMaxPeers = 8;
// One tensor map per rank for an operand that is ROW-SHARDED across the GPUs of a node (peer memory mapped
// into this process through NVLink / NVSwitch): the TMA producer reads each tile from the shard of the rank
// that owns it, so "all-gather, then GEMM" becomes ONE kernel whose loads cross NVLink tile by tile.
struct PeerMaps {
  CUtensorMap m[kMaxPeers];
};

struct GemmParams {
  OperandCfg a, b;
  int b_rows_dim2; // MN-major B whose contraction rows run along tensor dimension 2 (TMA box {32, 1, 32})
  int n_peers;     // > 0: operand B lives in n_peers shards of peer_rows rows each (row coordinate = dim 1)
  int peer_rows;
  int bn;          // N of one MMA / one accumulator (multiple of 16, <= 256)
  int taps_k;      // taps iterated inside the K loop (conv fwd / dgrad), >= 1
  int taps_n;      // taps held as separate accumulators (conv wgrad), >= 1
  int b_halo;      // MN-major B with taps_n > 1: ONE halo tile of 32 + b_halo k-rows per 32-wide MN block serves
                   // every tap (tap t = the same tile read from row t: a +128 B start-address shift of the
                   // shared-memory descriptor) instead of taps_n shifted copies
  int b_blk_bytes; // bytes per 32-wide MN block of the halo tile (1024-B multiple)
  int a_halo;      // K-major A with taps_k > 1 (conv fwd / dgrad): ONE halo slab of 128 + a_halo rows per k-block
                   // serves every tap (tap t = the slab read from row a_tap_row0 + t*a_tap_dir) and the stage
                   // carries the taps_k weight tiles: the activations cross L2 -> smem once, not taps_k times
  int a_slab_bytes, a_tap_row0, a_tap_dir, a_halo_row_shift;  // slab starts a_halo_row_shift rows from a.base[1]
  int a_wrap;      // > 0 (halo path): inner coordinates >= a_wrap wrap back by a_wrap -- a 3-pass conv reads the channel
                   // blocks [hi | lo | hi] of a tensor that stores [hi | lo] only
  int kin_count;   // inner k-blocks per (kout, tap)
  int kout_count;  // outer k iterations per tile (split along the z tile index when kout_split != 0)
  int kout_total;  // total outer k iterations (only used when kout_split != 0)
  int kout_split;  // 1: the z tile index selects a kout range [bz*kout_count, ...)
  int stages;
  int tmem_cols;   // allocated TMEM columns (power of two)
  int acc_bufs;    // 1 or 2 accumulator sets of taps_n*bn columns (2: epilogue of tile i overlaps MMA of tile i+1)
  int acc_chunk;   // > 0: the tensor core accumulates only acc_chunk k-blocks at a time; the epilogue warps add
                   // the chunks in fp32 round-to-nearest into a third TMEM region (columns [2*bn, 3*bn)).  The
                   // MMA's own accumulation truncates (a bias of ~3e-8 per K=8 step, i.e. 4e-4 over K = 120 000),
                   // which the fp32-accurate 3-pass projections cannot afford.  Needs acc_bufs == 2, taps_n == 1.
  int nx, ny, nz;  // tile grid (persistent CTAs walk tile = by + ny*(bx + nx*bz))
  // epilogue
  int M, N;        // valid extents of the output (guards)
  int tma_store;   // 1: tile staged in swizzled smem and written by TMA through tmC (clips ragged edges)
  int n_stride;    // logical column of a tile's first column = by*n_stride (launch_gemm default: bn)
  int c_col_base, c_col_mul;        // tmC column coordinate = c_col_base + by*c_col_mul + c0 (default 0, bn)
  int c_z_mul, c_y_mul, c_tap_mul;  // tmC z coordinate = bz*c_z_mul + by*c_y_mul + tn*c_tap_mul
  float* c;        // direct-store path (outputs TMA cannot address: pitch or base not 16-B aligned)
  long long ldc, c_z_stride, c_tap_stride;
  const float* bias;
  float alpha;
  int act;
  int round_tf32;
  double* colstat_part;  // EPI_ROWMAJOR with TMA store, bn % 64 == 0, ny * bn <= 256, may be NULL: (gridDim.x * 4, N, 2) per-(CTA, lane quadrant)
                         // column sums and sums of squares of the OUTPUT over the rows this CTA produced -- the
                         // BatchNorm batch statistics of a conv output without a pass over it (xm_bn_finalize_stats input)
  // InfoNCE epilogues
  const float* lse_row;
  const float* lse_col;
  float* partial;  // (ny, M)
  float* diag;     // (M)
  int diag_off;
  float coef;
  float shift;
};

// Epilogue warps per kernel flavour: 4 per TMEM lane quadrant "part"; the parts split a tile's columns.
// A single warp per scheduler runs a dependent TMEM-load -> ALU -> store chain at low IPC, so the storing
// flavours use 8 warps (the row-logsumexp flavour, which only reduces, 4).
template <int EPI>
struct EpiWarps {
  static constexpr int value = EPI == 2 ? 4 : 8;
};
constexpr int gemm_threads(int epi_warps) { return 64 + 32 * epi_warps; }
constexpr int kATileBytes = 128 * 128;  // 128 rows x 32 tf32
constexpr int kMaxStages = 8;
// staging: one 32 x 32 fp32 box (4 KB) per buffer; 2 buffers per warp up to 8 warps, 1 beyond
constexpr int staging_bufs(int epi_warps) { return epi_warps <= 8 ? 2 : 1; }
constexpr int staging_bytes(int epi_warps) { return epi_warps * staging_bufs(epi_warps) * 4096; }

// The whole producer warp runs the load loops; `lane` picks who issues which box, so the 4 KB boxes of an
// MN-major tile (and the per-tap weight tiles of a halo stage) are issued by different lanes of ONE warp
// instruction instead of one after another by a single thread.
XM_DEVICE void issue_operand_loads(const CUtensorMap* tm, uint64_t* bar, uint8_t* dst, const OperandCfg& o, int c0,
                                   int c1, int c2, int lane, int lane0 = 0) {
  if (!o.mn_major) {
    if (lane == lane0) ptx::tma_load_3d(tm, bar, dst, c0, c1, c2);
  } else {
    const int nbox = o.rows >> 5;
    const int bx = lane - lane0;
    if (bx >= 0 && bx < nbox) ptx::tma_load_3d(tm, bar, dst + bx * 4096, c0 + 32 * bx, c1, c2);
  }
}

struct TileCoord {
  int bx, by, bz, kout_lo, kout_n;
};
XM_DEVICE TileCoord decode_tile(const GemmParams& p, int tile) {
  TileCoord t;
  t.by = tile % p.ny;
  const int r = tile / p.ny;
  t.bx = r % p.nx;
  t.bz = r / p.nx;
  t.kout_lo = 0;
  t.kout_n = p.kout_count;
  if (p.kout_split) {
    t.kout_lo = t.bz * p.kout_count;
    t.kout_n = min(p.kout_count, p.kout_total - t.kout_lo);
  }
  return t;
}

XM_DEVICE float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 16-bit dropout lanes: one 64-bit hash per 4 consecutive columns of a row
XM_DEVICE uint64_t hash_u64(uint64_t idx, uint64_t seed) {
  uint64_t z = idx + epoch_seed(seed) * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

template <int EPI>
__global__ void __launch_bounds__(gemm_threads(EpiWarps<EPI>::value), 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ PeerMaps peers, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kMaxStages];
  __shared__ uint64_t empty_bar[kMaxStages];
  __shared__ uint64_t tmem_full_bar[2];
  __shared__ uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;
  constexpr int EW = EpiWarps<EPI>::value;
  constexpr int NPARTS = EW / 4;
  constexpr int NBUF = staging_bufs(EW);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int ntiles = p.nx * p.ny * p.nz;

  // 1024-B aligned operand ring (128B swizzle atoms are 1024 B), then the epilogue staging buffers
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int b_tile_bytes = p.b_halo ? (p.bn >> 5) * p.b_blk_bytes : p.bn * 128;
  const int a_bytes = p.a_halo ? p.a_slab_bytes : kATileBytes;
  const int stage_bytes = a_bytes + (p.b_halo ? 1 : (p.a_halo ? p.taps_k : p.taps_n)) * b_tile_bytes;
  const int stage_tx_bytes = p.b_halo   ? kATileBytes + (p.bn >> 5) * (32 + p.b_halo) * 128
                             : p.a_halo ? (128 + p.a_halo) * 128 + p.taps_k * b_tile_bytes
                                        : stage_bytes;
  const int taps_k_loop = p.a_halo ? 1 : p.taps_k;  // with a halo slab the taps live inside one stage
  uint8_t* staging = smem + (size_t)p.stages * stage_bytes;
  const int acc_cols = p.taps_n * p.bn;

  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    if (p.tma_store) ptx::prefetch_tensormap(&tmC);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&tmem_full_bar[b], 1);
      ptx::mbar_init(&tmem_empty_bar[b], EW);  // one arrival per epilogue warp
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(&tmem_base_slot, (uint32_t)p.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    {
      // ------------------------------------------------ TMA producer (full warp; lane 0 owns the barriers)
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(p, tile);
        int ca[3], cb[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          ca[d] = p.a.base[d] + t.bx * p.a.sx[d] + t.by * p.a.sy[d] + t.bz * p.a.sz[d];
          cb[d] = p.b.base[d] + t.bx * p.b.sx[d] + t.by * p.b.sy[d] + t.bz * p.b.sz[d];
        }
        for (int ko = 0; ko < t.kout_n; ++ko) {
          const int kout = t.kout_lo + ko;
          for (int tk = 0; tk < taps_k_loop; ++tk) {
            for (int kin = 0; kin < p.kin_count; ++kin) {
              ptx::mbar_wait(&empty_bar[s], ph ^ 1u);
              if (lane == 0) ptx::mbar_arrive_expect_tx(&full_bar[s], (uint32_t)stage_tx_bytes);
              __syncwarp();  // the transaction count is armed before any lane's load can complete
              uint8_t* sa = smem + (size_t)s * stage_bytes;
              if (p.a_halo) {
                // halo slab (rows of every tap) by lane 0 + the taps_k weight tiles of this k-block, one lane each
                if (lane == 0) {
                  int c0 = ca[0] + kin * p.a.kin_step[0] + kout * p.a.kout_step[0];
                  if (p.a_wrap > 0 && c0 >= p.a_wrap) c0 -= p.a_wrap;
                  ptx::tma_load_3d(&tmA, &full_bar[s], sa, c0,
                                   ca[1] + kin * p.a.kin_step[1] + kout * p.a.kout_step[1] + p.a_halo_row_shift,
                                   ca[2] + kin * p.a.kin_step[2] + kout * p.a.kout_step[2]);
                }
                for (int tp = 0; tp < p.taps_k; ++tp)
                  issue_operand_loads(&tmB, &full_bar[s], sa + a_bytes + tp * b_tile_bytes, p.b,
                                      cb[0] + kin * p.b.kin_step[0] + kout * p.b.kout_step[0] + tp * p.b.tap_step[0],
                                      cb[1] + kin * p.b.kin_step[1] + kout * p.b.kout_step[1] + tp * p.b.tap_step[1],
                                      cb[2] + kin * p.b.kin_step[2] + kout * p.b.kout_step[2] + tp * p.b.tap_step[2], lane,
                                      1 + tp);
                if (++s == p.stages) { s = 0; ph ^= 1u; }
                continue;
              }
              issue_operand_loads(&tmA, &full_bar[s], sa, p.a,
                                  ca[0] + kin * p.a.kin_step[0] + kout * p.a.kout_step[0] + tk * p.a.tap_step[0],
                                  ca[1] + kin * p.a.kin_step[1] + kout * p.a.kout_step[1] + tk * p.a.tap_step[1],
                                  ca[2] + kin * p.a.kin_step[2] + kout * p.a.kout_step[2] + tk * p.a.tap_step[2], lane);
              if (p.b_halo) {  // one halo box per MN block (rows of tap 0 ... tap taps_n-1 overlap), lanes 8..
                const int nbox = p.bn >> 5;
                const int bxi = lane - 8;
                if (bxi >= 0 && bxi < nbox)
                  ptx::tma_load_3d(&tmB, &full_bar[s], sa + a_bytes + bxi * p.b_blk_bytes,
                                   cb[0] + 32 * bxi + kin * p.b.kin_step[0] + kout * p.b.kout_step[0],
                                   cb[1] + kin * p.b.kin_step[1] + kout * p.b.kout_step[1],
                                   cb[2] + kin * p.b.kin_step[2] + kout * p.b.kout_step[2]);
              } else
              for (int tn = 0; tn < p.taps_n; ++tn) {
                const int tap = tk + tn;  // exactly one of taps_k / taps_n exceeds 1
                int b1 = cb[1] + kin * p.b.kin_step[1] + kout * p.b.kout_step[1] + tap * p.b.tap_step[1];
                const CUtensorMap* tb = &tmB;
                if (p.n_peers > 0) {  // row-sharded operand: this tile's rows live on rank b1 / peer_rows
                  const int owner = b1 / p.peer_rows;
                  b1 -= owner * p.peer_rows;
                  tb = &peers.m[owner];
                }
                issue_operand_loads(tb, &full_bar[s], sa + a_bytes + tn * b_tile_bytes, p.b,
                                    cb[0] + kin * p.b.kin_step[0] + kout * p.b.kout_step[0] + tap * p.b.tap_step[0], b1,
                                    cb[2] + kin * p.b.kin_step[2] + kout * p.b.kout_step[2] + tap * p.b.tap_step[2], lane, 8);
              }
              if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      // ------------------------------------------------ MMA issuer
      // The schedule runs warp-uniformly (all lanes wait on the barriers and form the descriptors, which the compiler
      // keeps in uniform registers); the MMAs and commits of a k-block issue from one `elect_one` branch.  Inside
      // `if (lane == 0)` every MMA paid an R2UR move and an ELECT loop (profiles/r2_mma_rate_probe.txt).
      const uint32_t idesc = ptx::make_idesc_tf32(128, p.bn, p.a.mn_major, p.b.mn_major);
      const uint32_t a_kstep = p.a.mn_major ? 1024u : 32u;  // bytes per K=8 slab
      const uint32_t b_kstep = p.b.mn_major ? 1024u : 32u;
      const uint32_t a_lbo = p.a.mn_major ? 4096u : 16u;
      const uint32_t b_lbo = p.b_halo ? 128u : (p.b.mn_major ? 4096u : 16u);
      const uint32_t idesc_halo = ptx::make_idesc_tf32(128, p.taps_n * 32, p.a.mn_major, p.b.mn_major);
      const uint32_t a_sbo = p.a.mn_major ? 512u : 1024u;
      const uint32_t b_sbo = p.b.mn_major ? 512u : 1024u;
      const uint32_t a_lt = p.a.mn_major ? 1u : 2u;
      const uint32_t b_lt = p.b.mn_major ? 1u : 2u;
      // Descriptors are built ONCE (for stage 0, K-slab 0, tap 0); every other operand view is the same
      // descriptor with its 14-bit start-address field advanced by (byte offset >> 4).  The single issuing
      // thread then spends ~4 integer instructions per MMA instead of ~30 -- with N = 64 tiles an MMA lasts
      // only 32 cycles, so descriptor arithmetic was what paced the conv kernels.
      const uint32_t smem0 = ptx::smem_u32(smem);
      const uint64_t da0 = ptx::make_smem_desc(smem0, a_lbo, a_sbo, a_lt);
      const uint64_t db0 = ptx::make_smem_desc(smem0 + a_bytes, b_lbo, b_sbo, b_lt);
      const uint32_t stage16 = (uint32_t)stage_bytes >> 4;
      const uint32_t a_k16 = a_kstep >> 4, b_k16 = b_kstep >> 4;
      const uint32_t b_tap16 = (uint32_t)(p.b_halo ? 128 : b_tile_bytes) >> 4;
      int s = 0;
      uint32_t ph = 0;
      int ait = 0;  // accumulator-set uses so far (one per tile, or one per chunk of a tile with acc_chunk)
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(p, tile);
        const int total_kb = t.kout_n * taps_k_loop * p.kin_count;
        int buf = (p.acc_bufs == 2) ? (ait & 1) : 0;
        uint32_t use = (uint32_t)(p.acc_bufs == 2 ? (ait >> 1) : ait);
        ++ait;
        ptx::mbar_wait(&tmem_empty_bar[buf], (use & 1u) ^ 1u);  // epilogue drained this accumulator set
        ptx::tc_fence_after_sync();
        uint32_t acc = tmem_base + (uint32_t)(buf * acc_cols);
        int kb0 = 0;  // first k-block of the running accumulation
        for (int kb = 0; kb < total_kb; ++kb) {
          if (p.acc_chunk > 0 && kb - kb0 == p.acc_chunk) {  // hand this chunk over, continue in the other set
            if (ptx::elect_one()) ptx::mma_commit(&tmem_full_bar[buf]);
            __syncwarp();
            buf = ait & 1;
            use = (uint32_t)(ait >> 1);
            ++ait;
            ptx::mbar_wait(&tmem_empty_bar[buf], (use & 1u) ^ 1u);
            ptx::tc_fence_after_sync();
            acc = tmem_base + (uint32_t)(buf * acc_cols);
            kb0 = kb;
          }
          ptx::mbar_wait(&full_bar[s], ph);
          ptx::tc_fence_after_sync();
          const uint64_t das = da0 + (uint64_t)((uint32_t)s * stage16);
          const uint64_t dbs = db0 + (uint64_t)((uint32_t)s * stage16);
          const bool last_kb = kb == total_kb - 1;
          if (p.a_halo) {  // taps inside the stage: A = halo slab shifted by whole rows (128 B), B = tap tile
            if (ptx::elect_one()) {
              for (int tp = 0; tp < p.taps_k; ++tp) {
                const uint64_t dat = das + (uint64_t)((uint32_t)(p.a_tap_row0 + tp * p.a_tap_dir) * 8u);
                const uint64_t dbt = dbs + (uint64_t)((uint32_t)tp * ((uint32_t)b_tile_bytes >> 4));
#pragma unroll
                for (int k8 = 0; k8 < 4; ++k8)
                  ptx::mma_tf32_ss(acc, dat + (uint64_t)(k8 * a_k16), dbt + (uint64_t)(k8 * b_k16), idesc,
                                   (kb > 0 || tp > 0 || k8 > 0) ? 1u : 0u);
              }
              ptx::mma_commit(&empty_bar[s]);
              if (last_kb) ptx::mma_commit(&tmem_full_bar[buf]);  // accumulator complete
            }
            __syncwarp();
            if (++s == p.stages) { s = 0; ph ^= 1u; }
            continue;
          }
          if (ptx::elect_one()) {
            if (p.b_halo) {
              // Halo tile: the taps are 128-B (one k-row) shifts of the same 32-wide block, so ONE MMA per
              // block covers all taps at once -- its N dimension walks taps_n "blocks" that are 128 B apart
              // (leading-dimension byte offset = 128).  N = 32*taps instead of bn per MMA: 3.5x fewer reads of
              // the A tile from shared memory, which is what bounded the small-N tap-by-tap MMAs.
              const int nblk = p.bn >> 5;
              for (int cb = 0; cb < nblk; ++cb) {
                const uint64_t dbc = dbs + (uint64_t)((uint32_t)cb * ((uint32_t)p.b_blk_bytes >> 4));
                const uint32_t acc_c = acc + (uint32_t)(cb * p.taps_n * 32);
#pragma unroll
                for (int k8 = 0; k8 < 4; ++k8)
                  ptx::mma_tf32_ss(acc_c, das + (uint64_t)(k8 * a_k16), dbc + (uint64_t)(k8 * b_k16), idesc_halo,
                                   (kb > 0 || k8 > 0) ? 1u : 0u);
              }
            } else
            for (int tn = 0; tn < p.taps_n; ++tn) {
              const uint64_t dbt = dbs + (uint64_t)((uint32_t)tn * b_tap16);
              const uint32_t acc_t = acc + (uint32_t)(tn * p.bn);
#pragma unroll
              for (int k8 = 0; k8 < 4; ++k8)
                ptx::mma_tf32_ss(acc_t, das + (uint64_t)(k8 * a_k16), dbt + (uint64_t)(k8 * b_k16), idesc,
                                 (kb > kb0 || k8 > 0) ? 1u : 0u);
            }
            ptx::mma_commit(&empty_bar[s]);  // frees the smem slot when these MMAs retire
            if (last_kb) ptx::mma_commit(&tmem_full_bar[buf]);  // accumulator complete
          }
          __syncwarp();
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
        if (total_kb == 0) {  // nothing to accumulate (cannot happen for a launched tile; keeps the hand-off complete)
          if (ptx::elect_one()) ptx::mma_commit(&tmem_full_bar[buf]);
          __syncwarp();
        }
      }
    }
  } else {
    // -------------------------------------------------- epilogue warps 2..5
    const int q = warp & 3;            // TMEM lane quadrant this warp may read (hardware: warp id % 4)
    const int part = (warp - 2) >> 2;  // which share of the tile's columns
    const int row = q * 32 + lane;
    uint8_t* stg = staging + (warp - 2) * (NBUF * 4096);  // this warp's staging buffer(s)
    int chunk_ctr = 0;
    int ait = 0;  // mirrors the MMA warp's accumulator-set counter
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t sum = tmem_base + (uint32_t)(2 * acc_cols) + lane_base;  // fp32 running sum (acc_chunk only)
    double cstat[4][2];  // column statistics of this warp's (up to 4) 32-column chunks: lane j owns column c0 + j
#pragma unroll
    for (int i = 0; i < 4; ++i) cstat[i][0] = cstat[i][1] = 0.0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      bool add_sum = false;
      if (EPI == EPI_ROWMAJOR && p.acc_chunk > 0) {
        // every chunk but the last: TMEM partial -> registers -> (+ running sum) -> TMEM sum region.  A warp only
        // ever touches its own lane quadrant and its own 32-column groups, here and in the store loop below.
        const int n_chunks = (t.kout_n * taps_k_loop * p.kin_count + p.acc_chunk - 1) / p.acc_chunk;
        for (int c = 0; c + 1 < n_chunks; ++c) {
          const int cbuf = ait & 1;
          ptx::mbar_wait(&tmem_full_bar[cbuf], (uint32_t)(ait >> 1) & 1u);
          ++ait;
          ptx::tc_fence_after_sync();
          const uint32_t src = tmem_base + (uint32_t)(cbuf * acc_cols) + lane_base;
          for (int c0 = part * 32; c0 < p.bn; c0 += 32 * NPARTS) {
            uint32_t r[32];
            ptx::tmem_ld_32x32(src + (uint32_t)c0, r);
            if (c > 0) {
              uint32_t u[32];
              ptx::tmem_ld_32x32(sum + (uint32_t)c0, u);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(u[j]));
            } else {
              ptx::tmem_ld_wait();
            }
            ptx::tmem_st_32x32(sum + (uint32_t)c0, r);
          }
          ptx::tmem_st_wait();
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[cbuf]);
        }
        add_sum = n_chunks > 1;
      }
      const int buf = (p.acc_bufs == 2) ? (ait & 1) : 0;
      const uint32_t use = (uint32_t)(p.acc_bufs == 2 ? (ait >> 1) : ait);
      ++ait;
      ptx::mbar_wait(&tmem_full_bar[buf], use & 1u);
      ptx::tc_fence_after_sync();
      const uint32_t acc = tmem_base + (uint32_t)(buf * acc_cols) + lane_base;
      const int m = t.bx * 128 + row;  // row index inside this z-slab
      const bool row_ok = m < p.M;
      const int n0 = t.by * p.n_stride;
      const int col0 = p.c_col_base + t.by * p.c_col_mul;  // tmC column of this tile's first column

      float lse_r = 0.f;
      if (EPI == EPI_NCE_GRAD && row_ok) lse_r = p.lse_row[m];
      float rowsum = 0.f;

      for (int tn = 0; tn < p.taps_n; ++tn) {
        if (EPI != EPI_LSE && p.tma_store) {
          // ---- 32-column chunks: TMEM -> registers -> swizzled smem -> TMA store (edges clipped by TMA)
          const int zc = t.bz * p.c_z_mul + t.by * p.c_y_mul + tn * p.c_tap_mul;
          for (int c0 = part * 32; c0 < p.bn; c0 += 32 * NPARTS) {
            if (n0 + c0 >= p.N) break;  // warp-uniform: chunk entirely outside the output
            uint32_t r[32];
            if (c0 + 32 <= p.bn) {
              ptx::tmem_ld_32x32(acc + (uint32_t)(tn * p.bn + c0), r);
            } else {  // bn % 32 == 16 tail
              uint32_t h[16];
              ptx::tmem_ld_32x16(acc + (uint32_t)(tn * p.bn + c0), h);
#pragma unroll
              for (int j = 0; j < 16; ++j) { r[j] = h[j]; r[16 + j] = 0u; }
            }
            ptx::tmem_ld_wait();
            if (EPI == EPI_ROWMAJOR && add_sum) {  // chunked accumulation: last partial + running sum
              uint32_t u[32];
              ptx::tmem_ld_32x32(sum + (uint32_t)c0, u);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(u[j]));
            }
            // Straight-line code here is executed once per chunk by a single warp per scheduler, so its
            // SIZE matters (instruction fetch): every runtime switch is hoisted out of the element loop.
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
            const int nb = n0 + c0;
            if (EPI == EPI_ROWMAJOR) {
              if (p.bias != nullptr) {
                if (nb + 32 <= p.N && ((reinterpret_cast<uintptr_t>(p.bias + nb) & 15) == 0)) {
                  const float4* b4 = reinterpret_cast<const float4*>(p.bias + nb);
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float4 b = __ldg(b4 + j);
                    v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
                  }
                } else {
                  const float bl = (nb + lane < p.N) ? __ldg(p.bias + nb + lane) : 0.f;
#pragma unroll
                  for (int j = 0; j < 32; ++j) v[j] += __shfl_sync(0xffffffffu, bl, j);
                }
              }
              if (p.act == XM_ACT_GELU) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
              } else if (p.act == XM_ACT_RELU) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
              }  // tanh / sigmoid are applied by a separate pass (host side): keeps this code compact and in registers
              if (p.round_tf32) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = round_tf32(v[j]);
              }
            } else {  // EPI_NCE_GRAD
              const float lc = (nb + lane < p.N) ? __ldg(p.lse_col + nb + lane) : 0.f;
              const int dj = m + p.diag_off - nb;  // column of this row's positive inside the chunk (if in [0, 32))
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float lcj = __shfl_sync(0xffffffffu, lc, j);
                float g = __expf(v[j] - lse_r) + __expf(v[j] - lcj);
                if (j == dj) g -= 2.0f;
                v[j] = g * p.coef;  // out-of-matrix parts are clipped by TMA
              }
              if (p.round_tf32) {  // G feeds a single-pass tf32 product directly (else: split 3-way first)
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = round_tf32(v[j]);
              }
            }
            if (EPI == EPI_ROWMAJOR && p.colstat_part != nullptr) {
              // batch statistics of the output: column sums over this warp's 32 rows (rows past M excluded), then
              // one fp64 add per lane and chunk -- rounding of the tile-level fp32 sums stays at the 1e-7 level
              uint32_t u[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) u[j] = row_ok ? __float_as_uint(v[j]) : 0u;
              const float s1 = warp_column_sums(u, lane);
#pragma unroll
              for (int j = 0; j < 32; ++j) u[j] = row_ok ? __float_as_uint(v[j] * v[j]) : 0u;
              const float s2 = warp_column_sums(u, lane);
              const int slot = t.by * (p.bn / (32 * NPARTS)) + (c0 - part * 32) / (32 * NPARTS);  // (N tile, chunk of this warp)
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (i == slot) {
                  cstat[i][0] += (double)s1;
                  cstat[i][1] += (double)s2;
                }
            }
            uint8_t* sb = stg + (NBUF == 2 ? (chunk_ctr & 1) * 4096 : 0);
            if (lane == 0) ptx::bulk_wait_read<NBUF - 1>();  // the store that last read this buffer has drained it
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j)  // 16-B chunk j of row `lane` lives at chunk j ^ (lane & 7) (SWIZZLE_128B)
              *reinterpret_cast<float4*>(sb + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                  make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              ptx::tma_store_3d(&tmC, sb, col0 + c0, t.bx * 128 + q * 32, zc);
              ptx::bulk_commit();
            }
            ++chunk_ctr;
          }
        } else {
          float* cbase = p.c ? p.c + (long long)t.bz * p.c_z_stride + (long long)tn * p.c_tap_stride : nullptr;
          for (int c0 = part * 16; c0 < p.bn; c0 += 16 * NPARTS) {
            uint32_t r[16];
            // halo layout: [ci block][tap][32 columns]; plain layout: [tap][bn columns]
            const uint32_t acol = p.b_halo ? (uint32_t)((c0 >> 5) * (p.taps_n * 32) + tn * 32 + (c0 & 31))
                                           : (uint32_t)(tn * p.bn + c0);
            ptx::tmem_ld_32x16(acc + acol, r);
            ptx::tmem_ld_wait();
            if (EPI == EPI_LSE) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int n = n0 + c0 + j;
                if (row_ok && n < p.N) {
                  const float sv = __uint_as_float(r[j]) * p.alpha;
                  rowsum += __expf(sv - p.shift);
                  if (n == m + p.diag_off) p.diag[m] = sv;
                }
              }
            } else if (row_ok) {
              float* crow = cbase + (long long)m * p.ldc + n0 + c0;
#pragma unroll 1
              for (int j = 0; j < 16; ++j) {  // small / unaligned outputs only: compact code over speed
                const int n = n0 + c0 + j;
                if (n >= p.N) break;
                float x = __uint_as_float(r[j]) * p.alpha;
                if (EPI == EPI_ROWMAJOR) {
                  if (p.bias != nullptr) x += __ldg(p.bias + n);
                  x = apply_act(x, p.act);
                  if (p.round_tf32) x = round_tf32(x);
                } else {
                  float g = __expf(x - lse_r) + __expf(x - __ldg(p.lse_col + n));
                  if (n == m + p.diag_off) g -= 2.0f;
                  x = g * p.coef;
                  if (p.round_tf32) x = round_tf32(x);
                }
                crow[j] = x;
              }
            }
          }
        }
      }
      if (EPI == EPI_LSE && row_ok) p.partial[(long long)t.by * p.M + m] = rowsum;
      // hand the accumulator set back to the MMA warp
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[buf]);
    }
    if (EPI == EPI_ROWMAJOR && p.colstat_part != nullptr) {
      double* dst = p.colstat_part + (long long)(blockIdx.x * 4 + q) * p.N * 2;
      const int per_tile = p.bn / (32 * NPARTS);  // 32-column chunks of one N tile that this warp owns (host: ny * per_tile <= 4)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int by = i / per_tile, ch = i - by * per_tile;
        const int col = by * p.n_stride + part * 32 + ch * 32 * NPARTS + lane;
        if (by < p.ny && col < p.N) {
          dst[col * 2 + 0] = cstat[i][0];
          dst[col * 2 + 1] = cstat[i][1];
        }
      }
    }
    if (lane == 0) ptx::bulk_wait_all();  // staged tiles fully written before the CTA (and its smem) retires
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ---------------------------------------------------------------- host side
struct TensorView3 {
  const void* ptr;
  unsigned long long dim[3];         // dim[0] innermost (contiguous)
  unsigned long long stride_bytes[2];  // strides of dim[1], dim[2]
};

int encode_tmap(CUtensorMap* out, const TensorView3& t, unsigned box0, unsigned box1, int mn_major, unsigned box2 = 1);
// `tc` describes the output for the TMA-store epilogue (dims {N, M, Z}); pass ptr == nullptr to use
// the direct-store path (p.c / p.ldc).  `grid` is the TILE grid; the launch is persistent.
int launch_gemm(int epi, const TensorView3& ta, const TensorView3& tb, const TensorView3& tc, GemmParams& p, dim3 grid,
                cudaStream_t stream, const void* const* b_peers = nullptr);

inline int tmem_cols_for(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

}  // namespace xm
