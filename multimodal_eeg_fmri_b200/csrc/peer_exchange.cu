// Small all-reduce over NVLink peer memory for the SyncBN statistics of the data-parallel step (SURVEY.md section 8e
// item 4): n <= 1024 doubles per call, 16 calls per step, each on the critical path of the stream that issues it.
// One CTA per rank, no collective library: every rank PUSHES its vector into a slot of every peer's symmetric buffer,
// publishes a sequence number behind a system-scope release, waits until all peers' sequence numbers for the slot have
// arrived in its own buffer, and sums the vectors in rank order (bit-identical on every rank).  A call costs one
// NVLink round of ~2 KB stores instead of a library launch, and -- unlike a communicator -- the two streams of the step
// (EEG encoder, fMRI branch) exchange independently: each uses its own channel (slots + sequence counter).
//
// Slot reuse: call k + S may overwrite the slot of call k in a peer's buffer only after that peer has read it.  A rank
// that issues call k + S has completed call k + S - 1, which needed every peer's push of call k + S - 1, which every
// peer issues after finishing its call k + S - 2: for S >= 2 slots the peer is done with call k.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/xmodal_b200.h"
#include "xm_common.cuh"

namespace xm {
namespace peer {

struct Ptrs {
  double* data_dst[8];               // peer p's slot, row of THIS rank
  unsigned long long* flag_dst[8];   // peer p's flag of THIS rank for the slot
};

XM_DEVICE void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
XM_DEVICE unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256)
peer_allreduce_kernel(const double* __restrict__ x, double* __restrict__ out, int n, Ptrs pp, int world,
                      const double* __restrict__ slot_data, const unsigned long long* __restrict__ slot_flags, int row_stride,
                      unsigned long long seq) {
  __shared__ int timed_out;
  if (threadIdx.x == 0) timed_out = 0;
  // A captured launch replays with its frozen `seq`; the device-resident seed epoch (xm_seed_epoch_advance, one step per
  // replay on every rank) keeps the published numbers distinct: replay k of the call publishes seq + k * 2^32.
  seq += xm_seed_epoch_c << 32;
  for (int p = 0; p < world; ++p)
    for (int i = threadIdx.x; i < n; i += blockDim.x) pp.data_dst[p][i] = x[i];
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < world) {
    st_release_sys(pp.flag_dst[threadIdx.x], seq);
    const long long t0 = clock64();
    while (ld_acquire_sys(slot_flags + threadIdx.x) != seq) {
      if (clock64() - t0 > 20000000000ll) {  // ~10 s: a peer never arrived (program error) -- poison instead of hanging
        timed_out = 1;
        break;
      }
    }
  }
  __syncthreads();
  __threadfence_system();
  const bool bad = timed_out != 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += slot_data[(long long)r * row_stride + i];
    out[i] = bad ? __longlong_as_double(0x7ff8000000000000ll) : s;
  }
}

}  // namespace peer
}  // namespace xm

XM_DEFINE_SEED_EPOCH_SLOT(peer_exchange)

using namespace xm;

extern "C" int xm_peer_allreduce_f64(const double* x, double* out, int64_t n, const void* const* data_dst,
                                     const void* const* flag_dst, int n_peers, const double* slot_data,
                                     const uint64_t* slot_flags, int64_t row_stride, uint64_t seq, void* stream) {
  if (!x || !out || !data_dst || !flag_dst || !slot_data || !slot_flags || n <= 0 || n > row_stride || n_peers < 1 ||
      n_peers > 8 || seq == 0)
    return XM_ERR_INVALID;
  peer::Ptrs pp{};
  for (int r = 0; r < n_peers; ++r) {
    if (!data_dst[r] || !flag_dst[r]) return XM_ERR_INVALID;
    pp.data_dst[r] = (double*)data_dst[r];
    pp.flag_dst[r] = (unsigned long long*)flag_dst[r];
  }
  peer::peer_allreduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(x, out, (int)n, pp, n_peers, slot_data,
                                                                 (const unsigned long long*)slot_flags, (int)row_stride,
                                                                 (unsigned long long)seq);
  return check_launch();
}
