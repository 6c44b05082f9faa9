// Radix-8/4/2 Stockham FFT building blocks shared by the CUDA band-power kernel and a
// host-side emulation used by the CPU test-suite (tests/test_spectral_core_host.py builds this
// header with g++ and checks the index algebra against numpy.fft).
//
// One "warp" (32 lanes) transforms one row of N2 = nfft/2 complex points z[n] = x[2n] + i x[2n+1]
// (real-input packing).  Every pass reads R points per butterfly, applies the Stockham twiddle
// W_{Ns*R}^{k*r}, runs an in-register radix-R butterfly and scatters to autosorted positions, so
// the output is in natural order with no bit-reversal pass.
#pragma once

#if defined(__CUDACC__)
#define XM_HD __host__ __device__ __forceinline__
#else
#define XM_HD inline
#endif

namespace xm {
namespace fft {

// padded index into the per-warp exchange arrays (keeps pass-0 scatters and all gathers
// bank-conflict free)
XM_HD int pad_idx(int i) { return i + (i >> 5); }
XM_HD int padded_len(int n) { return n + (n >> 5) + 1; }

XM_HD void cmul(float& ar, float& ai, float br, float bi) {
  const float r = ar * br - ai * bi;
  const float i = ar * bi + ai * br;
  ar = r;
  ai = i;
}

template <int R>
struct Butterfly;

template <>
struct Butterfly<2> {
  static XM_HD void run(float* re, float* im) {
    const float r0 = re[0] + re[1], i0 = im[0] + im[1];
    const float r1 = re[0] - re[1], i1 = im[0] - im[1];
    re[0] = r0; im[0] = i0; re[1] = r1; im[1] = i1;
  }
};

template <>
struct Butterfly<4> {
  static XM_HD void run(float* re, float* im) {
    const float t0r = re[0] + re[2], t0i = im[0] + im[2];
    const float t1r = re[0] - re[2], t1i = im[0] - im[2];
    const float t2r = re[1] + re[3], t2i = im[1] + im[3];
    // (a1 - a3) * (-i) = (y, -x)
    const float dr = re[1] - re[3], di = im[1] - im[3];
    const float t3r = di, t3i = -dr;
    re[0] = t0r + t2r; im[0] = t0i + t2i;
    re[1] = t1r + t3r; im[1] = t1i + t3i;
    re[2] = t0r - t2r; im[2] = t0i - t2i;
    re[3] = t1r - t3r; im[3] = t1i - t3i;
  }
};

template <>
struct Butterfly<8> {
  static XM_HD void run(float* re, float* im) {
    float er[4] = {re[0], re[2], re[4], re[6]}, ei[4] = {im[0], im[2], im[4], im[6]};
    float orr[4] = {re[1], re[3], re[5], re[7]}, oi[4] = {im[1], im[3], im[5], im[7]};
    Butterfly<4>::run(er, ei);
    Butterfly<4>::run(orr, oi);
    const float h = 0.70710678118654752440f;
    // O[k] *= W8^k : W8^1 = (h,-h), W8^2 = (0,-1), W8^3 = (-h,-h)
    {
      const float r = (orr[1] + oi[1]) * h, i = (oi[1] - orr[1]) * h;
      orr[1] = r; oi[1] = i;
    }
    {
      const float r = oi[2], i = -orr[2];
      orr[2] = r; oi[2] = i;
    }
    {
      const float r = (oi[3] - orr[3]) * h, i = -(orr[3] + oi[3]) * h;
      orr[3] = r; oi[3] = i;
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 4; ++k) {
      re[k] = er[k] + orr[k];     im[k] = ei[k] + oi[k];
      re[k + 4] = er[k] - orr[k]; im[k + 4] = ei[k] - oi[k];
    }
  }
};

// radix used by the pass whose remaining length is `rem`
XM_HD constexpr int radix_for(int rem) { return (rem % 8 == 0) ? 8 : ((rem % 4 == 0) ? 4 : 2); }

// One Stockham butterfly of pass (Ns, R) for butterfly index j in [0, N2/R):
// inputs v[r] = in[j + r*N2/R]; twiddle base w1 = W_{Ns*R}^{j % Ns} taken from the table
// tw[m] = exp(-2*pi*i*m/N2); outputs to out[(j/Ns)*Ns*R + (j%Ns) + r*Ns].
template <int R>
XM_HD void twiddle_and_butterfly(float* re, float* im, int k, int Ns, int N2, const float* tw_re, const float* tw_im) {
  if (Ns > 1) {
    const int m = k * (N2 / (Ns * R));
    const float w1r = tw_re[m], w1i = tw_im[m];
    float wr = w1r, wi = w1i;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 1; r < R; ++r) {
      cmul(re[r], im[r], wr, wi);
      if (r + 1 < R) cmul(wr, wi, w1r, w1i);
    }
  }
  Butterfly<R>::run(re, im);
}

XM_HD int scatter_base(int j, int Ns, int R) { return (j / Ns) * Ns * R + (j % Ns); }

}  // namespace fft
}  // namespace xm
