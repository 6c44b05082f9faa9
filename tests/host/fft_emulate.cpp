// Host emulation of multimodal_eeg_fmri_b200/csrc/spectral.cu's per-row band-power algorithm.
// Test infrastructure only: built with g++ by tests/test_spectral_core_host.py to check the
// Stockham index algebra, butterflies and real-FFT unpacking of csrc/spectral_core.cuh against
// numpy.fft on the CPU (the CUDA kernel runs the same helpers, one lane per butterfly).
#include <cmath>
#include <cstdint>
#include <vector>

#include "../../multimodal_eeg_fmri_b200/csrc/spectral_core.cuh"

using namespace xm;

template <int R>
static void run_pass(int N2, int Ns, const float* are, const float* aim, float* bre, float* bim, const float* tw_re,
                     const float* tw_im) {
  const int NB = N2 / R;
  for (int j = 0; j < NB; ++j) {
    float re[R], im[R];
    for (int r = 0; r < R; ++r) {
      const int i = fft::pad_idx(j + r * NB);
      re[r] = are[i];
      im[r] = aim[i];
    }
    fft::twiddle_and_butterfly<R>(re, im, j % Ns, Ns, N2, tw_re, tw_im);
    const int b = fft::scatter_base(j, Ns, R);
    for (int r = 0; r < R; ++r) {
      const int i = fft::pad_idx(b + r * Ns);
      bre[i] = re[r];
      bim[i] = im[r];
    }
  }
}

extern "C" int emulate_bandpower_row(const float* x, int win, const float* taper, int nfft, const int* band_bins,
                                     int n_bands, float scale, float* power, float* z_out /* 2*N2 or null */) {
  const int N2 = nfft / 2;
  const int PL = fft::padded_len(N2);
  std::vector<float> tw_re(N2), tw_im(N2), a_re(PL), a_im(PL), b_re(PL), b_im(PL);
  for (int m = 0; m < N2; ++m) {
    const double ang = -2.0 * M_PI * (double)m / (double)N2;
    tw_re[m] = (float)std::cos(ang);
    tw_im[m] = (float)std::sin(ang);
  }
  // pass 0 input: tapered, zero padded, packed z[n] = x[2n] + i x[2n+1] (stored in A as if gathered)
  for (int n = 0; n < N2; ++n) {
    float x0 = 0.f, x1 = 0.f;
    if (2 * n < win) x0 = x[2 * n] * taper[2 * n];
    if (2 * n + 1 < win) x1 = x[2 * n + 1] * taper[2 * n + 1];
    a_re[fft::pad_idx(n)] = x0;
    a_im[fft::pad_idx(n)] = x1;
  }
  float *sr = a_re.data(), *si = a_im.data(), *dr = b_re.data(), *di = b_im.data();
  int Ns = 1;
  while (Ns < N2) {
    const int R = fft::radix_for(N2 / Ns);
    if (R == 8) run_pass<8>(N2, Ns, sr, si, dr, di, tw_re.data(), tw_im.data());
    else if (R == 4) run_pass<4>(N2, Ns, sr, si, dr, di, tw_re.data(), tw_im.data());
    else run_pass<2>(N2, Ns, sr, si, dr, di, tw_re.data(), tw_im.data());
    Ns *= R;
    std::swap(sr, dr);
    std::swap(si, di);
  }
  const float* zre = sr;
  const float* zim = si;
  if (z_out)
    for (int k = 0; k < N2; ++k) {
      z_out[2 * k] = zre[fft::pad_idx(k)];
      z_out[2 * k + 1] = zim[fft::pad_idx(k)];
    }
  for (int b = 0; b < n_bands; ++b) {
    float acc = 0.f;
    for (int k = band_bins[2 * b]; k < band_bins[2 * b + 1]; ++k) {
      const int k1 = (k == N2) ? 0 : k;
      const int k2 = (k == 0 || k == N2) ? 0 : N2 - k;
      const float ar = zre[fft::pad_idx(k1)], ai = zim[fft::pad_idx(k1)];
      const float br = zre[fft::pad_idx(k2)], bi = -zim[fft::pad_idx(k2)];
      const float er = 0.5f * (ar + br), ei = 0.5f * (ai + bi);
      const float d_r = 0.5f * (ar - br), d_i = 0.5f * (ai - bi);
      const double ang = -2.0 * M_PI * (double)k / (double)nfft;
      const float co = (float)std::cos(ang), s = (float)std::sin(ang);
      const float xr = er + (co * d_i + s * d_r);
      const float xi = ei + (s * d_i - co * d_r);
      float pw = xr * xr + xi * xi;
      if (k != 0 && k != N2) pw *= 2.0f;
      acc += pw;
    }
    power[b] = acc * scale;
  }
  return 0;
}
