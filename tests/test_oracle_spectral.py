"""Known-answer tests that pin the AUTHORED band-power / windowing oracle (no reference
implementation exists -- SURVEY.md section 0): pure tone, Parseval, white noise, scipy cross-check."""
import numpy as np
import pytest
from scipy import signal

from oracle import spectral as osp


def test_window_indices_exact():
    starts, rec, lab, sub = osp.window_indices(3, 2000, 500, 250, rec_labels=[0, 1, 1], rec_subjects=[11, 12, 13])
    nw = (2000 - 500) // 250 + 1
    assert nw == 7 and starts.dtype == np.int64 and len(starts) == 3 * nw
    assert starts.tolist() == [w * 250 for w in range(nw)] * 3
    assert rec.tolist() == [r for r in range(3) for _ in range(nw)]
    assert lab.tolist() == [l for l in (0, 1, 1) for _ in range(nw)]
    assert sub.tolist() == [s for s in (11, 12, 13) for _ in range(nw)]
    # ragged / edge cases
    assert osp.n_windows(499, 500, 250) == 0
    assert osp.n_windows(500, 500, 250) == 1
    assert osp.n_windows(1023, 1024, 512) == 0 and osp.n_windows(1536, 1024, 512) == 2


def test_gather_matches_slicing():
    rng = np.random.default_rng(0)
    rec = rng.standard_normal((2, 3, 1000)).astype(np.float32)
    w = osp.gather_windows(rec, 256, 100)
    assert w.shape == (2 * 8, 3, 256)
    assert np.array_equal(w[8 + 5], rec[1, :, 500:756])


def test_band_bins_half_open():
    # fs=1000, nfft=1024: bin width 0.9766 Hz; theta [4,8) -> k in [5, 9); alpha [8,13) -> [9, 14); beta [13,30) -> [14, 31)
    b = osp.band_bins([(4, 8), (8, 13), (13, 30)], 1024, 1000.0)
    assert b.tolist() == [5, 9, 9, 14, 14, 31]
    # bin exactly on an edge belongs to the upper band: fs=512, nfft=512 -> 1 Hz bins
    b = osp.band_bins([(4, 8), (8, 13)], 512, 512.0)
    assert b.tolist() == [4, 8, 8, 13]


def test_pure_tone_all_power_in_alpha():
    fs, n = 1000.0, 1024
    t = np.arange(n) / fs
    f0 = 10.0 * fs / 1000.0 * (1024 / 1024)  # 10 Hz
    x = np.sin(2 * np.pi * 10.0 * t)[None, :]
    p = osp.band_power(x, fs)[0]
    assert p[1] / p.sum() > 0.999  # alpha
    # total power of a unit sine is 0.5 (Hann leakage stays inside alpha)
    assert abs(p[1] - 0.5) < 5e-3


def test_parseval_full_band():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((4, 512))
    n = 512
    p = osp.band_power(x, 256.0, bands=[(0.0, 129.0)], nfft=n, taper=np.ones(n))[:, 0]
    assert np.allclose(p, (x ** 2).mean(-1), rtol=1e-10)


def test_white_noise_proportional_to_bandwidth():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((4000, 1024))
    p = osp.band_power(x, 1000.0).mean(0)
    bins = osp.band_bins(list(osp.DEFAULT_BANDS.values()), 1024, 1000.0)
    widths = np.array([bins[1] - bins[0], bins[3] - bins[2], bins[5] - bins[4]], dtype=float)
    ratio = p / widths
    assert np.allclose(ratio / ratio.mean(), 1.0, atol=0.05)


@pytest.mark.parametrize("win,nfft,fs", [(1024, 1024, 1000.0), (500, 512, 250.0), (100, 128, 128.0)])
def test_matches_scipy_periodogram(win, nfft, fs):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((5, win))
    f, P = signal.periodogram(x, fs=fs, window="hann", nfft=nfft, detrend=False, scaling="density", axis=-1)
    bands = [(4, 8), (8, 13), (13, 30)]
    ref = np.stack([P[:, (f >= lo) & (f < hi)].sum(-1) * fs / nfft for lo, hi in bands], -1)
    got = osp.band_power(x, fs, bands=bands, nfft=nfft)
    assert np.allclose(got, ref, rtol=1e-9, atol=0)


def test_pw_layout_row_is_c_times_F_plus_f():
    power = np.arange(5 * 4 * 3, dtype=np.float64).reshape(5, 4, 3)  # (T, C, F)
    pw = osp.pw_layout(power)
    assert pw.shape == (12, 5)
    assert pw[2 * 3 + 1, 4] == power[4, 2, 1]


def test_normalize_and_label_rules():
    x = np.array([[1.0, 2.0], [3.0, 6.0]])
    z = osp.normalize_modality(x)
    assert abs(z.mean()) < 1e-12 and abs(z.std() - 1.0) < 1e-6
    # eeg_data_utils.py:42 precedence quirk
    assert [osp.binarise_score(s) for s in (1, 2, 3, 5)] == [0, 0, 1, 1]
    assert [osp.binarise_score(s, binary=False) for s in (1, 2, 3, 5)] == [0, 0, 3, 5]
    assert osp.binarise_score(2, threshold=1) == 1  # run_training_lite.py:290-291 variant
