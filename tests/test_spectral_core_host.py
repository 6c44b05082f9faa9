"""csrc/spectral_core.cuh (Stockham index algebra, radix-2/4/8 butterflies, real-FFT unpacking) compiled
for the host with g++ and checked against numpy.fft -- the CUDA band-power kernel runs the same helpers."""
import ctypes
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    out = tmp_path_factory.mktemp("fft") / "fft_emu.so"
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", str(out), str(ROOT / "tests" / "host" / "fft_emulate.cpp")], check=True)
    return ctypes.CDLL(str(out))


@pytest.mark.parametrize("nfft,win", [(64, 64), (64, 50), (128, 100), (256, 256), (512, 500), (1024, 1024), (1024, 1001), (2048, 2000)])
def test_host_emulation_matches_numpy(emu, nfft, win):
    rng = np.random.default_rng(nfft + win)
    x = rng.standard_normal(win).astype(np.float32)
    taper = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(win) / win)).astype(np.float32)
    n2 = nfft // 2
    bins = np.array([0, n2 + 1, 5, min(31, n2), 1, 2], dtype=np.int32)
    scale = np.float32(1.0 / (nfft * np.sum(taper.astype(np.float64) ** 2)))
    power = np.zeros(3, np.float32)
    z = np.zeros(2 * n2, np.float32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    emu.emulate_bandpower_row(p(x), win, p(taper), nfft, p(bins), 3, ctypes.c_float(scale), p(power), p(z))
    xw = x.astype(np.float64) * taper
    X = np.fft.rfft(xw, n=nfft)
    P = np.abs(X) ** 2
    P[1:-1] *= 2
    ref = np.array([P[bins[0]:bins[1]].sum(), P[bins[2]:bins[3]].sum(), P[bins[4]:bins[5]].sum()]) * scale
    zc = np.zeros(nfft)
    zc[:win] = xw
    zref = np.fft.fft(zc[0::2] + 1j * zc[1::2])
    assert np.abs((z[0::2] + 1j * z[1::2]) - zref).max() / np.abs(zref).max() < 2e-6
    assert np.all(np.abs(power - ref) / ref < 1e-5)
