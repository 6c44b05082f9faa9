#!/usr/bin/env python
"""Measurements for the SURVEY 8f rows 3-4 (run on a B200; prints one JSON line per measurement):

  * attribution: BridgeIntegratedGradients with all n_steps interpolation points in ONE batch (this repo) against
    the reference's schedule (n_steps sequential forward/backward passes through the same model), samples/s;
  * formats: XMSHARD1 file -> page-locked buffers -> device, GB/s (page cache warm), and the reference-style host
    stage it replaces (pandas CSV parse of the same ROI series), MB/s.

    python tools/widen_bench.py [--subjects 4096] [--n-steps 50] [--rows 2048]
"""
import argparse
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from multimodal_eeg_fmri_b200 import bridge_utils as bu, shards  # noqa: E402


def _sequential_ig(model, eeg, fmri, n_steps):
    """The reference's loop (bridge_utils.py:207-224) on the device: one forward/backward per alpha."""
    model.eval()
    tc = None
    ge, gf = torch.zeros_like(eeg), torch.zeros_like(fmri)
    for alpha in np.linspace(0, 1, n_steps):
        a, b, tc = bu._class_logit_gradients(model, float(alpha) * eeg, float(alpha) * fmri, tc)
        ge += a
        gf += b
    return (eeg * ge / n_steps).abs(), (fmri * gf / n_steps).abs()


def attribution(args):
    torch.manual_seed(0)
    model = bu.EEGfMRIBridgeFusionNet().cuda().eval()
    eeg, fmri = torch.randn(args.subjects, 128, device="cuda"), torch.randn(args.subjects, 64, device="cuda")
    ig = bu.BridgeIntegratedGradients(model, "cuda", n_steps=args.n_steps)
    ref = _sequential_ig(model, eeg, fmri, args.n_steps)
    got = ig.compute(eeg, fmri)
    err = float((torch.from_numpy(got["eeg"]).cuda() - ref[0]).norm() / ref[0].norm())
    out = {}
    for name, fn in (("batched", lambda: ig.compute(eeg, fmri)), ("sequential", lambda: _sequential_ig(model, eeg, fmri, args.n_steps))):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            fn()
        torch.cuda.synchronize()
        out[name] = args.subjects * args.reps / (time.perf_counter() - t0)
    print(json.dumps({"metric": "integrated-gradients samples/sec (bridge model)", "subjects": args.subjects, "n_steps": args.n_steps,
                      "batched": round(out["batched"], 1), "sequential": round(out["sequential"], 1),
                      "speedup": round(out["batched"] / out["sequential"], 2), "rel_diff_batched_vs_sequential": err,
                      "note": "wall clock incl. the device->host copy of the attributions (the API returns NumPy arrays)"}), flush=True)


def formats(args):
    g = np.random.default_rng(0)
    arrays = {"eeg": g.standard_normal((args.rows, 64, 500)).astype(np.float32), "roi": g.standard_normal((args.rows, 100, 200)).astype(np.float32),
              "conn": g.standard_normal((args.rows, 40000)).astype(np.float32)}
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "paired.xms")
        size = shards.write_shard(path, arrays)
        sh = shards.Shard(path)
        sh.to_device(list(arrays))  # warm-up (page cache, pinned allocations)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            dev = sh.to_device(list(arrays))
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.reps
        ok = bool(torch.equal(dev["conn"].cpu(), torch.from_numpy(arrays["conn"])))
        # the host stage this replaces: one CSV per subject, parsed with pandas (fmri_utils.py:135-139)
        import pandas as pd
        csv = os.path.join(d, "one.csv")
        pd.DataFrame(arrays["roi"][0]).to_csv(csv, index=False)
        t1 = time.perf_counter()
        for _ in range(20):
            pd.read_csv(csv).values.astype(np.float32)
        csv_s = (time.perf_counter() - t1) / 20
    print(json.dumps({"metric": "shard file -> pinned -> device", "bytes": size, "rows": args.rows, "gb_per_s": round(size / dt / 1e9, 2),
                      "bit_exact": ok, "csv_parse_mb_per_s_of_fp32": round(arrays["roi"][0].nbytes / csv_s / 1e6, 1),
                      "note": "page cache warm; includes the file read into page-locked memory and the H2D copy"}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--subjects", type=int, default=4096)
    ap.add_argument("--n-steps", type=int, default=50)
    ap.add_argument("--rows", type=int, default=2048)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("widen_bench.py needs a CUDA device")
    attribution(args)
    formats(args)


if __name__ == "__main__":
    main()
