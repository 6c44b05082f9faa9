#!/usr/bin/env python
"""BASELINE config 5: EEG preprocessing sweep -- theta/alpha/beta band power of 128-channel, 1 kHz
recordings over ~1M windows (win 1024, hop 512), streamed in chunks that fit HBM; windows are read in
place from the recordings (never materialised).  Under torchrun every rank sweeps its own share of the
windows (no collective: the path shards by recording).  Prints one JSON line (rank 0).

    python tools/bandpower_sweep.py [--windows 1048576] [--chunk-recordings 256]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from multimodal_eeg_fmri_b200 import eeg_data_utils as edu  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--windows", type=int, default=1 << 20)
    ap.add_argument("--chunk-recordings", type=int, default=256)
    ap.add_argument("--channels", type=int, default=128)
    ap.add_argument("--win", type=int, default=1024)
    ap.add_argument("--hop", type=int, default=512)
    ap.add_argument("--windows-per-recording", type=int, default=64)
    ap.add_argument("--paths", default="auto", help='comma list of "auto" | "dft" | "fft": one JSON line each')
    a = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl")
    n = a.win + (a.windows_per_recording - 1) * a.hop
    per_chunk = a.chunk_recordings * a.windows_per_recording
    chunks = max(1, a.windows // per_chunk // world)  # this rank's share
    g = torch.Generator(device="cuda").manual_seed(42 + rank)
    t = torch.arange(n, device="cuda", dtype=torch.float32) / 1000.0
    tone = sum(torch.sin(2 * torch.pi * f * t) for f in (6.0, 10.0, 20.0))
    bufs = [torch.randn(a.chunk_recordings, a.channels, n, device="cuda", generator=g) + tone for _ in range(2)]
    for path in a.paths.split(","):
        for i in range(3):
            edu.band_power(bufs[i & 1], 1000.0, a.win, a.hop, path=path)  # warm-up
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        # Timed region: the band-power launches only; every chunk's (windows, C, 3) features are written to HBM.
        # The checksum below is taken from the last two chunks AFTER the timed region (torch's strided sum over a
        # (16384, 128, 3) tensor costs about as much as the kernel itself and is not part of the path).
        e0.record()
        outs = [None, None]
        for c in range(chunks):
            outs[c & 1] = edu.band_power(bufs[c & 1], 1000.0, a.win, a.hop, path=path)  # chunks alternate (> L2 each)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        live = [o for o in outs if o is not None]
        acc = sum(o.double().sum((0, 1)) for o in live) / (len(live) * per_chunk * a.channels)
        if rank == 0:
            nwin = chunks * per_chunk * world
            by = nwin * (a.channels * a.win * 4 + a.channels * 3 * 4)
            sec = float(ms) * 1e-3
            print(json.dumps({"metric": "EEG band-power windows/sec", "value": round(nwin / sec, 1), "unit": "windows/s", "n_gpus": world,
                              "path": path, "windows": nwin, "ms_total": round(float(ms), 2), "channels": a.channels, "win": a.win,
                              "hop": a.hop, "algorithmic_gbs_per_gpu": round(by / sec / 1e9 / world, 1),
                              "mean_band_power": [round(float(v), 6) for v in acc.tolist()],
                              "note": "kernel launches only (features written to HBM); inputs resident in HBM, two alternating chunks"}),
                  flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
