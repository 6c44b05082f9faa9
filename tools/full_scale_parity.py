#!/usr/bin/env python
"""Direct parity of the paired step against the CPU oracle at a LARGE batch (default 2048 paired samples at
the BASELINE shapes: 64 ch x 500 samples, 200 ROI x 100 TR, conn 40 000; v4 encoder; dropout 0): loss and
parameter gradients.  The CPU oracle needs about a minute on 16 host cores.  Writes one JSON object.

    python tools/full_scale_parity.py [--batch 2048] [--out profiles/r1_full_scale_parity.json]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from multimodal_eeg_fmri_b200 import synthetic  # noqa: E402
from multimodal_eeg_fmri_b200.training import PairedBridgeModel  # noqa: E402
from oracle import paired_step as ps  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


class _RoundTf32(torch.autograd.Function):
    """Round to nearest-even tf32 (10 explicit mantissa bits), identity gradient."""

    @staticmethod
    def forward(ctx, t):
        b = t.float().contiguous().view(torch.int32)
        b = (b + 0xFFF + ((b >> 13) & 1)) & ~0x1FFF
        return b.view(torch.float32).to(t.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


def tf32_floor(P64, eeg, roi, conn, g64, zero):
    """Error floor of single-pass tf32 in the EEG encoder's FORWARD pass alone: the float64 oracle with only
    the operands of the EEG encoder's convs / linears / attention products rounded to tf32 (exact accumulation,
    exact backward).  Whatever error this run shows against the plain float64 run is inherent to the precision
    policy (the 1/tau = 14x logit scale amplifies the embedding noise into the softmax), not to a kernel."""
    from oracle import models as om
    om.OPERAND_ROUNDING = lambda pre, t: _RoundTf32.apply(t) if pre.startswith("eeg_encoder.") else t
    try:
        lt, gt = ps.paired_loss_and_grads(P64, eeg.double(), roi.double(), conn.double(), 0.07, "v4")
    finally:
        om.OPERAND_ROUNDING = None
    return lt, {k: rel(gt[k], g) for k, g in g64.items() if k not in zero}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--out", default="")
    ap.add_argument("--fp64", action="store_true", help="also run the oracle in float64 and report errors against it")
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(42)
    m = PairedBridgeModel(64, 200, None, 128, 64, 128, 0.0, 0.0, "v4")
    P = {k: v.detach().clone() for k, v in m.state_dict().items()}
    eeg, roi, conn = synthetic.paired_batch(a.batch, 64, 500, 200, 100, seed=42)
    m = m.cuda().train()
    t0 = time.time()
    loss = m(eeg.cuda(), roi.cuda(), conn.cuda())
    loss.backward()
    torch.cuda.synchronize()
    t_gpu = time.time() - t0
    t0 = time.time()
    oloss, og = ps.paired_loss_and_grads(P, eeg, roi, conn, 0.07, "v4")
    t_cpu = time.time() - t0
    named = dict(m.named_parameters())
    zero = set(ps.bias_before_batchnorm_keys(P))
    errs = {k: rel(named[k].grad, g) for k, g in og.items() if k not in zero}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:8]
    extra = {}
    if a.fp64:  # the fp32 oracle has rounding error of its own: measure both against an fp64 run of the oracle
        P64 = {k: (v.double() if v.is_floating_point() else v) for k, v in P.items()}
        l64, g64 = ps.paired_loss_and_grads(P64, eeg.double(), roi.double(), conn.double(), 0.07, "v4")
        e_gpu = {k: rel(named[k].grad, g) for k, g in g64.items() if k not in zero}
        e_cpu = {k: rel(og[k], g) for k, g in g64.items() if k not in zero}
        w = sorted(e_gpu.items(), key=lambda kv: -kv[1])[:8]
        lt, e_floor = tf32_floor(P64, eeg, roi, conn, g64, zero)
        extra = {"vs_fp64": {"loss_rel_err_gpu": abs(float(loss) - float(l64)) / abs(float(l64)),
                             "loss_rel_err_fp32_oracle": abs(float(oloss) - float(l64)) / abs(float(l64)),
                             "grad_rel_err_gpu_median": sorted(e_gpu.values())[len(e_gpu) // 2], "grad_rel_err_gpu_max": w[0][1],
                             "grad_rel_err_fp32_oracle_median": sorted(e_cpu.values())[len(e_cpu) // 2],
                             "grad_rel_err_fp32_oracle_max": max(e_cpu.values()),
                             "loss_rel_err_tf32_floor": abs(float(lt) - float(l64)) / abs(float(l64)),
                             "grad_rel_err_tf32_floor_median": sorted(e_floor.values())[len(e_floor) // 2],
                             "grad_rel_err_tf32_floor_max": max(e_floor.values()),
                             "worst_gpu": {k: {"gpu": v, "fp32_oracle": e_cpu[k], "tf32_floor": e_floor[k]} for k, v in w}}}
    res = {"batch": a.batch, "encoder": "v4", "loss_gpu": float(loss), "loss_oracle": float(oloss),
           "loss_rel_err": abs(float(loss) - float(oloss)) / abs(float(oloss)),
           "grad_rel_err_median": sorted(errs.values())[len(errs) // 2], "grad_rel_err_max": worst[0][1],
           "grad_rel_err_worst": dict(worst), "n_param_tensors": len(errs),
           "seconds_gpu_first_call": round(t_gpu, 2), "seconds_cpu_oracle": round(t_cpu, 2), "cpu_threads": torch.get_num_threads(),
           **extra}
    print(json.dumps(res, indent=1))
    if a.out:
        open(a.out, "w").write(json.dumps(res, indent=1) + "\n")


if __name__ == "__main__":
    main()
