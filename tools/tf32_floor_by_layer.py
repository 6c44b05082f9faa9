#!/usr/bin/env python
"""Which layers' tf32 operand rounding produces the parameter-gradient error floor?  (CPU only, float64 oracle.)

Runs the float64 oracle of the paired step with operands rounded to tf32 ONLY in a chosen subset of the EEG
encoder's contractions (exact accumulation, exact backward) and reports, per subset, the norm-wise relative error
of every parameter gradient against the plain float64 run.  Decides which layers must run in the fp32-accurate
3-pass mode for every gradient to meet the 1e-3 north-star tolerance.

    python tools/tf32_floor_by_layer.py [--batch 512] [--out profiles/r2_tf32_floor_by_layer.json]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from multimodal_eeg_fmri_b200 import synthetic  # noqa: E402
from multimodal_eeg_fmri_b200.training import PairedBridgeModel  # noqa: E402
from oracle import models as om  # noqa: E402
from oracle import paired_step as ps  # noqa: E402
from tools.full_scale_parity import _RoundTf32, rel  # noqa: E402

SUBSETS = {
    "all": lambda p: p.startswith("eeg_encoder."),
    "conv0": lambda p: p.startswith("eeg_encoder.conv_layers.0."),
    "conv4": lambda p: p.startswith("eeg_encoder.conv_layers.4."),
    "conv9": lambda p: p.startswith("eeg_encoder.conv_layers.9."),
    "tail": lambda p: p.startswith("eeg_encoder.") and not p.startswith("eeg_encoder.conv_layers."),
    "all_but_conv0_conv4": lambda p: p.startswith("eeg_encoder.") and not p.startswith(("eeg_encoder.conv_layers.0.", "eeg_encoder.conv_layers.4.")),
    "all_but_convs": lambda p: p.startswith("eeg_encoder.") and not p.startswith("eeg_encoder.conv_layers."),
}


class _RoundGrad(torch.autograd.Function):
    """Identity whose BACKWARD rounds the incoming gradient to tf32 (the dy operand of dgrad / wgrad)."""

    @staticmethod
    def forward(ctx, t):
        return t.view_as(t)

    @staticmethod
    def backward(ctx, g):
        return _RoundTf32.apply(g)


def _bwd_rounded(fn, x, w, b):
    """fn(x, w, b) with exact forward VALUE but a backward whose three operands (dy, w in dgrad, x in wgrad) are
    tf32-rounded: the graph runs through fn(round(x), round(w)), the exact value is added back detached."""
    xr, wr = _RoundTf32.apply(x), _RoundTf32.apply(w)
    y = fn(xr, wr, b)
    y = y + (fn(x, w, b) - y).detach()
    return _RoundGrad.apply(y)


def install_backward_rounding(sel):
    """Model single-pass tf32 dgrad / wgrad in the convs and linears whose key prefix satisfies `sel`."""
    import torch.nn.functional as F
    conv0, lin0 = om._conv, om._lin

    def conv(P, pre, x):
        if not sel(pre):
            return conv0(P, pre, x)
        w = P[pre + "weight"]
        return _bwd_rounded(lambda a, ww, bb: F.conv1d(om._r(pre, a), om._r(pre, ww), bb, padding=ww.shape[-1] // 2), x, w,
                            P.get(pre + "bias"))

    def lin(P, pre, x):
        if not sel(pre):
            return lin0(P, pre, x)
        return _bwd_rounded(lambda a, ww, bb: F.linear(om._r(pre, a), om._r(pre, ww), bb), x, P[pre + "weight"], P.get(pre + "bias"))

    om._conv, om._lin = conv, lin
    return lambda: (setattr(om, "_conv", conv0), setattr(om, "_lin", lin0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bwd", default="", help="also round the backward operands of this subset's convs / linears")
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--out", default="")
    ap.add_argument("--subsets", default=",".join(SUBSETS))
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(42)
    m = PairedBridgeModel(64, 200, None, 128, 64, 128, 0.0, 0.0, "v4")
    P64 = {k: (v.detach().double() if v.is_floating_point() else v.detach().clone()) for k, v in m.state_dict().items()}
    eeg, roi, conn = (t.double() for t in synthetic.paired_batch(a.batch, 64, 500, 200, 100, seed=42))
    zero = set(ps.bias_before_batchnorm_keys(P64))
    _, g64 = ps.paired_loss_and_grads(P64, eeg, roi, conn, 0.07, "v4")
    res = {"batch": a.batch}
    for name in a.subsets.split(","):
        sel = SUBSETS[name]
        om.OPERAND_ROUNDING = lambda pre, t, sel=sel: _RoundTf32.apply(t) if sel(pre) else t
        undo = install_backward_rounding(SUBSETS[a.bwd]) if a.bwd else (lambda: None)
        try:
            _, g = ps.paired_loss_and_grads(P64, eeg, roi, conn, 0.07, "v4")
        finally:
            om.OPERAND_ROUNDING = None
            undo()
        errs = {k: rel(g[k], g64[k]) for k in g64 if k not in zero}
        worst = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
        name = name + ("+bwd:" + a.bwd if a.bwd else "")
        res[name] = {"median": sorted(errs.values())[len(errs) // 2], "max": worst[0][1], "worst": dict(worst),
                     "n_above_1e-3": sum(v > 1e-3 for v in errs.values())}
        print(name, json.dumps(res[name]), flush=True)
    if a.out:
        open(a.out, "w").write(json.dumps(res, indent=1) + "\n")


if __name__ == "__main__":
    main()
