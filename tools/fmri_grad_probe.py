#!/usr/bin/env python
"""Probe: per-stage forward / backward error of the fMRI net alone at a large batch vs a float64 oracle run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from multimodal_eeg_fmri_b200 import fmri_utils, synthetic, functional as XF
from oracle import models as om

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-300))


torch.manual_seed(1)
m = fmri_utils.fMRIFusionNet(400, 40000, 64, 2, 0.0)
P = {k: v.detach().clone().double().requires_grad_(True) if v.is_floating_point() else v.clone() for k, v in m.state_dict().items()}
_, roi, conn = synthetic.paired_batch(B, 64, 500, 200, 100, seed=42)
gf = torch.randn(B, 64)
gf = gf - 0.9 * gf.mean(0, keepdim=True) + 0.5
m = m.cuda().train()
act = fmri_utils.aggregate_roi_timeseries(roi.cuda(), "both")
actd = om.roi_meanstd(roi.double())
print("act features", rel(act, actd))
# GPU, stage by stage
enc = m.connectivity_encoder
mods = list(enc.children()) if not hasattr(enc, "encoder") else None
a = m.activation_encoder(act); a.retain_grad()
c = m.connectivity_encoder(conn.cuda()); c.retain_grad()
w = torch.softmax(torch.stack([m.activation_weight, m.connectivity_weight]), dim=0)
comb = torch.cat([a * w[0], c * w[1]], dim=1); comb.retain_grad()
fused = XF.linear_bn_act(comb, m.fusion[0], m.fusion[1], "relu", 0.0, True)
(fused * gf.cuda()).sum().backward()
# oracle, stage by stage
ad = om.fmri_mlp_encoder(P, "activation_encoder.", actd); ad.retain_grad()
x0 = conn.double()
y0 = om._lin(P, "connectivity_encoder.encoder.0.", x0)
h0 = F.relu(om._bn(P, "connectivity_encoder.encoder.1.", y0, True))
cd = F.relu(om._bn(P, "connectivity_encoder.encoder.5.", om._lin(P, "connectivity_encoder.encoder.4.", h0), True)); cd.retain_grad()
wd = torch.softmax(torch.stack([P["activation_weight"], P["connectivity_weight"]]), dim=0)
combd = torch.cat([ad * wd[0], cd * wd[1]], dim=1); combd.retain_grad()
fd = F.relu(om._bn(P, "fusion.1.", om._lin(P, "fusion.0.", combd), True))
(fd * gf.double()).sum().backward()
print("y0: mean|y| %.3e batch-std %.3e" % (float(y0.abs().mean()), float(y0.std(0).mean())))
from multimodal_eeg_fmri_b200 import ops
y0g = ops.linear_fwd_precise(conn.cuda(), m.connectivity_encoder.encoder[0].weight.detach(), m.connectivity_encoder.encoder[0].bias.detach())
print("y0 err rel-to-norm %.3e, err/std %.3e" % (rel(y0g, y0), float(((y0g.cpu().double() - y0.detach()) / y0.detach().std(0)).abs().mean())))
print("fwd a", rel(a, ad), "c", rel(c, cd), "fused", rel(fused, fd), "flips c", int(((c > 0).cpu() != (cd > 0)).sum()), "flips fused", int(((fused > 0).cpu() != (fd > 0)).sum()))
print("bwd dcomb", rel(comb.grad, combd.grad), "da", rel(a.grad, ad.grad), "dc", rel(c.grad, cd.grad))
for k, p in m.named_parameters():
    if p.grad is None or P[k].grad is None:
        continue
    g = P[k].grad
    if float(g.norm()) < 1e-9:
        continue
    print(f"{k:55s} {float((p.grad.cpu().double() - g).norm() / (g.norm() + 1e-300)):.3e}")
