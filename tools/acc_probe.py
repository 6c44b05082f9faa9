"""Accuracy probe: paired-step gradients vs the CPU oracle with the long connectivity projection in single-pass
tf32 vs the 3-pass mode (decides functional._PRECISE_MAX_K).  python tools/acc_probe.py  (needs a B200)."""
import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from multimodal_eeg_fmri_b200 import synthetic, functional as XF
from multimodal_eeg_fmri_b200.training import PairedBridgeModel
from oracle import paired_step as ps
from conftest import rel_err
for maxk in (4096, 1 << 30):  # single-pass above 4096 vs always 3-pass
    XF._PRECISE_MAX_K = maxk
    torch.manual_seed(1)
    m = PairedBridgeModel(64, 200, None, 96, 64, 128, 0.0, 0.0, "lite")
    P = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    eeg, roi, conn = synthetic.paired_batch(64, 64, 500, 200, 100, seed=2)
    m = m.cuda().train()
    loss = m(eeg.cuda(), roi.cuda(), conn.cuda()); loss.backward()
    oloss, og = ps.paired_loss_and_grads(P, eeg, roi, conn, 0.07, "lite")
    named = dict(m.named_parameters())
    print("precise_max_k", maxk, "loss rel", abs(float(loss)-float(oloss))/float(oloss))
    for k in ("fmri_net.connectivity_encoder.encoder.0.weight","fmri_net.connectivity_encoder.encoder.4.weight","fmri_net.activation_encoder.encoder.0.weight","fmri_net.fusion.0.weight","bridge.fmri_proj.0.weight","eeg_encoder.conv_layers.0.weight"):
        print("   ", k, "%.2e" % rel_err(named[k].grad, og[k]))
