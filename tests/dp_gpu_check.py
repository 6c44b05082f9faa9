"""Multi-GPU check of the data-parallel paired step (run under torchrun on >= 2 B200s of one node):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dp_gpu_check.py

Rank r trains on rows [r*Bl, (r+1)*Bl) of a global batch; the sharded run (SyncBN partial sums, peer-memory
fused all-gather + InfoNCE, flat gradient all-reduce) must reproduce the single-process global-batch oracle:
same global loss and same parameters after the steps.  Runs both the NVLink peer-memory path and the NCCL
all-gather path and prints one line per path."""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (HERE, os.path.dirname(HERE)):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    from multimodal_eeg_fmri_b200 import functional as XF
    from multimodal_eeg_fmri_b200 import synthetic
    from multimodal_eeg_fmri_b200.training import PairedBridgeModel, PairedTrainer, init_distributed
    from oracle import paired_step as ps

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    init_distributed("nccl")
    ok = True
    # Bl = 128: remote tiles read inside the GEMMs; Bl = 640: gather-once through the peer mappings
    for Bl, steps, peer in ((128, 2, True), (128, 2, False), (640, 1, True)):
        eeg, roi, conn = synthetic.paired_batch(Bl * world, 8, 64, 12, 20, seed=5)
        sl = slice(rank * Bl, (rank + 1) * Bl)
        XF.set_parallel_context(XF.ParallelContext(group=None, sync_bn=True, peer_memory=peer))
        torch.manual_seed(7)
        model = PairedBridgeModel(eeg_channels=8, n_roi=12, eeg_hidden=32, fmri_hidden=16, bridge_dim=32, dropout=0.0,
                                  fmri_dropout=0.0, encoder="lite")
        P = {k: v.detach().clone() for k, v in model.state_dict().items()}
        trainer = PairedTrainer(model.cuda().train())
        losses, norms = [], []
        for _ in range(steps):
            loss = trainer.step(eeg[sl].cuda(), roi[sl].cuda(), conn[sl].cuda()).clone()
            dist.all_reduce(loss)
            losses.append(float(loss))
            norms.append(float(trainer.last_grad_norm))  # norm of the ALL-REDUCED gradient (both buckets)
        used_peer = peer and not XF._PeerShards._failed
        if rank == 0:
            state, want, want_norm = {}, [], []
            for _ in range(steps):
                l, nrm = ps.paired_train_step(P, state, eeg, roi, conn, 0.07, "lite")
                want.append(float(l))
                want_norm.append(float(nrm))
            lerr = max(abs(a - b) / abs(b) for a, b in zip(losses, want))
            nerr = max(abs(a - b) / abs(b) for a, b in zip(norms, want_norm))
            keys = set(ps.trainable_keys(P)) - set(ps.bias_before_batchnorm_keys(P))
            sd = model.state_dict()
            perr = max(float((sd[k].cpu() - P[k]).abs().max()) for k in keys)
            # AdamW steps of lr 1e-4: a sign flip of a ~zero gradient moves a weight by 2*lr per step
            # (AdamW normalises the step, so the parameters alone would not notice a gradient bucket that missed its
            # all-reduce: the pre-clip norm of the reduced gradient does)
            good = lerr < 1e-3 and perr < 2.25e-4 * steps and nerr < 2e-3
            ok = ok and good
            path = ("peer-memory/in-GEMM" if Bl <= 512 else "peer-memory/gather-once") if used_peer else "nccl-all-gather"
            print(f"dp_gpu_check world={world} local_batch={Bl} path={path} loss_rel_err={lerr:.2e} grad_norm_rel_err={nerr:.2e} "
                  f"param_max_abs_err={perr:.2e} {'OK' if good else 'FAIL'}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
