"""Worker for the world_size-2 gloo test of the data-parallel paired step (CPU, fake device ops)."""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (HERE, os.path.dirname(HERE)):
    if p not in sys.path:
        sys.path.insert(0, p)


def make_model(encoder):
    from multimodal_eeg_fmri_b200.training import PairedBridgeModel
    torch.manual_seed(7)
    return PairedBridgeModel(eeg_channels=8, n_roi=12, eeg_hidden=32, fmri_hidden=16, bridge_dim=32, dropout=0.0,
                             fmri_dropout=0.0, encoder=encoder)


def run(rank, world, port, encoder, steps, out_dir):
    import fake_ops
    from multimodal_eeg_fmri_b200 import synthetic
    from multimodal_eeg_fmri_b200 import functional as XF
    from multimodal_eeg_fmri_b200.training import PairedTrainer

    torch.set_num_threads(1)
    fake_ops.install()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    XF.set_parallel_context(XF.ParallelContext(group=None, sync_bn=True))
    model = make_model(encoder).train()
    trainer = PairedTrainer(model)
    B = 16 // world
    eeg, roi, conn = synthetic.paired_batch(16, 8, 64, 12, 20, seed=5)
    sl = slice(rank * B, (rank + 1) * B)
    losses = []
    for _ in range(steps):
        loss = trainer.step(eeg[sl], roi[sl], conn[sl]).clone()
        dist.all_reduce(loss)  # the global loss is the sum of the per-rank shares
        losses.append(float(loss))
    if rank == 0:
        torch.save({"losses": losses, "sd": model.state_dict()}, os.path.join(out_dir, "dp.pt"))
    dist.barrier()
    dist.destroy_process_group()
