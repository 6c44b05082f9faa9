"""Multi-GPU check of the CUDA-graph replay of the data-parallel step (run under torchrun on >= 2 B200s of one node):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tests/dp_graph_check.py

Every rank captures its step (SyncBN peer exchanges, symmetric-memory barrier + in-GEMM peer reads or gather-once, NCCL
logsumexp all-gather and two-bucket gradient all-reduce inside the graph) and replays it three times; a twin trainer then
makes the same three steps eagerly with the same host seeds at seed epochs 1..3.  Losses and parameters must agree."""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (HERE, os.path.dirname(HERE)):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    from multimodal_eeg_fmri_b200 import functional as XF, ops
    from multimodal_eeg_fmri_b200 import synthetic
    from multimodal_eeg_fmri_b200.training import PairedBridgeModel, PairedTrainer, init_distributed

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    init_distributed("nccl")
    ok = True
    for Bl in (128, 640):
        eeg, roi, _ = synthetic.paired_batch(Bl * world, 8, 64, 12, 20, seed=5)
        sl = slice(rank * Bl, (rank + 1) * Bl)
        eeg, roi = eeg[sl].cuda(), roi[sl].cuda()

        def make():
            torch.manual_seed(7)
            m = PairedBridgeModel(eeg_channels=8, n_roi=12, eeg_hidden=128, fmri_hidden=16, bridge_dim=32, dropout=0.1,
                                  fmri_dropout=0.2, encoder="v4")
            return m, PairedTrainer(m.cuda().train(), lr=1e-3)

        ma, ta = make()
        mb, tb = make()
        tb.use_device_state()
        ops.seed_epoch_set(0)
        XF.manual_seed(99)
        g = ta.capture(eeg, roi)
        base = XF.seed_state()[0]
        lg = [float(g.replay()) for _ in range(3)]
        torch.cuda.synchronize()
        le = []
        for k in (1, 2, 3):
            XF.set_seed_state((base, 0))
            ops.seed_epoch_set(k)
            le.append(float(tb.step(eeg, roi)))
        ops.seed_epoch_set(0)
        torch.cuda.synchronize()
        perr = max(float((a - b).abs().max()) for a, b in zip(ma.state_dict().values(), mb.state_dict().values()))
        lerr = max(abs(a - b) for a, b in zip(lg, le))
        good = lerr <= 1e-6 * abs(le[0]) and perr <= 1e-7 and lg[2] < lg[0]
        flag = torch.tensor([1.0 if good else 0.0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = ok and bool(flag.item())
        if rank == 0:
            print(f"dp_graph_check world={world} local_batch={Bl} calls_captured={g.launches_captured} losses={['%.6f' % v for v in lg]} "
                  f"loss_abs_diff={lerr:.2e} param_max_abs_diff={perr:.2e} {'OK' if flag.item() else 'FAIL'}", flush=True)
        del g
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
