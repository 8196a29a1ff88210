"""Data-parallel training step alone (BASELINE configs[4]: 64 pairs of 64^3 split over the ranks), for A/B runs of the
communication paths under torchrun:
    TOMOENC_PEER_ALLREDUCE=0|1  peer-memory one-shot all-reduce of the SyncBN sums (csrc/peer_comm.cu) vs NCCL
    TOMOENC_GRAD_BUCKETS=0|1    gradient all-reduce in buckets on a side stream during the backward pass vs one flat call
Prints one line per run from rank 0: ms per step (max over ranks, CUDA events), loss.
ALWAYS run it under `timeout`: the combination peer = 1, buckets = 1 hung at 8 ranks."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tomoencoders_b200.neural_nets._train import UnetTrainer  # noqa: E402

M_A01 = dict(n_filters=[16, 32, 64], n_blocks=3, activation="lrelu", batch_norm=True, isconcat=[True] * 3, pool_size=[2, 2, 2])


def main():
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    tb = max(8, 64 // world)
    tr = UnetTrainer(M_A01, (64,) * 3, tb, mode="bf16", seed=13)
    g = torch.Generator(device=dev).manual_seed(8 + rank)
    x = torch.rand((tb, 64, 64, 64), generator=g, device=dev)
    y = (torch.rand((tb, 64, 64, 64), generator=g, device=dev) < 0.5).to(torch.uint8)
    for _ in range(5):
        tr.train_step(x, y)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    n = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        loss = tr.train_step(x, y)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("ranks %d batch/rank %d peer=%s buckets=%s: %.3f ms/step, loss %.6f, graph launches %d" % (
            world, tb, os.environ.get("TOMOENC_PEER_ALLREDUCE", "0"), os.environ.get("TOMOENC_GRAD_BUCKETS", "0"), float(ms.item()),
            loss, tr.graph_launches), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
