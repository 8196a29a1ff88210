"""GPU: tests added at the very end of round 2 whose code paths were last changed after the final GPU run of the round
(the GPU budget was exhausted).  They live in a file that sorts last so that `pytest -x` reaches every
other test first.

  * peer-memory all-reduce (csrc/peer_comm.cu, sharding.PeerReducer; needs two GPUs: NCCL + CUDA IPC) against NCCL, inside
    a CUDA graph, and a data-parallel fit step against the single-rank batch -- validated on a 2-GPU box BEFORE the
    reducer's set-up code was made failure-safe (same protocol, same kernel);
  * TimerGPU of misc/voxel_processing.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


# ------------------------------------------------------------------ peer-memory all-reduce + data-parallel training step
def _peer_worker(rank, world_size, port, out_dir):
    ngpu = torch.cuda.device_count()
    torch.cuda.set_device(rank)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), TOMOENC_PEER_ALLREDUCE="1")    # the peer path is opt-in
    dist.init_process_group("nccl", rank=rank, world_size=world_size, device_id=torch.device("cuda", rank))
    try:
        from tomoencoders_b200 import sharding
        dev = torch.device("cuda", rank)
        red = sharding.peer_reducer()
        assert red is not None, "peer-memory reducer not available on a box with %d GPUs" % ngpu
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        # 200 reductions of changing lengths: the 64-slot ring wraps three times
        for it in range(200):
            n = 1 + (37 * it) % 2048
            x = torch.randn(n, generator=g, device=dev, dtype=torch.float64)
            want = x.clone()
            dist.all_reduce(want)
            got = sharding.allreduce_sum_(x.clone())
            assert torch.allclose(got, want, rtol=1e-13, atol=1e-13), (it, n)
            # rank order summation: every rank holds the same bits
            both = [torch.empty_like(got) for _ in range(world_size)]
            dist.all_gather(both, got)
            assert torch.equal(both[0], both[1])
        # inside a CUDA graph, replayed: the epoch advances on the device
        buf = torch.zeros(512, dtype=torch.float64, device=dev)
        src = torch.full((512,), float(rank + 1), dtype=torch.float64, device=dev)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=s):
                buf.copy_(src)
                sharding.allreduce_sum_(buf)
                buf.mul_(2.0)
                sharding.allreduce_sum_(buf)
            for _ in range(70):
                gr.replay()
        torch.cuda.synchronize()
        tot = sum(range(1, world_size + 1))
        assert torch.equal(buf, torch.full_like(buf, 2.0 * tot * world_size))
        # ---- one data-parallel fit step: the two ranks' half batches reproduce the single-rank loss and gradient
        from tomoencoders_b200.neural_nets._train import UnetTrainer
        P = dict(n_filters=[16, 32], n_blocks=2, activation="lrelu", batch_norm=True, isconcat=[True] * 2, pool_size=[2, 2])
        rng = np.random.default_rng(9)
        x = rng.uniform(0, 1, (4, 32, 32, 32)).astype(np.float32)
        y = (rng.uniform(0, 1, (4, 32, 32, 32)) < 0.5).astype(np.uint8)
        tr = UnetTrainer(P, (32,) * 3, 2, mode="fp32", seed=3)
        tr.forward_backward(x[2 * rank:2 * rank + 2], y[2 * rank:2 * rank + 2])
        tr._post_backward()
        torch.cuda.synchronize()
        if rank == 0:
            np.save(os.path.join(out_dir, "dp_loss.npy"), sharding.allreduce_sum_(tr.loss.clone()).cpu().numpy())
            np.save(os.path.join(out_dir, "dp_grad.npy"), tr.grad.cpu().numpy())
        else:
            sharding.allreduce_sum_(tr.loss.clone())
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL + CUDA IPC peer mapping)")
def test_peer_memory_allreduce_matches_nccl_and_data_parallel_step_matches_single_rank(tmp_path):
    mp.spawn(_peer_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    from tomoencoders_b200.neural_nets._train import UnetTrainer
    P = dict(n_filters=[16, 32], n_blocks=2, activation="lrelu", batch_norm=True, isconcat=[True] * 2, pool_size=[2, 2])
    rng = np.random.default_rng(9)
    x = rng.uniform(0, 1, (4, 32, 32, 32)).astype(np.float32)
    y = (rng.uniform(0, 1, (4, 32, 32, 32)) < 0.5).astype(np.uint8)
    tr = UnetTrainer(P, (32,) * 3, 4, mode="fp32", seed=3)
    tr.forward_backward(x, y)
    torch.cuda.synchronize()
    loss1, grad1 = tr.loss.cpu().numpy(), tr.grad.cpu().numpy()
    loss2, grad2 = np.load(tmp_path / "dp_loss.npy"), np.load(tmp_path / "dp_grad.npy")
    assert abs(float(loss2[0]) - float(loss1[0])) < 1e-6 * abs(float(loss1[0]))
    assert np.abs(grad2 - grad1).max() < 1e-4 * np.abs(grad1).max()


def test_timer_gpu_measures_stream_time():
    """misc/voxel_processing.py:29-51 of the reference (CuPy events there, CUDA events of the current torch stream here)."""
    import torch
    from tomoencoders_b200.misc.voxel_processing import TimerGPU
    t = TimerGPU("ms")
    t.tic()
    x = torch.zeros(64 << 20, dtype=torch.uint8, device="cuda")
    x.add_(1)
    ms = t.toc()
    assert 0.0 < ms < 1000.0
    t2 = TimerGPU("secs")
    t2.tic()
    assert 0.0 <= t2.toc() < 1.0
