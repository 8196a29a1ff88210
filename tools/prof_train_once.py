"""One eager training step of M_a01 (8 x 64^3) after warm-up, for `ncu --metrics gpu__time_duration.sum` launch lists:
cudaProfilerStart/Stop bracket exactly one step."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
os.environ["TOMOENC_TRAIN_GRAPH"] = "0"
from tomoencoders_b200.neural_nets import _train  # noqa: E402

M_A01 = dict(n_filters=[16, 32, 64], n_blocks=3, pool_size=[2, 2, 2], isconcat=[True] * 3)
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
tr = _train.UnetTrainer(M_A01, (64,) * 3, batch, mode="bf16", seed=1)
g = np.random.default_rng(0)
x = torch.from_numpy(g.uniform(0, 1, (batch, 64, 64, 64)).astype(np.float32)).cuda()
y = (x > 0.5).to(torch.uint8)
for _ in range(2):
    tr.train_step(x, y)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
tr.train_step(x, y)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
