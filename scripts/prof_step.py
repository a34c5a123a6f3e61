#!/usr/bin/env python3
"""A few fused train steps at BASELINE configs[1] size (B=200, C=8192, K=4) -- the command profiled under ncu."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import signaltrain_b200 as st
from signaltrain_b200 import data
from signaltrain_b200.train import FusedTrainer

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 200
torch.manual_seed(218)
model = st.nn_proc.st_model(1, 4, 4).cuda()
lr, _ = st.learningrate.get_1cycle_schedule(1e-4, 200000, 1000, 200)
tr = FusedTrainer(model, lr)
x, y, k = (torch.from_numpy(a).cuda() for a in data.make_pool(B, model.in_chunk_size, model.out_chunk_size, data.Compressor_4c()))
for _ in range(steps):
    loss = tr.step(x, y, k)
torch.cuda.synchronize()
print("loss", loss.item())
