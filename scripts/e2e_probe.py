import sys, time, os
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import signaltrain_b200 as st
from signaltrain_b200 import data
from signaltrain_b200.train import FusedTrainer
torch.manual_seed(218)
model = st.nn_proc.st_model(1, 4, 4).cuda()
lr, _ = st.learningrate.get_1cycle_schedule(1e-4, 200000, 1000, 200)
tr = FusedTrainer(model, lr)
B = 200
pool = data.make_pool(B * 24, 8192, 2048, data.Compressor_4c(), 44100, seed=1)
pinned = [torch.from_numpy(a).pin_memory() for a in pool]
host = [tuple(p[i * B:(i + 1) * B] for p in pinned) for i in range(24)]
print("pinned views:", [t.is_pinned() for t in host[3]])
dv = [tuple(t.cuda() for t in b) for b in host[:4]]
for b in dv: tr.step(*b)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter(); tr.run_host_batches(host[:20]); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("run_host_batches 20 steps: %.3f ms/step" % ((t1 - t0) * 1e3 / 20))
t0 = time.perf_counter()
for i in range(20): tr.step(*dv[i % 4])
torch.cuda.synchronize(); print("device-resident: %.3f ms/step" % ((time.perf_counter() - t0) * 1e3 / 20))
# raw H2D bandwidth from the pinned views
xs = torch.empty_like(dv[0][0]); torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(20): xs.copy_(host[i][0], non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("H2D x only: %.3f ms per copy, %.1f GB/s" % (dt * 1e3 / 20, 20 * xs.numel() * 4 / dt / 1e9))
