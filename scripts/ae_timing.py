#!/usr/bin/env python3
"""Region breakdown of the autoencoder kernels (cycles of one producer / consumer warp per CTA, summed over CTAs) at B=200."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import signaltrain_b200 as st
from signaltrain_b200 import data
from signaltrain_b200.train import FusedTrainer
torch.manual_seed(218)
model = st.nn_proc.st_model(1, 4, 4).cuda()
lr, _ = st.learningrate.get_1cycle_schedule(1e-4, 200000, 1000, 200)
tr = FusedTrainer(model, lr)
x, y, k = (torch.from_numpy(a).cuda() for a in data.make_pool(200, 8192, 2048, data.Compressor_4c()))
for _ in range(3):
    tr.step(x, y, k)
eng = tr.eng
buf = (ctypes.c_longlong * 24)()
eng.lib.st_debug_ae_timing(eng.h, 1, None)
n = 5
for _ in range(n):
    tr.step(x, y, k)
eng.lib.st_debug_ae_timing(eng.h, 1, buf)
names = ["producer: slot acquire (empty wait)", "producer: cp.async wait", "producer: output side", "producer: dgrad layers 9..2", "producer: layer 1 + input side", "-", "consumer u1a: waiting for full", "consumer u1a: wgrad work"]
if os.environ.get("ST_DISABLE_FFMA2_AE_BWD") == "1":
    names = ["sync0(prev store/step0)", "stage issue", "bwd-data mma", "cp.async wait+sync", "wgrad", "bias_grad", "sync+store_gz", "L1 epilogue+sync"]
ncta = 148
for ae in range(2):
    tot = sum(buf[8 * ae + i] for i in range(8))
    print("AE", ae, "total cycles/CTA/step", tot / ncta / n)
    for i in range(8):
        print("   %-38s %9.0f cyc/CTA/step  %5.1f%%" % (names[i], buf[8 * ae + i] / ncta / n, 100.0 * buf[8 * ae + i] / max(tot, 1)))
eng.lib.st_debug_ae_timing(eng.h, 0, None)

fnames = ["input stage", "layers (FFMA2 + reload)", "record copies", "output stage"]
print("FFMA2 forward, warp 0, cycles/CTA/step:")
for ae in range(2):
    tot = sum(buf[16 + 4 * ae + i] for i in range(4))
    print("AE", ae, "total", tot / ncta / n)
    for i in range(4):
        print("   %-22s %9.0f  %5.1f%%" % (fnames[i], buf[16 + 4 * ae + i] / ncta / n, 100.0 * buf[16 + 4 * ae + i] / max(tot, 1)))
