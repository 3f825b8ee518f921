"""Throughput of the pilots' forward pass (tensor cores): N frames of 120x160 u8 on the device -> (N,2) model outputs.
usage: python tools/pilot_bench.py [frames] [max_batch] [model: cnn_2d | cnn_2d_full_house ...]"""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
from oracle import pilot_ref as ref
from triton_racer_sim_b200 import synth
from triton_racer_sim_b200.pilot import ModelType, PilotNet, _KIND
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
cap = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
mt = ModelType(sys.argv[3]) if len(sys.argv) > 3 else ModelType.CNN_2D_FULL_HOUSE
h, w = 120, 160
pool = torch.from_numpy(synth.frame_pool(256, h, w)).cuda()
frames = synth.expand_torch(pool, n)
wts = ref.random_weights(_KIND[mt], h, w, seed=1)
net = PilotNet(mt, wts, h, w, device=0, max_batch=cap)
spd = torch.rand(n, device='cuda'); loc = torch.rand(n, device='cuda') * 10
out = torch.empty((n, 2), dtype=torch.float32, device='cuda')
for _ in range(3):
    net.forward_device(frames, spd, loc, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    net.forward_device(frames, spd, loc, out=out)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / reps * 1e-3
macs = 0
hh, ww, c = h, w, 3
for k, s, f in ref.CONVS:
    hh, ww = (hh - k) // s + 1, (ww - k) // s + 1
    macs += hh * ww * f * k * k * c
    c = f
macs += hh * ww * c * 100 * (2 if mt == ModelType.CNN_2D_FULL_HOUSE else 1)
print(f"{mt.value}: {n} frames, chunk {cap}: {n / t / 1e6:.3f} M frames/s ({t * 1e3:.2f} ms), {2 * macs * n / t / 1e12:.1f} TFLOP/s algorithmic "
      f"({2 * macs / 1e6:.1f} MFLOP per frame)")
