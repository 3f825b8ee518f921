#!/usr/bin/env python
"""Frames/s of every reference-generated configuration (tests/golden/images.npz cases_json) at one frame size, device-resident inputs.
usage: tools/bench_configs.py [h w n [config,config,...]]   (default 240 320 8192, every configuration) -> one JSON object on stdout"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from tests.helpers import cfg_for
from triton_racer_sim_b200 import ImgPreprocessing, synth

h, w, n = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (240, 320, 8192)
cases = json.loads(bytes(np.load(os.path.join(ROOT, "tests", "golden", "images.npz"))["cases_json"]).decode())
pool = torch.from_numpy(synth.frame_pool(128, h, w)).cuda()
frames = synth.expand_torch(pool, n)
u8, f32 = torch.empty_like(frames), torch.empty(frames.shape, dtype=torch.float32, device="cuda")
out = {"h": h, "w": w, "frames": n, "configs": {}}
only = sys.argv[4].split(",") if len(sys.argv) >= 5 else None
for name, over in cases.items():
    if only and name not in only:
        continue
    print("config", name, file=sys.stderr, flush=True)
    comp = ImgPreprocessing(cfg_for(over), device=0)
    for want_f32 in (False, True):
        fn = lambda: comp.process_device(frames, out_u8=u8, out_f32=f32 if want_f32 else None, want_f32=want_f32)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out["configs"].setdefault(name, {})["u8+f32" if want_f32 else "u8"] = n / (e0.elapsed_time(e1) / 10 * 1e-3)
    comp.onShutdown()
base = out["configs"].get("full_house")
for name, v in out["configs"].items():
    if base:
        v["vs_full_house"] = {k: v[k] / base[k] for k in ("u8", "u8+f32")}
print(json.dumps(out))
