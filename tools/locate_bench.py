#!/usr/bin/env python
"""Nearest-waypoint kernels on 1M car states (BASELINE.json configs[3]) for both shipped centre lines: the default path (grid walk) against the
scanning kernel.  usage: tools/locate_bench.py [n]  -> one JSON object on stdout"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from triton_racer_sim_b200 import LocationTracker, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
tracks = np.load(os.path.join(ROOT, "tests", "golden", "tracks.npz"))


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


out = {"states": n}
for name in ("generated_track", "mountain_track"):
    wp = tracks[f"wp/{name}"]
    xyz, _, _, _ = synth.car_states(wp, n, seed=4)
    d = torch.from_numpy(xyz).cuda()
    res = {}
    ref = None
    for which in ("grid", "thread"):
        os.environ["TRS_LOCATE"] = which
        trk = LocationTracker(wp, device=0)
        del os.environ["TRS_LOCATE"]
        idx, seg = trk.locate_device(d)
        if ref is None:
            ref = (idx.clone(), seg.clone())
        else:
            res["identical"] = bool(torch.equal(idx, ref[0]) and torch.equal(seg, ref[1]))
        t = timed(lambda: trk.locate_device(d))
        res[which] = {"ms": t * 1e3, "states_per_s": n / t}
        trk.onShutdown()
    res["speedup"] = res["thread"]["ms"] / res["grid"]["ms"]
    out[name] = res
print(json.dumps(out))
