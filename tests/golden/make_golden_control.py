#!/usr/bin/env python
"""Generate tests/golden/control.npz by running the reference's own ControlMultiplexer, DriverAssistance and three_segment_map
(imported unmodified from /root/reference; `pygame`, which components/controller.py imports at module level for its joystick
classes, is replaced by an empty stub module: only the DriveMode enum of that file is used).

The multiplexer's launch locks end from sleeping threads (controlmultiplexer.py:55-58, 67-70), so the sequences below are driven in
real time with durations of 0.05 s / 0.15 s and steps placed at least 40 ms away from every lock boundary; the step times are stored
with the outputs.  Run in the build container: ``python tests/golden/make_golden_control.py``.
"""
import json
import os
import sys
import time
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.modules.setdefault("pygame", types.ModuleType("pygame"))

from TritonRacerSim.components.controller import DriveMode  # noqa: E402
from TritonRacerSim.components.controlmultiplexer import ControlMultiplexer  # noqa: E402
from TritonRacerSim.components.driver_assistance import DriverAssistance  # noqa: E402
from TritonRacerSim.core.config import config as REF_DEFAULTS  # noqa: E402
from TritonRacerSim.utils.mapping import three_segment_map  # noqa: E402

MODES = [DriveMode.HUMAN, DriveMode.AI_STEERING, DriveMode.AI]

MUX_CASES = {
    "no_locks": dict(),
    "both_locks": dict(ai_launch_boost_throttle_enabled=True, ai_launch_boost_throttle_value=0.8, ai_launch_boost_throttle_duration=0.15,
                       ai_launch_lock_steering_enabled=True, ai_launch_lock_steering_value=-0.25, ai_launch_lock_steering_duration=0.05),
    "throttle_lock": dict(ai_launch_boost_throttle_enabled=True, ai_launch_boost_throttle_value=1.0, ai_launch_boost_throttle_duration=0.05),
}
# per car: a script of (delay before the step in seconds, mode index)
SCRIPTS = [
    [(0.0, 0), (0.0, 2), (0.0, 2), (0.10, 2), (0.10, 2), (0.0, 0), (0.0, 2), (0.0, 1), (0.10, 1), (0.10, 2)],
    [(0.0, 1), (0.0, 1), (0.0, 2), (0.10, 2), (0.0, 2), (0.10, 2), (0.0, 1), (0.0, 2), (0.10, 0), (0.0, 2)],
    [(0.0, 2), (0.0, 2), (0.10, 0), (0.0, 2), (0.10, 2), (0.10, 2), (0.0, 2), (0.0, 2), (0.0, 0), (0.0, 0)],
]


def run_mux(rng):
    arrays, meta = {}, {}
    for cname, over in MUX_CASES.items():
        cfg = dict(REF_DEFAULTS)
        cfg.update(over)
        meta[cname] = over
        for si, script in enumerate(SCRIPTS):
            mux = ControlMultiplexer(cfg)
            t0 = time.perf_counter()
            rows = []
            for delay, mi in script:
                if delay:
                    time.sleep(delay)
                vals = rng.uniform(-1, 1, 6)
                now = time.perf_counter() - t0
                out = mux.step(MODES[mi], *[float(v) for v in vals])
                rows.append([now, mi, *vals, *out])
            time.sleep(0.2)                                  # let the lock threads end before the next multiplexer prints over them
            arrays[f"mux/{cname}/{si}"] = np.asarray(rows, np.float64)
    return arrays, meta


def run_assist(rng):
    arrays = {}
    n = 4000
    st = rng.uniform(-1.2, 1.2, n)
    th = rng.uniform(-1, 1, n)
    br = rng.uniform(0, 1, n)
    sp = rng.uniform(-2, 25, n)
    st[::97] = 0.0
    sp[::89] = 0.0
    arrays["assist/in"] = np.stack([st, th, br, sp])
    for mode in ("steering", "speed"):
        for k in (5, 2.5):
            cfg = dict(REF_DEFAULTS)
            cfg.update(drive_assist_limit_mode=mode, drive_assist_limit_k=k)
            da = DriverAssistance(cfg)
            out = np.asarray([da.step(float(a), float(b), float(c), float(d)) for a, b, c, d in zip(st, th, br, sp)], np.float64).T
            arrays[f"assist/{mode}/{k}"] = out
    return arrays


def run_pwm(rng):
    v = np.concatenate([rng.uniform(-1.5, 1.5, 2000), [0.0, -0.0, 1.0, -1.0, 1e-300, -1e-300]])
    arrays = {"pwm/in": v}
    for name, (a, b, c) in {"steering": (430, 350, 300), "throttle": (400, 370, 330), "odd": (-3.5, 0.25, 11.0)}.items():
        arrays[f"pwm/{name}"] = np.asarray([three_segment_map(float(x), a, b, c) for x in v], np.float64)
        arrays[f"pwm/{name}/map"] = np.asarray([a, b, c], np.float64)
    return arrays


def main():
    rng = np.random.default_rng(20261018)
    arrays, meta = run_mux(rng)
    arrays.update(run_assist(rng))
    arrays.update(run_pwm(rng))
    arrays["mux_cases_json"] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    np.savez_compressed(os.path.join(HERE, "control.npz"), **arrays)
    print("wrote control.npz:", len(arrays), "arrays")


if __name__ == "__main__":
    main()
