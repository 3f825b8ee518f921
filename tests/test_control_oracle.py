"""The control post-processing oracle (oracle/control.py) against vectors produced by the reference's own ControlMultiplexer,
DriverAssistance and three_segment_map (tests/golden/control.npz, tests/golden/make_golden_control.py)."""
import json
import os

import numpy as np
import pytest

from oracle import control as oc
from triton_racer_sim_b200.config import default_config

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def golden_control():
    return np.load(os.path.join(HERE, "golden", "control.npz"))


def mux_cases(g):
    return json.loads(bytes(g["mux_cases_json"]).decode())


def replay_mux(rows, cfg, step_fn):
    """Feed one car's recorded sequence through step_fn(mode (1,), usr (3,1), ai (3,1), now, last_mode, launch_time) -> (3,1)."""
    last_mode = np.zeros(1, np.int32)
    launch = np.full((oc.LAUNCH_SLOTS, 1), oc.NEVER)
    outs = []
    for r in rows:
        now, mi = r[0], int(r[1])
        usr, ai = r[2:5].reshape(3, 1).copy(), r[5:8].reshape(3, 1).copy()
        outs.append(step_fn(np.array([mi], np.int32), usr, ai, now, last_mode, launch)[:, 0])
    return np.asarray(outs)


def test_mux_oracle_matches_reference_sequences(golden_control):
    n = 0
    for cname, over in mux_cases(golden_control).items():
        cfg = default_config()
        cfg.update(over)
        for si in range(3):
            rows = golden_control[f"mux/{cname}/{si}"]
            got = replay_mux(rows, cfg, lambda *a: oc.control_mux(*a, cfg))
            assert np.array_equal(got, rows[:, 8:11]), f"{cname}/{si}"
            n += len(rows)
    assert n == 90


def test_assist_and_pwm_oracle_match_reference(golden_control):
    st, th, br, sp = golden_control["assist/in"]
    for mode in ("steering", "speed"):
        for k in (5, 2.5):
            cfg = default_config()
            cfg.update(drive_assist_limit_mode=mode, drive_assist_limit_k=k)
            got = np.stack(oc.driver_assist(st, th, br, sp, cfg))
            assert np.array_equal(got, golden_control[f"assist/{mode}/{k}"]), f"{mode}/{k}"
    v = golden_control["pwm/in"]
    for name in ("steering", "throttle", "odd"):
        a, b, c = golden_control[f"pwm/{name}/map"]
        assert np.array_equal(oc.three_segment_map(v, a, b, c), golden_control[f"pwm/{name}"]), name
