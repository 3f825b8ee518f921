#!/usr/bin/env python
"""Generate tests/golden/telemetry.npz by running the reference's own GymInterface.on_msg_recv and GymInterface.step
(TritonRacerSim/components/gyminterface.py:66-104, imported UNMODIFIED from /root/reference) on simulator-style telemetry packets.

`gym_donkeycar` (the socket client the class derives from) is not installed: a stand-in `SDClient` without a socket is injected, the object is
made without its connecting constructor, and `send_controls` is a no-op; everything between a received packet and the values the component
publishes (`cam/img`, `gym/x`, `gym/y`, `gym/z`, `gym/speed`, `gym/cte`) is the reference's code.
Run in the build container: ``python tests/golden/make_golden_telemetry.py``.
"""
import base64
import io
import json
import os
import sys
import types

import numpy as np
from PIL import Image

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

for name in ("gym_donkeycar", "gym_donkeycar.core", "gym_donkeycar.core.sim_client"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["gym_donkeycar.core.sim_client"].SDClient = type("SDClient", (), {})

from TritonRacerSim.components.gyminterface import DEFAULT_GYM_CONFIG, GymInterface  # noqa: E402

from triton_racer_sim_b200 import synth  # noqa: E402


def main():
    n, h, w = 6, 24, 32
    frames = synth.frame_pool(n, h, w, seed=83)
    rng = np.random.default_rng(83)
    gi = GymInterface.__new__(GymInterface)
    gi.gym_config = dict(DEFAULT_GYM_CONFIG)
    gi.send_controls = lambda *a: None
    packets, images, floats = [], [], []
    for k in range(n):
        buf = io.BytesIO()
        Image.fromarray(frames[k]).save(buf, format="JPEG")                      # the simulator sends img_enc 'JPG' (gyminterface.py:33)
        txt = base64.b64encode(buf.getvalue()).decode()
        # the simulator's JSON carries the numbers as numbers or as strings; the reference takes float() of either (gyminterface.py:100-104)
        pkt = {"msg_type": "telemetry", "image": txt, "pos_x": float(rng.normal(50, 20)), "pos_y": str(round(float(rng.normal(0.56, 0.01)), 6)),
               "pos_z": float(rng.normal(30, 20)), "speed": str(float(rng.uniform(0, 20))), "cte": float(rng.normal())}
        gi.on_msg_recv(pkt)
        out = gi.step(0.0, 0.0, None, False)                                     # -> cam/img, gym/x, gym/y, gym/z, gym/speed, gym/cte
        packets.append(pkt)
        images.append(out[0])
        floats.append([float(v) for v in out[1:]])
    images = np.stack(images)
    assert images.dtype == np.uint8 and images.shape == (n, h, w, 3)
    np.savez_compressed(os.path.join(HERE, "telemetry.npz"), packets_json=np.frombuffer(json.dumps(packets).encode(), np.uint8),
                        images=images, floats=np.asarray(floats, np.float64))
    print("telemetry.npz:", images.shape, np.asarray(floats).shape)


if __name__ == "__main__":
    main()
