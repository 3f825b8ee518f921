"""Shared helpers for the parity tests."""
import json

import numpy as np

from triton_racer_sim_b200.config import default_config


def image_cases(golden_images):
    return json.loads(bytes(golden_images["cases_json"]).decode())


def cfg_for(case_overrides):
    cfg = default_config()
    cfg.update(json.loads(json.dumps(case_overrides)))
    return cfg


def golden_pairs(golden_images):
    """Yield (set_name, case_name, cfg, frames, expected)."""
    cases = image_cases(golden_images)
    for key in golden_images.files:
        if not key.startswith("out/"):
            continue
        _, sname, cname = key.split("/")
        yield sname, cname, cfg_for(cases[cname]), golden_images[f"in/{sname}"], golden_images[key]


def speed_cases(golden_speed):
    return json.loads(bytes(golden_speed["cases_json"]).decode())


def rel_close(a, b, rtol):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) <= rtol * np.maximum(np.abs(b), 1e-30)
