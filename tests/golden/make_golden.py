#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE ITSELF (imported unmodified from /root/reference).

Run in the build container (``python tests/golden/make_golden.py``); the GPU box has no /root/reference, so
the vectors produced here are committed.  Inputs are stored next to the outputs so nothing depends on a
random generator reproducing itself.

What is executed (nothing is restated except where the module cannot be imported):
  * ImgPreprocessing(cfg)._ImgPreprocessing__process(img)    components/img_preprocessing.py:37-54
  * LocationTracker(path).step(x, y, z)                       components/track_data_process.py:81-84
  * calcThrottle / calcBreak                                   utils/mapping.py:23-35
  * keras_pilot.py cannot be imported (tensorflow absent); its 15-line speed-control call sequence
    (keras_pilot.py:80-95, == 99-118) and __cap/__smooth_steering (142-153) are replayed here around the
    imported calcThrottle/calcBreak with the same numpy scalar types the model returns (np.float32).
"""
import json
import os
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from TritonRacerSim.components.img_preprocessing import ImgPreprocessing  # noqa: E402
from TritonRacerSim.components.track_data_process import LocationTracker  # noqa: E402
from TritonRacerSim.core.config import config as REF_DEFAULTS  # noqa: E402
from TritonRacerSim.utils.mapping import calcBreak, calcThrottle  # noqa: E402

from triton_racer_sim_b200 import synth  # noqa: E402

TRACK_DIR = os.path.join(REF, "TritonRacerSim/car_templates/track_data")


def ref_cfg(**kw):
    cfg = dict(REF_DEFAULTS)
    cfg["preprocessing_preview_enabled"] = False       # headless OpenCV has no imshow
    cfg.update(kw)
    return json.loads(json.dumps(cfg))                 # the JSON round trip of generate/read_config: tuples -> lists


IMAGE_CASES = {
    # name: overrides on the reference defaults
    "full_house": dict(preprocessing_color_filter_enabled=True, preprocessing_edge_detection_enabled=True),
    "colour_only": dict(preprocessing_color_filter_enabled=True),
    "edge_only": dict(preprocessing_edge_detection_enabled=True),
    "adjust_only": dict(),
    "dyn_contrast": dict(preprocessing_color_filter_enabled=True, preprocessing_edge_detection_enabled=True,
                         preprocessing_dynamic_brightness_enabled=True, preprocessing_contrast_enhancement_ratio=1.3,
                         preprocessing_brightness_baseline=420),
    "exotic": dict(preprocessing_color_filter_enabled=True, preprocessing_edge_detection_enabled=True,
                   preprocessing_dynamic_brightness_enabled=True, preprocessing_contrast_enhancement_ratio=0.7,
                   preprocessing_contrast_enhancement_offset=100.5, preprocessing_brightness_baseline=400.25,
                   preprocessing_color_filter_hsvs=[((170, 40.5, 60), (180, 255, 250.5)), ((0, 0, 0), (12.5, 300, 200)),
                                                    ((90, 30, 30), (130, 255, 255))],
                   preprocessing_color_filter_destination_channels=[2, 0, 0],
                   preprocessing_edge_detection_threshold_a=150.7, preprocessing_edge_detection_threshold_b=50.2,
                   preprocessing_edge_detection_destination_channel=1),
    "edge_tight": dict(preprocessing_edge_detection_enabled=True, preprocessing_edge_detection_threshold_a=10,
                       preprocessing_edge_detection_threshold_b=400, preprocessing_edge_detection_destination_channel=0),
}


def special_frames():
    """Adversarial frames for the edge filter: long weak chains hanging off one strong pixel, plateaus (ties), borders."""
    h, w = 120, 160
    out = []
    rng = np.random.default_rng(99)
    # 1. a spiral of low contrast with one high-contrast blob at its centre
    img = np.full((h, w, 3), 100, np.uint8)
    y, x, dy, dx, step = 60, 80, 0, 1, 2
    pts = []
    for leg in range(60):
        for _ in range(step):
            if 2 <= y < h - 2 and 2 <= x < w - 2:
                pts.append((y, x))
            y, x = y + dy, x + dx
        dy, dx = dx, -dy
        if leg % 2 == 1:
            step += 4
    for (yy, xx) in pts:
        img[yy, xx] = 118
    img[58:63, 78:83] = 255
    out.append(img)
    # 2. vertical / horizontal / diagonal ramps and plateaus (equal magnitudes -> tie rules)
    img = np.zeros((h, w, 3), np.uint8)
    img[:, :, 0] = (np.arange(w)[None, :] // 8 * 12) % 256
    img[:, :, 1] = (np.arange(h)[:, None] // 6 * 20) % 256
    img[:, :, 2] = ((np.arange(w)[None, :] + np.arange(h)[:, None]) // 10 * 25) % 256
    out.append(img)
    # 3. checkerboard + border-touching bright frame
    img = ((np.indices((h, w)).sum(0) // 4 % 2) * 70 + 60).astype(np.uint8)[..., None].repeat(3, 2)
    img[0, :] = 255; img[-1, :] = 255; img[:, 0] = 0; img[:, -1] = 255
    out.append(img)
    # 4. long diagonal weak line with strong end + random weak speckle
    img = np.full((h, w, 3), 60, np.uint8)
    for i in range(110):
        img[5 + i, 10 + i] = (75, 60, 60)
        img[115 - i, 30 + i] = (60, 78, 60)
    img[5:8, 10:13] = (255, 255, 255)
    sp = rng.random((h, w)) < 0.02
    img[sp] = (60, 60, 90)
    out.append(img)
    # 5/6. constant frames
    out.append(np.zeros((h, w, 3), np.uint8))
    out.append(np.full((h, w, 3), 255, np.uint8))
    # 7. saturated colours for the HSV tie rules (v==r==g etc.)
    img = np.zeros((h, w, 3), np.uint8)
    vals = np.array([0, 1, 64, 127, 128, 129, 130, 155, 180, 254, 255], np.uint8)
    combos = np.array(np.meshgrid(vals, vals, vals)).reshape(3, -1).T
    flat = img.reshape(-1, 3)
    flat[: len(combos)] = combos
    flat[len(combos): 2 * len(combos)] = combos[::-1]
    out.append(img)
    return np.stack(out)


def frame_sets():
    sets = {
        "f120": np.concatenate([synth.frame_pool(8, 120, 160), special_frames()]),
        "f240": synth.frame_pool(4, 240, 320)[[1, 3]],
        "odd_7x9": synth.frame_pool(3, 7, 9, seed=5),
        "odd_33x50": synth.frame_pool(3, 33, 50, seed=6),        # fewer than 40 rows: empty brightness ROI
        "odd_45x160": synth.frame_pool(2, 45, 160, seed=7),      # ROI clipped to rows 40..44
        "odd_121x163": synth.frame_pool(2, 121, 163, seed=8),    # row bytes not a multiple of 4
        "odd_64x96": synth.frame_pool(2, 64, 96, seed=9),
    }
    # brightness-shifted variants so dynamic brightness sees different ROI means
    sets["f120_shift"] = synth.expand_numpy(synth.frame_pool(4, 120, 160), 6, start=4 * 3 + 1)
    return sets


def run_images():
    sets = frame_sets()
    arrays = {}
    for sname, frames in sets.items():
        arrays[f"in/{sname}"] = frames
        for cname, over in IMAGE_CASES.items():
            if (sname.startswith("odd") or sname == "f240") and cname in ("colour_only", "adjust_only", "edge_tight"):
                continue
            if sname == "f120_shift" and cname not in ("dyn_contrast", "exotic"):
                continue
            cfg = ref_cfg(**over)
            comp = ImgPreprocessing(cfg)
            outs = []
            for img in frames:
                before = img.copy()
                res = comp._ImgPreprocessing__process(img.copy())
                assert np.array_equal(img, before)
                outs.append(res)
            arrays[f"out/{sname}/{cname}"] = np.stack(outs)
    arrays["cases_json"] = np.frombuffer(json.dumps(IMAGE_CASES).encode(), np.uint8)
    np.savez_compressed(os.path.join(HERE, "images.npz"), **arrays)
    print("images.npz:", {k: v.shape for k, v in arrays.items() if k.startswith("in/")})


def run_tracks():
    arrays = {}
    for name in ("generated_track", "mountain_track"):
        path = os.path.join(TRACK_DIR, name + ".json")
        wp = np.asarray(json.load(open(path)), np.float64)
        arrays[f"wp/{name}"] = wp
        n = 3000 if name == "generated_track" else 1500
        xyz, _, _, _ = synth.car_states(wp, n, seed=11)
        # a few hand-made states: exactly on waypoints (duplicates -> first index), far away (sentinel -> 0), distance ~100
        xyz[0] = wp[len(wp) // 2]
        xyz[1] = wp[-1]
        xyz[2] = wp[0] + np.array([60.0, 20.0, 20.0])            # L1 distance exactly 100 from wp[0] -> not < 100
        xyz[3] = wp[0] + np.array([59.999, 20.0, 20.0])
        for lo, hi, tag in ((0, 10, "0_10"), (-2.5, 7.25, "m2p5_7p25")):
            tracker = LocationTracker(path, lo, hi)
            seg = np.array([tracker.step(float(p[0]), float(p[1]), float(p[2]))[0] for p in xyz], np.float64)
            arrays[f"seg/{name}/{tag}"] = seg
        tracker = LocationTracker(path)
        idx = np.array([tracker._LocationTracker__find_closest((float(p[0]), float(p[1]), float(p[2])))[0] for p in xyz], np.int32)
        arrays[f"xyz/{name}"] = xyz
        arrays[f"idx/{name}"] = idx
    np.savez_compressed(os.path.join(HERE, "tracks.npz"), **arrays)
    print("tracks.npz:", {k: v.shape for k, v in arrays.items()})


SPD_CASES = {
    "default": dict(),
    "break": dict(spd_ctl_break=True),
    "break_smooth": dict(spd_ctl_break=True, smooth_steering_enabled=True, spd_ctl_break_multiplier=1.7,
                         spd_ctl_reverse_multiplier=0.6, spd_ctl_threshold=0.95, smooth_steering_threshold=0.75),
}


def pilot_tail(cfg, real_spd, model_steer, model_spd, numpy1=False):
    """keras_pilot.py:80-95 replayed with the reference's own calcThrottle/calcBreak.  numpy1: what NumPy 1.x scalar promotion makes of
    line 83 (`np.float32 * 20` is float64 there, so everything after it is float64 too); NumPy >= 2 keeps float32 (NEP 50)."""
    steering = model_steer                                           # numpy()[0][0] -> np.float32
    if steering < -1.0: steering = -1.0                              # __cap, :142-145
    elif steering > 1.0: steering = 1.0
    if numpy1:
        model_spd = np.float64(model_spd)
    predicted_speed = model_spd * 20                                 # :83
    breaking = 0.0
    throttle = calcThrottle(real_spd, predicted_speed * cfg['spd_ctl_threshold'], cfg['spd_ctl_reverse_multiplier'])
    if cfg['spd_ctl_break']:
        throttle = 1.0 if predicted_speed - real_spd > 0.0 else 0.0
        breaking = calcBreak(real_spd, predicted_speed * cfg['spd_ctl_threshold'], cfg['spd_ctl_break_multiplier'])
    if cfg['smooth_steering_enabled']:                               # __smooth_steering, :147-153
        if steering > cfg['smooth_steering_threshold']: steering = 1.0
        elif steering < cfg['smooth_steering_threshold'] * -1: steering = -1.0
    return float(steering), float(throttle), float(breaking)


def run_speed():
    wp = synth.synthetic_track(400)
    _, cur, model_spd, steer = synth.car_states(wp, 20000, seed=21)
    arrays = {"cur": cur, "model_spd": model_spd, "model_steer": steer}
    for cname, over in SPD_CASES.items():
        cfg = ref_cfg(**over)
        res = np.array([pilot_tail(cfg, float(c), s, m) for c, s, m in zip(cur, steer, model_spd)], np.float64)
        arrays[f"out/{cname}"] = res
        arrays[f"out_numpy1/{cname}"] = np.array([pilot_tail(cfg, float(c), s, m, numpy1=True) for c, s, m in zip(cur, steer, model_spd)], np.float64)
    arrays["feature"] = np.array([np.asarray(float(c) / 20, dtype=np.float32) for c in cur], np.float32)   # keras_pilot.py:100
    arrays["cases_json"] = np.frombuffer(json.dumps(SPD_CASES).encode(), np.uint8)
    np.savez_compressed(os.path.join(HERE, "speed.npz"), **arrays)
    print("speed.npz:", {k: v.shape for k, v in arrays.items()})


if __name__ == "__main__":
    run_images()
    run_tracks()
    run_speed()
