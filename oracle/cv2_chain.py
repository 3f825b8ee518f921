"""The reference's per-frame call sequence over OpenCV/numpy/math.  TEST INFRASTRUCTURE ONLY.

The reference's pixel arithmetic is delegated to third-party code (OpenCV; ``pip install opencv-python``
unpinned at /root/reference/README.md:32; this image pins opencv-python-headless 4.13.0.92, numpy 2.3.5).
This module re-issues the same library calls in the same order as
TritonRacerSim/components/img_preprocessing.py:37-102, so it runs with the reference's CPU cost and the
reference's results on any box that has cv2, without needing ``/root/reference`` (absent on the GPU box).
``tests/test_oracle_golden.py`` pins it, like the C restatement, against golden vectors made by the real
reference.  ``bench.py`` times it as the reference CPU path.
"""
from __future__ import annotations

import math

import numpy as np

try:  # cv2 is part of the image; keep the import soft so the C oracle stays usable without it
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

DEFAULT_HSVS = [((0, 0, 130), (180, 64, 255)), ((25, 180, 155), (43, 255, 255))]


def adjust(img: np.ndarray, cfg: dict) -> np.ndarray:
    """img_preprocessing.py:81-102 (__trim_brightness_contrast)."""
    ratio = cfg.get("preprocessing_contrast_enhancement_ratio", 1.0)
    offset = cfg.get("preprocessing_contrast_enhancement_offset", 125)
    baseline = cfg.get("preprocessing_brightness_baseline", 550)
    now = sum(list(cv2.mean(img[40:119, :, :])))          # :88
    shift = (baseline - now) / 3                           # :89
    work = img.astype(np.float32)                          # :92
    if cfg.get("preprocessing_dynamic_brightness_enabled", False):
        work += shift                                      # :94
    work -= offset                                         # :95
    work *= ratio                                          # :96
    work += offset                                         # :97
    return np.clip(work, 0, 255).astype(np.uint8)          # :98-99


def colour_layers(img: np.ndarray, cfg: dict):
    """img_preprocessing.py:65-74 (__color_filter)."""
    hsv = cv2.cvtColor(img.copy(), cv2.COLOR_RGB2HSV)
    return [cv2.inRange(hsv, tuple(lo), tuple(hi)) for lo, hi in cfg.get("preprocessing_color_filter_hsvs", DEFAULT_HSVS)]


def edge_layer(img: np.ndarray, cfg: dict) -> np.ndarray:
    """img_preprocessing.py:76-79 (__edge_detection)."""
    return cv2.Canny(img, cfg.get("preprocessing_edge_detection_threshold_a", 60),
                     cfg.get("preprocessing_edge_detection_threshold_b", 100))


def process(img: np.ndarray, cfg: dict) -> np.ndarray:
    """img_preprocessing.py:37-62 (__process, __merge)."""
    img = adjust(img, cfg)
    layers, dests = [], []
    if cfg.get("preprocessing_color_filter_enabled", False):
        layers.extend(colour_layers(img, cfg))
        dests.extend(cfg.get("preprocessing_color_filter_destination_channels", [0, 1]))
    if cfg.get("preprocessing_edge_detection_enabled", False):
        layers.append(edge_layer(img, cfg))
        dests.append(cfg.get("preprocessing_edge_detection_destination_channel", 2))
    assert len(layers) == len(dests)
    for layer, ch in zip(layers, dests):
        img[:, :, ch] = layer
    return img


def normalise(img: np.ndarray) -> np.ndarray:
    """keras_pilot.py:49-50."""
    arr = np.array(img, dtype=np.float32)   # asarray of a u8 frame always makes a fresh f32 array
    arr /= 255
    return arr


def full_chain(img: np.ndarray, cfg: dict):
    """process + the pilot's float normalisation, as one car-loop tick sees it."""
    out = process(img, cfg)
    f = np.asarray(out, dtype=np.float32)
    f /= 255
    return out, f


def resize_nearest(img: np.ndarray, w_out: int, h_out: int) -> np.ndarray:
    """camera.py:36 — pygame.transform.scale is nearest-neighbour; cv2.INTER_NEAREST has the same index rule."""
    return cv2.resize(img, (w_out, h_out), interpolation=cv2.INTER_NEAREST)


def locate(waypoints, point, min_map=0, max_map=10):
    """track_data_process.py:89-107 — pure-Python loop, as the reference runs it."""
    best_i, best_d = 0, 100
    for i, wp in enumerate(waypoints):
        d = abs(point[0] - wp[0]) + abs(point[1] - wp[1]) + abs(point[2] - wp[2])
        if d < best_d:
            best_i, best_d = i, d
    return best_i, best_i / float(len(waypoints)) * (max_map - min_map) + min_map


def throttle_law(cur, target, mult):
    """utils/mapping.py:23-28."""
    t = mult * math.atan((target - cur) * 2) / (math.pi / 2)
    return 0.0 if -0.2 < t < 0.0 else t


def brake_law(cur, target, mult):
    """utils/mapping.py:30-35."""
    b = -1.0 * mult * math.atan((target - cur) * 1.0) / (math.pi / 2)
    return 0.0 if b < 0.4 else b


def pilot_tail(real_spd: float, model_steer: np.float32, model_spd: np.float32, cfg: dict):
    """keras_pilot.py:80-95 (== 99-118) with __cap (142-145) and __smooth_steering (147-153)."""
    steering = model_steer
    if steering < -1.0:
        steering = -1.0
    elif steering > 1.0:
        steering = 1.0
    predicted = model_spd * 20
    thr_k = cfg.get("spd_ctl_threshold", 1.1)
    throttle = throttle_law(real_spd, predicted * thr_k, cfg.get("spd_ctl_reverse_multiplier", 1.0))
    breaking = 0.0
    if cfg.get("spd_ctl_break", False):
        throttle = 1.0 if predicted - real_spd > 0.0 else 0.0
        breaking = brake_law(real_spd, predicted * thr_k, cfg.get("spd_ctl_break_multiplier", 1.0))
    if cfg.get("smooth_steering_enabled", False):
        st = cfg.get("smooth_steering_threshold", 0.9)
        if steering > st:
            steering = 1.0
        elif steering < st * -1:
            steering = -1.0
    return float(steering), float(throttle), float(breaking)
