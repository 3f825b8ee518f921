"""CPU oracle for the observation hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package; the product (``triton-racer-sim_b200/``) never does.

Two layers:

* ``trs_oracle.c`` (built to ``libtrs_oracle.so`` by ``oracle/build.py``): a plain-C restatement of every
  arithmetic stage, each function citing the reference file:line it follows.  This is *the* checker.
* ``oracle/cv2_chain.py``: the reference's call sequence over the same third-party library the reference
  uses (OpenCV).  It is used to cross-check the C restatement on fresh random inputs and as the
  "reference CPU path" timed by the benchmark.

Pinning: the reference ships no golden vectors, so both layers are pinned against
``tests/golden/*.npz``, produced by importing and running the reference itself
(``tests/golden/make_golden.py``, run in the build container where ``/root/reference`` exists).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtrs_oracle.so")
MAX_HSV = 4


class PreprocParams(C.Structure):
    _fields_ = [
        ("contrast_ratio", C.c_double), ("contrast_offset", C.c_double), ("brightness_baseline", C.c_double),
        ("dynamic_brightness", C.c_int32), ("color_filter_enabled", C.c_int32), ("n_hsv", C.c_int32),
        ("edge_enabled", C.c_int32),
        ("hsv_lo", (C.c_double * 3) * MAX_HSV), ("hsv_hi", (C.c_double * 3) * MAX_HSV),
        ("color_dest", C.c_int32 * MAX_HSV), ("edge_dest", C.c_int32),
        ("canny_a", C.c_double), ("canny_b", C.c_double),
    ]


class SpdParams(C.Structure):
    _fields_ = [
        ("threshold", C.c_double), ("reverse_multiplier", C.c_double), ("break_multiplier", C.c_double),
        ("use_break", C.c_int32), ("smooth_steering", C.c_int32), ("smooth_threshold", C.c_double),
        ("numpy_legacy_promotion", C.c_int32), ("reserved", C.c_int32),
    ]


def build(force: bool = False) -> str:
    """Compile trs_oracle.c with gcc (idempotent)."""
    src = os.path.join(_HERE, "trs_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        cmd = ["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-ffp-contract=off", "-o", _LIB_PATH, src, "-lm"]
        subprocess.check_call(cmd)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_max_threads.restype = C.c_int
    return _lib


def _ptr(a, ty):
    return a.ctypes.data_as(C.POINTER(ty)) if a is not None else None


def params_from_cfg(cfg: dict) -> PreprocParams:
    """cfg uses the reference's key names (core/config.py:15-28)."""
    p = PreprocParams()
    p.contrast_ratio = float(cfg.get("preprocessing_contrast_enhancement_ratio", 1.0))
    p.contrast_offset = float(cfg.get("preprocessing_contrast_enhancement_offset", 125))
    p.brightness_baseline = float(cfg.get("preprocessing_brightness_baseline", 550))
    p.dynamic_brightness = int(bool(cfg.get("preprocessing_dynamic_brightness_enabled", False)))
    p.color_filter_enabled = int(bool(cfg.get("preprocessing_color_filter_enabled", False)))
    hsvs = cfg.get("preprocessing_color_filter_hsvs", [((0, 0, 130), (180, 64, 255)), ((25, 180, 155), (43, 255, 255))])
    dests = cfg.get("preprocessing_color_filter_destination_channels", [0, 1])
    if p.color_filter_enabled:
        assert len(hsvs) == len(dests) and len(hsvs) <= MAX_HSV
        p.n_hsv = len(hsvs)
        for k, (lo, hi) in enumerate(hsvs):
            for c in range(3):
                p.hsv_lo[k][c] = float(lo[c])
                p.hsv_hi[k][c] = float(hi[c])
            p.color_dest[k] = int(dests[k])
    p.edge_enabled = int(bool(cfg.get("preprocessing_edge_detection_enabled", False)))
    p.edge_dest = int(cfg.get("preprocessing_edge_detection_destination_channel", 2))
    p.canny_a = float(cfg.get("preprocessing_edge_detection_threshold_a", 60))
    p.canny_b = float(cfg.get("preprocessing_edge_detection_threshold_b", 100))
    return p


# ------------------------------------------------------------------------------------------------
# stage-level entry points (numpy in, numpy out)
# ------------------------------------------------------------------------------------------------
def brightness_lut(img: np.ndarray, cfg: dict):
    img = np.ascontiguousarray(img, np.uint8)
    h, w, _ = img.shape
    lut = np.zeros(256, np.uint8)
    sums = np.zeros(3, np.uint64)
    p = params_from_cfg(cfg)
    lib().orc_brightness_lut(_ptr(img, C.c_uint8), h, w, C.byref(p), _ptr(lut, C.c_uint8), _ptr(sums, C.c_uint64))
    return lut, sums


def rgb2hsv(rgb: np.ndarray) -> np.ndarray:
    rgb = np.ascontiguousarray(rgb, np.uint8)
    out = np.empty_like(rgb)
    lib().orc_rgb2hsv(_ptr(rgb, C.c_uint8), C.c_size_t(rgb.size // 3), _ptr(out, C.c_uint8))
    return out


def hsv_tables():
    s = np.zeros(256, np.int32)
    h = np.zeros(256, np.int32)
    lib().orc_hsv_tables(_ptr(s, C.c_int32), _ptr(h, C.c_int32))
    return s, h


def inrange(hsv: np.ndarray, lo, hi) -> np.ndarray:
    hsv = np.ascontiguousarray(hsv, np.uint8)
    out = np.empty(hsv.shape[:-1], np.uint8)
    lo_a = (C.c_double * 3)(*[float(x) for x in lo])
    hi_a = (C.c_double * 3)(*[float(x) for x in hi])
    lib().orc_inrange(_ptr(hsv, C.c_uint8), C.c_size_t(hsv.size // 3), lo_a, hi_a, _ptr(out, C.c_uint8))
    return out


def canny3(rgb: np.ndarray, thr_a: float, thr_b: float, taps: bool = False):
    rgb = np.ascontiguousarray(rgb, np.uint8)
    h, w, _ = rgb.shape
    out = np.zeros((h, w), np.uint8)
    mag = np.zeros((h, w), np.uint16) if taps else None
    mp = np.zeros((h, w), np.uint8) if taps else None
    lib().orc_canny3(_ptr(rgb, C.c_uint8), h, w, C.c_double(thr_a), C.c_double(thr_b), _ptr(out, C.c_uint8),
                     _ptr(mag, C.c_uint16), _ptr(mp, C.c_uint8))
    return (out, mag, mp) if taps else out


def process_batch(frames: np.ndarray, cfg: dict, want_f32: bool = False, nthreads: int = 0):
    """ImgPreprocessing.__process over (N,H,W,3) u8 [+ /255 normalise].  Returns u8 or (u8, f32)."""
    frames = np.ascontiguousarray(frames, np.uint8)
    n, h, w, _ = frames.shape
    out = np.empty_like(frames)
    f32 = np.empty(frames.shape, np.float32) if want_f32 else None
    p = params_from_cfg(cfg)
    lib().orc_process_batch(_ptr(frames, C.c_uint8), n, h, w, C.byref(p), _ptr(out, C.c_uint8), _ptr(f32, C.c_float),
                            int(nthreads))
    return (out, f32) if want_f32 else out


def process_frame(img: np.ndarray, cfg: dict) -> np.ndarray:
    return process_batch(img[None], cfg)[0]


def normalise(u8: np.ndarray) -> np.ndarray:
    u8 = np.ascontiguousarray(u8, np.uint8)
    out = np.empty(u8.shape, np.float32)
    lib().orc_normalise(_ptr(u8, C.c_uint8), C.c_size_t(u8.size), _ptr(out, C.c_float))
    return out


def crop_resize(frames: np.ndarray, roi, out_hw, want_f32=True):
    frames = np.ascontiguousarray(frames, np.uint8)
    n, h, w, _ = frames.shape
    y0, y1, x0, x1 = roi
    ho, wo = out_hw
    u8 = np.empty((n, ho, wo, 3), np.uint8)
    f32 = np.empty((n, ho, wo, 3), np.float32) if want_f32 else None
    lib().orc_crop_resize(_ptr(frames, C.c_uint8), n, h, w, y0, y1, x0, x1, ho, wo, _ptr(u8, C.c_uint8), _ptr(f32, C.c_float))
    return (u8, f32) if want_f32 else u8


def locate(waypoints: np.ndarray, xyz: np.ndarray, min_map=0.0, max_map=10.0, nthreads: int = 0):
    wp = np.ascontiguousarray(waypoints, np.float64)
    xyz = np.ascontiguousarray(xyz, np.float64)
    n = xyz.shape[0]
    idx = np.empty(n, np.int32)
    seg = np.empty(n, np.float64)
    lib().orc_locate(_ptr(wp, C.c_double), wp.shape[0], C.c_double(min_map), C.c_double(max_map), _ptr(xyz, C.c_double), n,
                     _ptr(idx, C.c_int32), _ptr(seg, C.c_double), int(nthreads))
    return idx, seg


def spd_params_from_cfg(cfg: dict) -> SpdParams:
    p = SpdParams()
    p.threshold = float(cfg.get("spd_ctl_threshold", 1.1))
    p.reverse_multiplier = float(cfg.get("spd_ctl_reverse_multiplier", 1.0))
    p.break_multiplier = float(cfg.get("spd_ctl_break_multiplier", 1.0))
    p.use_break = int(bool(cfg.get("spd_ctl_break", False)))
    p.smooth_steering = int(bool(cfg.get("smooth_steering_enabled", False)))
    p.smooth_threshold = float(cfg.get("smooth_steering_threshold", 0.9))
    p.numpy_legacy_promotion = int(bool(cfg.get("spd_ctl_numpy_legacy_promotion", False)))
    return p


def speed_control(cur_spd, model_spd, model_steer, cfg: dict):
    cur = np.ascontiguousarray(cur_spd, np.float64)
    ms = np.ascontiguousarray(model_spd, np.float32)
    st = np.ascontiguousarray(model_steer, np.float32)
    n = cur.shape[0]
    so, th, br = (np.empty(n, np.float64) for _ in range(3))
    ft = np.empty(n, np.float32)
    p = spd_params_from_cfg(cfg)
    lib().orc_speed_control(_ptr(cur, C.c_double), _ptr(ms, C.c_float), _ptr(st, C.c_float), n, C.byref(p),
                            _ptr(so, C.c_double), _ptr(th, C.c_double), _ptr(br, C.c_double), _ptr(ft, C.c_float))
    return so, th, br, ft


def max_threads() -> int:
    return int(lib().orc_max_threads())


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
