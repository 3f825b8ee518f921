"""Pin the oracle (C restatement + cv2 call-sequence layer) against vectors produced by the reference itself."""
import numpy as np
import pytest

import oracle
from oracle import cv2_chain
from tests.helpers import cfg_for, golden_pairs, speed_cases
from triton_racer_sim_b200 import synth


def test_c_oracle_matches_reference_images(golden_images):
    n = 0
    for sname, cname, cfg, frames, expected in golden_pairs(golden_images):
        got = oracle.process_batch(frames, cfg)
        assert np.array_equal(got, expected), f"{sname}/{cname}: {np.count_nonzero(got != expected)} bytes differ"
        n += 1
    assert n >= 30


def test_cv2_chain_matches_reference_images(golden_images):
    if cv2_chain.cv2 is None:
        pytest.skip("cv2 missing")
    for sname, cname, cfg, frames, expected in golden_pairs(golden_images):
        got = np.stack([cv2_chain.process(f.copy(), cfg) for f in frames])
        assert np.array_equal(got, expected), f"{sname}/{cname}"


def test_c_oracle_matches_cv2_on_fresh_frames():
    """Beyond the committed vectors: fresh seeded frames, several sizes, several configs."""
    if cv2_chain.cv2 is None:
        pytest.skip("cv2 missing")
    cfgs = [
        cfg_for(dict(preprocessing_color_filter_enabled=True, preprocessing_edge_detection_enabled=True)),
        cfg_for(dict(preprocessing_color_filter_enabled=True, preprocessing_edge_detection_enabled=True,
                     preprocessing_dynamic_brightness_enabled=True, preprocessing_contrast_enhancement_ratio=1.6,
                     preprocessing_contrast_enhancement_offset=90, preprocessing_edge_detection_threshold_a=33.3,
                     preprocessing_edge_detection_threshold_b=210)),
    ]
    for (h, w, seed) in [(120, 160, 1), (240, 320, 2), (17, 23, 3), (41, 64, 4), (119, 8, 5), (1, 40, 6), (40, 1, 7)]:
        frames = synth.frame_pool(8, h, w, seed=seed)
        for cfg in cfgs:
            want = np.stack([cv2_chain.process(f.copy(), cfg) for f in frames])
            got = oracle.process_batch(frames, cfg)
            assert np.array_equal(got, want), (h, w)


def test_hsv_exhaustive_against_cv2():
    """All 2^24 colours."""
    if cv2_chain.cv2 is None:
        pytest.skip("cv2 missing")
    cv2 = cv2_chain.cv2
    v = np.arange(256, dtype=np.uint8)
    for r0 in range(0, 256, 64):
        rr, gg, bb = np.meshgrid(v[r0:r0 + 64], v, v, indexing="ij")
        rgb = np.stack([rr, gg, bb], -1).reshape(64 * 256, 256, 3)
        assert np.array_equal(oracle.rgb2hsv(rgb), cv2.cvtColor(rgb, cv2.COLOR_RGB2HSV))


def test_inrange_bounds_against_cv2():
    if cv2_chain.cv2 is None:
        pytest.skip("cv2 missing")
    cv2 = cv2_chain.cv2
    rng = np.random.default_rng(3)
    hsv = rng.integers(0, 256, size=(64, 64, 3), dtype=np.uint8)
    for lo, hi in [((0, 0, 130), (180, 64, 255)), ((25, 180, 155), (43, 255, 255)), ((10.5, -3, 20.2), (99.5, 300, 128.7)),
                   ((50, 50, 50), (40, 255, 255)), ((0, 0, 0), (255, 255, 255)), ((40.5, 41.5, 2.5), (100.5, 101.5, 3.5)),
                   ((-1e10, 0, 0), (1e10, 255, 255)), ((0, 0, 255.5), (10, 10, 256)), ((-0.5, -0.6, 254.5), (255.5, 1e9, 255.4))]:
        assert np.array_equal(oracle.inrange(hsv, lo, hi), cv2.inRange(hsv, tuple(lo), tuple(hi)))


def test_normalise_is_true_division():
    u8 = np.arange(256, dtype=np.uint8)
    want = u8.astype(np.float32)
    want /= 255
    assert np.array_equal(oracle.normalise(u8), want)
    assert np.array_equal(cv2_chain.normalise(u8), want)


def test_crop_resize_against_cv2():
    if cv2_chain.cv2 is None:
        pytest.skip("cv2 missing")
    frames = synth.frame_pool(2, 240, 320)
    for (ho, wo) in [(120, 160), (200, 200), (240, 320), (60, 107)]:
        u8, f32 = oracle.crop_resize(frames, (0, 240, 0, 320), (ho, wo))
        want = np.stack([cv2_chain.resize_nearest(f, wo, ho) for f in frames])
        assert np.array_equal(u8, want)
        assert np.array_equal(f32, oracle.normalise(want))
    u8 = oracle.crop_resize(frames, (40, 119, 10, 300), (79, 290), want_f32=False)
    assert np.array_equal(u8, frames[:, 40:119, 10:300])


def test_locate_matches_reference(golden_tracks):
    for name in ("generated_track", "mountain_track"):
        wp, xyz = golden_tracks[f"wp/{name}"], golden_tracks[f"xyz/{name}"]
        idx, seg = oracle.locate(wp, xyz, 0, 10)
        assert np.array_equal(idx, golden_tracks[f"idx/{name}"])
        assert np.array_equal(seg, golden_tracks[f"seg/{name}/0_10"])
        _, seg2 = oracle.locate(wp, xyz, -2.5, 7.25)
        assert np.array_equal(seg2, golden_tracks[f"seg/{name}/m2p5_7p25"])
        # the python-loop layer agrees too (spot check)
        for k in range(0, 40):
            i, s = cv2_chain.locate(wp.tolist(), xyz[k].tolist())
            assert i == idx[k] and s == seg[k]


def test_speed_control_matches_reference(golden_speed):
    cur, ms, st = golden_speed["cur"], golden_speed["model_spd"], golden_speed["model_steer"]
    for cname, over in speed_cases(golden_speed).items():
        cfg = cfg_for(over)
        so, th, br, ft = oracle.speed_control(cur, ms, st, cfg)
        want = golden_speed[f"out/{cname}"]
        # same libm atan on the same box: expect exact equality for the C restatement
        assert np.array_equal(so, want[:, 0]), cname
        assert np.allclose(th, want[:, 1], rtol=1e-12, atol=0), cname
        assert np.allclose(br, want[:, 2], rtol=1e-12, atol=0), cname
        assert np.array_equal(ft, golden_speed["feature"])
        for k in range(0, 200):
            got = cv2_chain.pilot_tail(float(cur[k]), st[k], ms[k], cfg)
            assert got == tuple(want[k]), (cname, k)
        # NumPy 1.x scalar promotion (np.float32 * 20 -> float64): the reference functions fed float64, see make_golden.py
        so, th, br, _ = oracle.speed_control(cur, ms, st, dict(cfg, spd_ctl_numpy_legacy_promotion=True))
        want1 = golden_speed[f"out_numpy1/{cname}"]
        assert np.array_equal(so, want1[:, 0]), cname
        assert np.allclose(th, want1[:, 1], rtol=1e-12, atol=0) and np.allclose(br, want1[:, 2], rtol=1e-12, atol=0), cname
        assert np.array_equal(th == 0, want1[:, 1] == 0) and np.array_equal(br == 0, want1[:, 2] == 0), cname
