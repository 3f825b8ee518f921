"""Host-side logic that needs no GPU: sharding, workload generator, plugin API mirror, 2-rank gloo plumbing."""
import inspect
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.conftest import ROOT
from triton_racer_sim_b200 import Component, default_config, sharding, synth


def test_shard_ranges_partition_everything():
    for n in (0, 1, 7, 8, 1000, 1048576):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (s0, e0), (s1, e1) in zip(spans, spans[1:]):
                assert e0 == s1
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_synth_is_deterministic_and_shard_invariant():
    pool = synth.frame_pool(8, 24, 32)
    assert np.array_equal(pool, synth.frame_pool(8, 24, 32))
    whole = synth.expand_numpy(pool, 40)
    parts = [synth.expand_numpy(pool, e - s, start=s) for s, e in (sharding.shard_range(40, r, 3) for r in range(3))]
    assert np.array_equal(whole, np.concatenate(parts))
    t = synth.expand_torch(torch.from_numpy(pool), 40)
    assert np.array_equal(t.numpy(), whole)
    wp = synth.synthetic_track(300)
    xyz, cur, ms, st = synth.car_states(wp, 1000)
    assert xyz.shape == (1000, 3) and xyz.dtype == np.float64 and ms.dtype == np.float32 and st.dtype == np.float32
    assert len(np.unique(wp, axis=0)) < len(wp)            # duplicates present -> ties


def test_component_base_mirrors_reference_api():
    # TritonRacerSim/components/component.py:3-28
    c = Component(inputs=['a'], outputs=['b', 'c'], threaded=True)
    assert c.step_inputs == ['a'] and c.step_outputs == ['b', 'c'] and c.threaded is True
    src = ['a']
    c2 = Component(inputs=src)
    c2.step_inputs[0] = 'z'
    assert src == ['a']                                      # lists are copied (callers mutate them: manage.py:50,104)
    for hook in ('onStart', 'step', 'thread_step', 'onShutdown', 'getName'):
        assert callable(getattr(c, hook))
    assert c.getName() == 'Generic Component'
    assert c.step(1, 2) is None and c.onStart() is None
    assert list(inspect.signature(Component.__init__).parameters) == ['self', 'inputs', 'outputs', 'threaded']


def test_config_defaults_match_reference_values():
    cfg = default_config()
    # core/config.py:8-28,65-66,76-80
    assert (cfg['img_w'], cfg['img_h'], cfg['cam_resolution']) == (160, 120, [320, 240])
    assert cfg['preprocessing_contrast_enhancement_ratio'] == 1.0 and cfg['preprocessing_contrast_enhancement_offset'] == 125
    assert cfg['preprocessing_brightness_baseline'] == 550 and cfg['preprocessing_dynamic_brightness_enabled'] is False
    assert cfg['preprocessing_color_filter_hsvs'] == [((0, 0, 130), (180, 64, 255)), ((25, 180, 155), (43, 255, 255))]
    assert cfg['preprocessing_color_filter_destination_channels'] == [0, 1]
    assert (cfg['preprocessing_edge_detection_threshold_a'], cfg['preprocessing_edge_detection_threshold_b'],
            cfg['preprocessing_edge_detection_destination_channel']) == (60, 100, 2)
    assert (cfg['spd_ctl_threshold'], cfg['spd_ctl_reverse_multiplier'], cfg['spd_ctl_break'], cfg['spd_ctl_break_multiplier']) == (1.1, 1.0, False, 1.0)
    assert (cfg['smooth_steering_enabled'], cfg['smooth_steering_threshold']) == (False, 0.9)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = sharding.init_distributed("gloo")
    assert (r, w) == (rank, world)
    s, e = sharding.shard_range(n_total, rank, world)
    pool = synth.frame_pool(4, 8, 12)
    local = torch.from_numpy(synth.expand_numpy(pool, e - s, start=s))       # each rank builds only its own block
    stats = torch.zeros(16, dtype=torch.int64)
    stats[0] = e - s
    stats[5] = int(local.sum())
    total = sharding.reduce_stats(stats)
    whole = sharding.gather_shards(local, n_total)
    slowest = sharding.max_over_ranks(1.0 + rank)
    if rank == 0:
        ref = synth.expand_numpy(pool, n_total)
        q.put((int(total[0]), int(total[5]) == int(ref.astype(np.int64).sum()), bool(np.array_equal(whole.numpy(), ref)), slowest))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_stats():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 37, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    frames, sum_ok, equal, slowest = q.get(timeout=10)
    assert frames == 37 and sum_ok and equal and slowest == 2.0


def test_tub_labels_and_features_follow_the_reference_loaders():
    """keras_train.py:113-119,264-299: what each DataLoader subclass keeps of a record; float32 as at keras_train.py:48-49."""
    import numpy as np

    from triton_racer_sim_b200 import tub
    recs = [{'mux/steering': 0.25, 'mux/throttle': 0.5, 'gym/speed': 7.0, 'loc/segment': 3.5},
            {'mux/steering': -1.0, 'mux/throttle': 0.1, 'gym/speed': 20.0, 'loc/segment': 0.0}]
    lab, ft = tub.labels_and_features(recs, "DataLoader")
    assert lab.dtype == np.float32 and ft is None and np.array_equal(lab, np.float32([[0.25, 0.5], [-1.0, 0.1]]))
    lab, ft = tub.labels_and_features(recs, "cnn_2d_speed_as_feature")
    assert np.array_equal(ft, np.float32([[7.0 / 20], [1.0]])) and np.array_equal(lab[:, 1], np.float32([0.5, 0.1]))
    lab, ft = tub.labels_and_features(recs, "SpeedCtlDataLoader")
    assert ft is None and np.array_equal(lab, np.float32([[0.25, 7.0 / 20], [-1.0, 1.0]]))
    lab, ft = tub.labels_and_features(recs, "cnn_2d_full_house")
    assert np.array_equal(ft, np.float32([[7.0 / 20, 3.5], [1.0, 0.0]])) and np.array_equal(lab, np.float32([[0.25, 7.0 / 20], [-1.0, 1.0]]))
    # the same values the reference's classes produce (np.asarray(..., dtype=float32) of its tuples)
    for r in recs:
        assert np.array_equal(tub.labels_and_features([r], "FullHouseDataLoader")[1][0],
                              np.asarray(np.asarray((r['gym/speed'] / 20, r['loc/segment'])), dtype=np.float32))


def test_tub_loaders_against_the_reference_run(tmp_path):
    """tests/golden/tub.npz: a tub written by the reference's own recorder (DataStorage.step + its file thread) and `loader.dataset` of the
    reference's four loader classes over it, all run unmodified (tests/golden/make_golden_tub.py).  Labels and feature vectors must be the
    reference's float32 values bit for bit; the record count must start at 1 (the recorder numbers from 0, so its first frame is never loaded)
    and stop at the first missing file; the JPEG files must decode (Pillow here; the CUDA decoder is held to Pillow in test_jpeg_gpu.py) to the
    frames the reference divided by 255."""
    import io
    import json

    from PIL import Image

    from triton_racer_sim_b200 import tub
    g = np.load(os.path.join(ROOT, "tests", "golden", "tub.npz"))
    recorded = json.loads(bytes(g["records_json"]).decode())                    # record_0 .. record_7 as the recorder wrote them
    ends = np.cumsum(g["jpeg_sizes"])
    files = [bytes(g["jpeg_blob"][e - s:e]) for s, e in zip(g["jpeg_sizes"], ends)]
    assert len(recorded) == len(files) == 8 and recorded[3]["cam/img"] == "img_3.jpg"
    for i, (f, r) in enumerate(zip(files, recorded)):
        (tmp_path / f"img_{i}.jpg").write_bytes(f)
        (tmp_path / f"record_{i}.json").write_text(json.dumps(r))
    assert tub.count_records(str(tmp_path)) == 7                                              # records 1 .. 7
    records = recorded[1:]
    for i, f in enumerate(files[1:]):
        assert np.array_equal(np.asarray(Image.open(io.BytesIO(f))), g["frames_u8"][i]) and tub.jpeg_size(f) == g["frames_u8"].shape[1:3]
    for cname in ("DataLoader", "SpeedFeatureDataLoader", "SpeedCtlDataLoader", "FullHouseDataLoader"):
        lab, ft = tub.labels_and_features(records, cname)
        want_l, want_f = g[f"labels/{cname}"], g[f"features/{cname}"]
        assert lab.dtype == np.float32 and np.array_equal(lab, want_l), cname
        if ft is None:
            # the reference's np.asarray(None, dtype=float32) for loaders without features: a NaN scalar per record
            assert want_f.shape == (7,) and np.isnan(want_f).all(), cname
        else:
            assert ft.dtype == np.float32 and np.array_equal(ft, want_f), cname
    for model, cname in tub.LOADER_OF_MODEL.items():                                          # keras_train.py:384-398
        assert np.array_equal(tub.labels_and_features(records, model)[0], g[f"labels/{cname}"])
    (tmp_path / "record_4.json").unlink()
    assert tub.count_records(str(tmp_path)) == int(g["count_with_record_4_missing"]) == 3


# ---- nearest waypoint through the grid (csrc/loc_grid.h compiled for the host) ------------------------------------------------------------------
@pytest.fixture(scope="module")
def locate_host(tmp_path_factory):
    import ctypes as C
    import subprocess
    out = tmp_path_factory.mktemp("loc") / "liblocchk.so"
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", str(out), os.path.join(ROOT, "tests", "host_locate_check.cpp")])
    lib = C.CDLL(str(out))
    lib.locate_grid_host.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]

    def run(wp, xyz):
        wp, xyz = np.ascontiguousarray(wp, np.float64), np.ascontiguousarray(xyz, np.float64)
        idx = np.zeros(len(xyz), np.int32)
        stats = np.zeros(6, np.int64)
        assert lib.locate_grid_host(wp.ctypes.data, len(wp), xyz.ctypes.data, len(xyz), idx.ctypes.data, stats.ctypes.data) == 0
        return idx, stats
    return run


def _adversarial_cars(wp, rng, scale, n=4000):
    lo, hi = wp.min(0), wp.max(0)
    ext = max(hi[0] - lo[0], hi[2] - lo[2])
    c = 2.0 ** np.ceil(np.log2(ext / 32.0))                     # the library's cell side
    return np.ascontiguousarray(np.concatenate([
        wp[rng.integers(0, len(wp), n)] + rng.standard_normal((n, 3)) * np.array([1.5, 0.05, 1.5]) * scale,                       # near the line
        wp[rng.integers(0, len(wp), 300)],                                                                                       # on points (repeats included)
        np.stack([rng.uniform(lo[0] - 20 * scale, hi[0] + 20 * scale, n), rng.uniform(lo[1], hi[1], n),
                  rng.uniform(lo[2] - 20 * scale, hi[2] + 20 * scale, n)], 1),                                                   # anywhere around (many rings)
        np.stack([c * rng.integers(np.floor(lo[0] / c) - 2, np.floor(hi[0] / c) + 3, 1500), rng.uniform(lo[1], hi[1], 1500),
                  c * rng.integers(np.floor(lo[2] / c) - 2, np.floor(hi[2] / c) + 3, 1500)], 1),                                 # cell corners
        np.stack([c * rng.integers(np.floor(lo[0] / c), np.floor(hi[0] / c) + 1, 1500), rng.uniform(lo[1], hi[1], 1500),
                  rng.uniform(lo[2], hi[2], 1500)], 1),                                                                           # cell edges
        np.stack([lo[0] - rng.uniform(95, 105, 500), rng.uniform(lo[1], hi[1], 500), rng.uniform(lo[2], hi[2], 500)], 1),        # around the 100 limit
        np.array([[np.nan, 1, 1], [1, np.nan, 1], [1, 1, np.nan], [np.inf, 0, 0], [0, 0, -np.inf], [1e300, 1e300, 1e300], [-1e300, 0, 1e300]])]))


def test_grid_walk_returns_the_reference_index(locate_host):
    """The product's grid walk (the source the kernel compiles, built for the host) against the oracle's scan, which follows
    track_data_process.py:89-107: both shipped centre lines, shifted far from the origin and scaled to other cell sizes; cars near the line, on
    points, on cell corners and edges, anywhere around the bounding box (put off to the scan after six rings), at the 100 limit, NaN / infinite."""
    import oracle
    tracks = np.load(os.path.join(ROOT, "tests", "golden", "tracks.npz"))
    rng = np.random.default_rng(2024)
    for name in ("generated_track", "mountain_track"):
        base = tracks[f"wp/{name}"]
        # the reference's own outputs first
        idx, stats = locate_host(base, tracks[f"xyz/{name}"])
        assert np.array_equal(idx, tracks[f"idx/{name}"])
        assert stats[0] == 1 and stats[5] == 2 and stats[1] <= 34 and stats[2] <= 34          # cells of side 4
        put_off = 0
        for scale, shift in ((1.0, 0.0), (1.0, 1.0e6), (1e-3, 0.0), (1e3, -5.0e4), (1.0, -3.0)):
            wp = base * scale + shift
            xyz = _adversarial_cars(wp, rng, scale)
            with np.errstate(invalid="ignore", over="ignore"):
                want, _ = oracle.locate(wp, xyz)
            got, stats = locate_host(wp, xyz)
            bad = np.nonzero(got != want)[0]
            assert len(bad) == 0, f"{name} x{scale} +{shift}: {len(bad)} of {len(want)} differ, first {bad[:5]}: {xyz[bad[:5]]}"
            put_off += int(stats[3])
        assert put_off > 0                                       # the deferred path is exercised
    # a nearest point at a distance of exactly 100 inside the bounding box: the reference's strict `<` against its starting minimum keeps index 0
    wp = np.array([[i, 0, 0] for i in range(40)] + [[1000 + i, 0, 1000] for i in range(40)], np.float64)
    xyz = np.array([[89, 0, 50], [89, 0, 49.999], [89, 0, 50.001], [88, 1, 50], [1000, -50, 950], [1039, 0, 1100]], np.float64)
    want = oracle.locate(wp, xyz)[0]
    assert list(want) == [0, 39, 0, 0, 0, 0]
    got, stats = locate_host(wp, xyz)
    assert stats[0] == 1 and stats[5] == 6 and np.array_equal(got, want)
    # no grid: a handful of points, a non-finite point
    wp = synth.synthetic_track(40)
    xyz, _, _, _ = synth.car_states(wp, 2000, seed=8)
    got, stats = locate_host(wp, xyz)
    assert stats[0] == 0 and np.array_equal(got, oracle.locate(wp, xyz)[0])
    wp = tracks["wp/generated_track"].copy()
    wp[17, 1] = np.inf
    xyz, _, _, _ = synth.car_states(tracks["wp/generated_track"], 2000, seed=9)
    got, stats = locate_host(wp, xyz)
    with np.errstate(invalid="ignore"):
        assert stats[0] == 0 and np.array_equal(got, oracle.locate(wp, xyz)[0])


# ---- bench.py contract pieces that run without a GPU ----------------------------------------------------------------------------------------------
def test_reference_arm_prints_a_contract_line():
    """`bench.py --impl reference` times the reference's call sequence (the port under oracle/) on the host cores and prints ONE JSON line with
    the contract's keys; without a GPU the product arm refuses to run instead of falling back."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["metric"].startswith("preprocessed frames/sec (120x160)") and d["config"]["workload"] == "full_chain_120x160"
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    if not torch.cuda.is_available():
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                             timeout=600, cwd=ROOT)
        assert not [ln for ln in out.stdout.splitlines() if ln.startswith("{")]          # no number without the GPU
        assert "no CPU path" in (out.stdout + out.stderr)


def test_grid_walk_fuzz_under_the_sanitizers(tmp_path):
    """csrc/loc_grid.h built for the host with -fsanitize=address,undefined: 300 random centre lines x 400 cars against the reference's loop
    (tests/host_locate_fuzz.cpp) — exact indices, no out-of-bounds cell access, no integer overflow in the ring arithmetic."""
    import subprocess
    exe = tmp_path / "loc_fuzz"
    build = subprocess.run(["g++", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer", "-o", str(exe),
                            os.path.join(ROOT, "tests", "host_locate_fuzz.cpp")], capture_output=True, text=True)
    if build.returncode != 0:
        pytest.skip("no sanitizer runtime in this toolchain: " + build.stderr[-200:])
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0 and run.stdout.strip() == "ok 120000 cars", (run.stdout[-500:], run.stderr[-2000:])


# ---- per-pixel arithmetic of the kernels (csrc/pixel_math.cuh compiled for the host) ---------------------------------------------------------------
@pytest.fixture(scope="module")
def math_host(tmp_path_factory):
    import ctypes as C
    import subprocess
    out = tmp_path_factory.mktemp("math") / "libmathchk.so"
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", str(out), os.path.join(ROOT, "tests", "host_math_check.cpp")])
    lib = C.CDLL(str(out))
    lib.math_rgb2hsv_range.argtypes = [C.c_uint, C.c_uint, C.c_void_p]
    lib.math_in_range.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.math_canny_dirs.argtypes = [C.c_int, C.c_void_p]
    lib.math_adjust_table.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p]
    lib.math_brightness_delta.argtypes = [C.c_ulonglong, C.c_ulonglong, C.c_ulonglong, C.c_double, C.c_double]
    lib.math_brightness_delta.restype = C.c_double
    lib.math_check_sat_thresholds.argtypes = [C.c_int, C.c_int]
    lib.math_check_sat_thresholds.restype = C.c_longlong
    lib.math_check_hue_thresholds.argtypes = [C.c_int, C.c_int]
    lib.math_check_hue_thresholds.restype = C.c_longlong
    return lib


def test_pixel_math_header_against_opencv_and_the_oracle(math_host):
    """The header the kernels compile, built for the host: RGB -> HSV over all 2^24 colours against cv2.cvtColor itself, the direction classes of
    all gradients in [-400, 400]^2 against the angle they stand for, the brightness / contrast table and the dynamic-brightness statistic against
    the oracle (which test_oracle_golden.py holds to the reference's own outputs)."""
    import cv2

    import oracle
    from tests.helpers import cfg_for
    # (1) every colour, in slabs of 2^20
    for first in range(0, 1 << 24, 1 << 20):
        got = np.empty((1 << 20, 3), np.uint8)
        math_host.math_rgb2hsv_range(first, 1 << 20, got.ctypes.data)
        c = np.arange(first, first + (1 << 20), dtype=np.uint32)
        rgb = np.stack([(c >> 16) & 255, (c >> 8) & 255, c & 255], 1).astype(np.uint8)
        want = cv2.cvtColor(rgb.reshape(1024, 1024, 3), cv2.COLOR_RGB2HSV).reshape(-1, 3)
        assert np.array_equal(got, want), first
    # (2) direction classes: 0 within 22.5 degrees of the x axis, 1 within 22.5 degrees of the y axis, else the diagonal with the sign of dx dy;
    # gradients within 0.01 degrees of a sector boundary are left to the integer rule (tan 22.5 as 13573 / 2^15)
    m = 400
    dirs = np.empty((2 * m + 1, 2 * m + 1), np.uint8)
    math_host.math_canny_dirs(m, dirs.ctypes.data)
    dy, dx = np.mgrid[-m:m + 1, -m:m + 1]
    ang = np.degrees(np.arctan2(np.abs(dy), np.abs(dx)))
    want = np.where(ang < 22.5, 0, np.where(ang > 67.5, 1, np.where((dx ^ dy) < 0, 3, 2)))
    clear = (np.abs(ang - 22.5) > 0.01) & (np.abs(ang - 67.5) > 0.01) & ((dx != 0) | (dy != 0))      # (the zero gradient never passes the magnitude test)
    assert np.array_equal(dirs[clear], want[clear])
    # (3) brightness / contrast tables against the oracle's, static and dynamic
    rng = np.random.default_rng(12)
    from triton_racer_sim_b200 import synth
    for over in (dict(preprocessing_contrast_enhancement_ratio=1.3, preprocessing_contrast_enhancement_offset=100.5),
                 dict(preprocessing_dynamic_brightness_enabled=True, preprocessing_brightness_baseline=400.25, preprocessing_contrast_enhancement_ratio=0.7),
                 dict(preprocessing_dynamic_brightness_enabled=True, preprocessing_brightness_baseline=550)):
        cfg = cfg_for(over)
        for img in synth.frame_pool(6, 120, 160, seed=int(rng.integers(1000))):
            lut, sums = oracle.brightness_lut(img, cfg)
            dyn = bool(cfg["preprocessing_dynamic_brightness_enabled"])
            delta = math_host.math_brightness_delta(int(sums[0]), int(sums[1]), int(sums[2]), 79.0 * 160, float(cfg["preprocessing_brightness_baseline"]))
            got = np.zeros(256, np.uint8)
            math_host.math_adjust_table(int(dyn), np.float32(delta), np.float32(cfg["preprocessing_contrast_enhancement_offset"]),
                                        np.float32(cfg["preprocessing_contrast_enhancement_ratio"]), got.ctypes.data)
            assert np.array_equal(got, lut), over


def test_threshold_tables_replace_saturation_and_hue_exactly(math_host):
    """The fast kernels never compute the saturation, and for the reference's default ranges not the hue either: they compare the delta / the hue
    numerator with per-value thresholds (csrc/pixel_math.cuh sat_threshold_entry / hue_threshold_entry, the functions init_tables calls).  Exhaustive
    proof on the host build: every (value, delta) pair against every saturation bound from -2 to 300, and every (delta, numerator) pair against every
    hue interval with a lower bound >= 0 and an upper bound up to 149 (the variant's conditions: both hue bounds live, so the lower one is positive;
    the host picks it for upper bounds below 60 only)."""
    assert math_host.math_check_sat_thresholds(-2, 300) == 0
    for lo in list(range(0, 151, 7)) + [1, 25]:                  # (a live lower bound is positive; a negative one would admit wrapped hues)
        for hi in (0, 12, 43, 59, 90, 120, 149):
            assert math_host.math_check_hue_thresholds(lo, hi) == 0, (lo, hi)
