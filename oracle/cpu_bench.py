"""Times the reference CPU path (oracle/cv2_chain.py: the reference's own call sequence over OpenCV/numpy) on the
host cores of whatever box this runs on.  TEST/BENCH INFRASTRUCTURE ONLY — used by bench.py's ``cpu_baseline`` leg
and by ``bench.py --impl reference``.  Kept free of torch so spawned workers start fast.

One worker process per core, ``cv2.setNumThreads(1)`` in each; a worker builds its own frames
from the seeded pool, warms up, then times ``frames`` passes of the chain.  Throughput = total frames / slowest worker.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def usable_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:  # pragma: no cover
        return max(1, os.cpu_count() or 1)


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def _init():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    try:
        import cv2
        cv2.setNumThreads(1)
    except Exception:
        pass


def _work(args):
    (kind, frames, h, w, cfg, worker) = args
    _init()
    import numpy as np

    from oracle import cv2_chain
    from triton_racer_sim_b200 import synth

    if kind == "frames":
        pool = synth.frame_pool(32, h, w)
        imgs = synth.expand_numpy(pool, 64, start=worker * 64)
        if cv2_chain.cv2 is None:
            import oracle
            t0 = time.perf_counter()
            oracle.process_batch(np.concatenate([imgs] * max(1, frames // 64))[:frames], cfg, want_f32=True, nthreads=1)
            return time.perf_counter() - t0
        for i in range(8):
            cv2_chain.full_chain(imgs[i].copy(), cfg)
        t0 = time.perf_counter()
        for i in range(frames):
            cv2_chain.full_chain(imgs[i & 63].copy(), cfg)          # the reference copies the frame too (img_preprocessing.py:20,29)
        return time.perf_counter() - t0
    if kind == "cars":
        # the reference's recorded centre line (car_templates/track_data/generated_track.json, 1,185 points) as committed with the goldens
        with np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "tracks.npz")) as z:
            wp = z["wp/generated_track"].tolist()
        xyz, cur, ms, st = synth.car_states(np.asarray(wp), frames, seed=worker)
        t0 = time.perf_counter()
        for k in range(frames):
            cv2_chain.locate(wp, (float(xyz[k, 0]), float(xyz[k, 1]), float(xyz[k, 2])))
            cv2_chain.pilot_tail(float(cur[k]), st[k], ms[k], cfg)
        return time.perf_counter() - t0
    raise ValueError(kind)


def run(kind: str, per_worker: int, h: int, w: int, cfg: dict, workers: int | None = None):
    """Returns dict(value=units/s, cores=workers, wall_s=..., units=...)."""
    workers = workers or usable_cores()
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers, initializer=_init) as pool:
        pool.map(_work, [("frames" if kind == "frames" else "cars", 4, h, w, cfg, i) for i in range(workers)])   # start + import cost out of the timing
        t0 = time.perf_counter()
        times = pool.map(_work, [(kind, per_worker, h, w, cfg, i) for i in range(workers)])
        wall = time.perf_counter() - t0
    slowest = max(times)
    total = per_worker * workers
    return {"value": total / slowest, "cores": workers, "wall_s": wall, "units": total, "slowest_worker_s": slowest,
            "single_core_value": per_worker / (sum(times) / len(times))}
