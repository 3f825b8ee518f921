"""CPU oracle for the per-car control post-processing.  TEST INFRASTRUCTURE ONLY.

numpy restatement of three scalar components of the reference, vectorised over N cars:

* ``control_mux``   ControlMultiplexer.step            components/controlmultiplexer.py:24-43 (+ launch locks 48-70)
* ``driver_assist`` DriverAssistance.step              components/driver_assistance.py:13-31
* ``three_segment_map``                                utils/mapping.py:9-16 (cap: 18-21)

The reference ends its launch locks from sleeping threads (controlmultiplexer.py:48-70): every AI launch sets the flag and starts a
thread that clears it ``duration`` seconds later, whatever happened in between.  A re-launch while an older thread is still sleeping is
therefore cut short by that older thread (the flag is cleared at the OLDEST pending end time after the latest launch, and stays clear
until the next launch).  Here the clock is an argument and the state is the last ``LAUNCH_SLOTS`` launch times per car, which reproduces
that exactly while at most ``LAUNCH_SLOTS`` lock threads would be pending at once.  Pinned against tests/golden/control.npz, produced by
running the reference's own classes in real time (tests/golden/make_golden_control.py).
"""
import numpy as np

MODE_HUMAN, MODE_AI_STEERING, MODE_AI = 0, 1, 2          # DriveMode.HUMAN / AI_STEERING / AI (components/controller.py:7-10)
NEVER = -1.0e300


LAUNCH_SLOTS = 4


def lock_active(launch_times, now, duration):
    """launch_times (K, N): flag set at the latest launch L, cleared at the earliest end time t_i + duration that lies after L."""
    latest = launch_times.max(axis=0)
    ends = launch_times + float(duration)
    pending = np.where((launch_times > NEVER) & (ends > latest), ends, np.inf)
    return (latest > NEVER) & (now < pending.min(axis=0))


def control_mux(mode, usr, ai, now, last_mode, launch_times, cfg):
    """mode (N,) int; usr, ai (3, N) f64 = steering, throttle, breaking; last_mode (N,) and launch_times (LAUNCH_SLOTS, N) are the
    per-car state, updated in place.  Returns (3, N) f64."""
    mode = np.asarray(mode)
    usr = np.asarray(usr, np.float64)
    ai = np.asarray(ai, np.float64)
    out = np.empty_like(usr)
    human, full = mode == MODE_HUMAN, mode == MODE_AI
    out[0] = np.where(human, usr[0], ai[0])                     # :26-31  steering from the pilot in both AI modes
    out[1] = np.where(full, ai[1], usr[1])
    out[2] = np.where(full, ai[2], usr[2])
    launch = (last_mode != MODE_AI) & full                      # :33  AI launch detection
    if launch.any():                                            # :34-35  both locks start together: the oldest remembered launch drops out
        oldest = launch_times.argmin(axis=0)
        cols = np.nonzero(launch)[0]
        launch_times[oldest[cols], cols] = now
    if cfg['ai_launch_lock_steering_enabled']:                  # :37-38
        out[0] = np.where(lock_active(launch_times, now, cfg['ai_launch_lock_steering_duration']), float(cfg['ai_launch_lock_steering_value']), out[0])
    if cfg['ai_launch_boost_throttle_enabled']:                 # :39-40
        out[1] = np.where(lock_active(launch_times, now, cfg['ai_launch_boost_throttle_duration']), float(cfg['ai_launch_boost_throttle_value']), out[1])
    last_mode[:] = mode                                         # :42
    return out


def driver_assist(steering, throttle, breaking, speed, cfg):
    """driver_assistance.py:13-31 for N cars (all inputs present)."""
    st = np.array(steering, np.float64)
    th = np.array(throttle, np.float64)
    br = np.array(breaking, np.float64)
    sp = np.asarray(speed, np.float64)
    k = float(cfg['drive_assist_limit_k'])
    if cfg['drive_assist_limit_mode'] == 'steering':
        ok = sp != 0
        with np.errstate(divide='ignore', invalid='ignore'):
            mx = k / sp
        hi = ok & (st > mx)                                      # :18-20
        lo = ok & ~hi & (st < mx * -1)                           # :21-23
        st = np.where(hi, mx, np.where(lo, mx * -1, st))
        th = np.where(hi | lo, -0.1, th)
    elif cfg['drive_assist_limit_mode'] == 'speed':
        ok = st != 0
        with np.errstate(divide='ignore', invalid='ignore'):
            mx = k / st
        cut = ok & (sp > mx)                                     # :27-29
        th = np.where(cut, 0.0, th)
        br = np.where(cut, 0.0, br)
    return st, th, br


def three_segment_map(val, min_map, mid_map, max_map):
    """utils/mapping.py:9-16: cap to [-1, 1]; 0 -> mid; negative side scaled by (mid - min), positive by (max - mid)."""
    v = np.clip(np.asarray(val, np.float64), -1.0, 1.0)
    neg = mid_map + (mid_map - min_map) * v
    pos = mid_map + (max_map - mid_map) * v
    return np.where(v == 0, float(mid_map), np.where(v < 0, neg, pos))
