"""The pilots' networks on the GPU, behind the reference's ``KerasPilot`` plugin interface.

``PilotNet`` is the batched forward pass of Keras_2D_CNN / Keras_2D_FULL_HOUSE (TritonRacerSim/components/keras_train.py:127-245) on
the tcgen05 tensor cores (csrc/pilot_kernels.cuh); ``KerasPilot`` mirrors TritonRacerSim/components/keras_pilot.py:16-153 — same
constructor arguments, input / output keys and per-model-type glue — with every input gaining a leading N.

Weights: a dict of numpy arrays in Keras layout under the reference's layer names ("conv1/kernel", "dense1/bias", …), an ``.npz`` file
of the same, or (when h5py is importable) the ``.h5`` file Keras wrote.  There is no CPU evaluation: without the CUDA library the
constructor raises.
"""
from __future__ import annotations

import ctypes as C
from enum import Enum

import numpy as np
import torch

from . import _native as nat
from .component import Component
from .config import default_config


class ModelType(Enum):
    """TritonRacerSim/utils/types.py:3-9 (the 2-D models; CNN_3D / RNN have no model in the reference)."""
    CNN_2D = 'cnn_2d'
    CNN_2D_SPD_FTR = 'cnn_2d_speed_as_feature'
    CNN_2D_SPD_CTL = 'cnn_2d_speed_control'
    CNN_2D_FULL_HOUSE = 'cnn_2d_full_house'


_KIND = {ModelType.CNN_2D: 0, ModelType.CNN_2D_SPD_FTR: 1, ModelType.CNN_2D_SPD_CTL: 2, ModelType.CNN_2D_FULL_HOUSE: 3}


def _model_type(m) -> ModelType:
    if isinstance(m, ModelType):
        return m
    v = getattr(m, "value", m)                      # the reference's own enum, or its string value
    return ModelType(v)


def load_weights(path) -> dict:
    """name -> float32 array.  ``.npz``: keys are the tensor names; ``.h5``: Keras's model_weights group (needs h5py)."""
    path = str(path)
    if path.endswith(".npz"):
        with np.load(path) as z:
            return {k: np.asarray(z[k], np.float32) for k in z.files}
    try:
        import h5py
    except ImportError as e:                                              # pragma: no cover
        raise RuntimeError(f"{path}: reading Keras .h5 files needs h5py, which is not installed; export the weights to .npz") from e
    out = {}                                                              # pragma: no cover
    with h5py.File(path, "r") as f:                                       # pragma: no cover
        g = f["model_weights"] if "model_weights" in f else f

        def visit(name, obj):
            if isinstance(obj, h5py.Dataset):
                parts = name.split("/")
                out[f"{parts[-2]}/{parts[-1].split(':')[0]}"] = np.asarray(obj, np.float32)
        g.visititems(visit)
    return out


class PilotNet:
    """Batched forward pass: (N,H,W,3) u8 frames [+ speed / 20, + loc/segment] -> (N,2) float32 model outputs."""

    LAYERS = 8          # debug taps: 1..7 conv outputs, 8 first-Dense partial sums

    def __init__(self, model_type, weights, h=120, w=160, device=None, max_batch=16384, initial_batch=64):
        """max_batch: largest internal chunk (frames per launch sequence); the activation workspace (~0.6 MB per 120x160 frame) starts
        at `initial_batch` frames and grows to the largest batch seen, up to max_batch, so a single-car pilot stays small."""
        self.model_type = _model_type(model_type)
        self.device = torch.cuda.current_device() if device is None else torch.device(device).index if not isinstance(device, int) else device
        self.ctx = nat.Context(self.device)
        self.h, self.w = int(h), int(w)
        if not isinstance(weights, dict):
            weights = load_weights(weights)
        self._weights = {k: np.ascontiguousarray(v, np.float32) for k, v in weights.items()}      # kept for workspace growth
        for nm, v in self._weights.items():
            if v.ndim > 4:
                raise ValueError(f"weight {nm!r} has {v.ndim} dimensions")
        self.max_batch = int(max_batch)
        self.handle = None
        self.capacity = 0
        self._create(min(self.max_batch, max(1, int(initial_batch))))

    def _create(self, capacity):
        names = [k.encode() for k in self._weights]
        arr = (nat.Tensor * len(names))()
        for i, (nm, v) in enumerate(zip(names, self._weights.values())):
            arr[i].name = nm
            arr[i].data = v.ctypes.data_as(C.POINTER(C.c_float))
            arr[i].ndim = v.ndim
            for d in range(v.ndim):
                arr[i].shape[d] = v.shape[d]
        h_ = C.c_void_p()
        nat.check(self.ctx.lib.trs_pilot_create(self.ctx.handle, _KIND[self.model_type], self.h, self.w, arr, len(names), int(capacity),
                                                C.byref(h_)), "trs_pilot_create")
        if self.handle:
            torch.cuda.synchronize(self.device)          # work queued on the old workspace
            self.ctx.lib.trs_pilot_destroy(self.handle)
        self.handle, self.capacity = h_, int(capacity)

    def reserve(self, n):
        """Make the workspace hold chunks of min(n, max_batch) frames."""
        want = min(int(n), self.max_batch)
        if want > self.capacity:
            self._create(want)

    def forward_device(self, frames: torch.Tensor, spd_feature: torch.Tensor = None, loc_feature: torch.Tensor = None, out=None):
        if not (isinstance(frames, torch.Tensor) and frames.is_cuda and frames.device.index == self.device and frames.dtype == torch.uint8
                and frames.dim() == 4 and tuple(frames.shape[1:]) == (self.h, self.w, 3)):
            raise ValueError(f"frames must be (N,{self.h},{self.w},3) uint8 on cuda:{self.device}")
        frames = frames.contiguous()
        n = frames.shape[0]
        if out is not None and not (out.is_cuda and out.device.index == self.device and out.dtype == torch.float32
                                    and tuple(out.shape) == (n, 2) and out.is_contiguous()):
            raise ValueError(f"out must be a contiguous ({n}, 2) float32 tensor on cuda:{self.device}")
        self.reserve(n)
        if out is None:
            out = torch.empty((n, 2), dtype=torch.float32, device=frames.device)
        f32 = lambda t: None if t is None else t.to(device=frames.device, dtype=torch.float32).contiguous()
        spd_feature, loc_feature = f32(spd_feature), f32(loc_feature)
        for t, name in ((spd_feature, "speed feature"), (loc_feature, "loc/segment feature")):
            if t is not None and t.numel() != n:
                raise ValueError(f"{name} must have {n} elements, got {t.numel()}")
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        stream = C.c_void_p(torch.cuda.current_stream(frames.device).cuda_stream)
        nat.check(self.ctx.lib.trs_pilot_forward(self.handle, ptr(frames), n, ptr(spd_feature), ptr(loc_feature), ptr(out), stream),
                  "trs_pilot_forward")
        return out

    def layer_shape(self, layer):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        nat.check(self.ctx.lib.trs_pilot_layer_shape(self.handle, layer, C.byref(a), C.byref(b), C.byref(c)), "trs_pilot_layer_shape")
        return a.value, b.value, c.value

    def activation(self, layer, n):
        """Debug tap: the first n frames of the most recent chunk at `layer` as a numpy array (fp16 NHWC; layer 8: fp32)."""
        ho, wo, c = self.layer_shape(layer)
        buf = np.empty((n, ho, wo, c), np.float32 if layer == 8 else np.float16)
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        nat.check(self.ctx.lib.trs_pilot_debug_activation(self.handle, layer, buf.ctypes.data_as(C.c_void_p), buf.nbytes, stream),
                  "trs_pilot_debug_activation")
        return buf

    def close(self):
        if getattr(self, "handle", None):
            self.ctx.lib.trs_pilot_destroy(self.handle)
            self.handle = None
            self.ctx.close()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


def _is_ai(mode):
    v = getattr(mode, "value", mode)
    return v in ('ai_steering', 'ai', 1, 2)           # DriveMode.AI_STEERING / DriveMode.AI (components/controller.py:7-10)


class KerasPilot(Component):
    """Drop-in for TritonRacerSim/components/keras_pilot.py:16-153, N cars per call.

    step(cam/img (N,H,W,3) u8, gym/speed (N,), loc/segment (N,), gym/cte, usr/mode) -> ai/steering, ai/throttle, ai/breaking.
    Tensors in, tensors out (float64, like the reference's Python floats); one numpy frame and Python scalars in (the reference's own
    call), Python floats out.  ``usr/mode`` is one mode for the whole batch, as one pilot serves it.
    """

    def __init__(self, cfg, model_path, model_type, device=None, max_batch=16384):
        inputs = ['cam/img', 'gym/speed', 'loc/segment', 'gym/cte', 'usr/mode']                   # keras_pilot.py:18-19
        outputs = ['ai/steering', 'ai/throttle', 'ai/breaking']
        Component.__init__(self, inputs=inputs, outputs=outputs, threaded=False)
        self.cfg = default_config()
        self.cfg.update(cfg or {})
        self.model_type = _model_type(model_type)
        h, w = int(self.cfg.get('img_h', 120)), int(self.cfg.get('img_w', 160))
        self.model = PilotNet(self.model_type, model_path, h, w, device=device, max_batch=max_batch)
        self.device = self.model.device
        self.spd_params = nat.spd_params_from_cfg(self.cfg)
        self.smooth_steering = bool(self.cfg['smooth_steering_enabled'])
        self.smooth_steering_threshold = float(self.cfg['smooth_steering_threshold'])
        self.on = True

    def pilot_device(self, frames: torch.Tensor, speed: torch.Tensor = None, segment: torch.Tensor = None):
        """Batched body of KerasPilot.step (keras_pilot.py:48-117): three float64 tensors (N,)."""
        if not (isinstance(frames, torch.Tensor) and frames.is_cuda and frames.device.index == self.device):
            raise ValueError(f"frames must live on cuda:{self.device}")
        n = frames.shape[0]
        dev = frames.device
        mt = self.model_type
        if mt != ModelType.CNN_2D and (speed is None or speed.numel() != n):
            raise ValueError(f"gym/speed must hold {n} values for this model type")
        lib, ctx = self.model.ctx.lib, self.model.ctx.handle
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        spd_feat = None
        if mt in (ModelType.CNN_2D_SPD_FTR, ModelType.CNN_2D_FULL_HOUSE):
            # spd = np.asarray(real_spd / 20, float32): float64 division, then one rounding (keras_pilot.py:68,100)
            spd_feat = (speed.to(device=dev, dtype=torch.float64) / 20).to(torch.float32)
        out = self.model.forward_device(frames, spd_feat, segment if mt == ModelType.CNN_2D_FULL_HOUSE else None)
        res = torch.empty((3, n), dtype=torch.float64, device=dev)
        p = lambda t: C.c_void_p(t.data_ptr())
        if mt in (ModelType.CNN_2D, ModelType.CNN_2D_SPD_FTR):                                     # keras_pilot.py:59-63, 71-76
            nat.check(lib.trs_pilot_cap(ctx, p(out), n, int(self.smooth_steering), self.smooth_steering_threshold, p(res[0]), p(res[1]),
                                        p(res[2]), stream), "trs_pilot_cap")
            return res[0], res[1], res[2]
        cur = speed.to(device=dev, dtype=torch.float64).contiguous()
        steer, mspd = out[:, 0].contiguous(), out[:, 1].contiguous()
        nat.check(lib.trs_speed_control(ctx, p(cur), p(mspd), p(steer), n, C.byref(self.spd_params), p(res[0]), p(res[1]), p(res[2]),
                                        None, stream), "trs_speed_control")
        return res[0], res[1], res[2]

    def step(self, *args):
        if args[0] is None:
            return 0.0, 0.0, 0.0                                                                  # keras_pilot.py:46-47
        if not _is_ai(args[-1]):
            return 0.0, 0.0, 0.0                                                                  # keras_pilot.py:118
        img, speed, segment = args[0], args[1], args[2]
        dev = f"cuda:{self.device}"
        if isinstance(img, torch.Tensor) and img.dim() == 4:
            return self.pilot_device(img, speed, segment)
        img = np.asarray(img, np.uint8)
        single = img.ndim == 3
        frames = torch.from_numpy(img.reshape((-1,) + img.shape[-3:])).to(dev)
        as_t = lambda v: None if v is None else torch.as_tensor(np.atleast_1d(np.asarray(v, np.float64)), device=dev)
        s, t, b = self.pilot_device(frames, as_t(speed), as_t(segment))
        res = torch.stack([s, t, b]).cpu().numpy()                      # one device -> host copy for the three outputs
        if single:
            return float(res[0, 0]), float(res[1, 0]), float(res[2, 0])
        return res[0], res[1], res[2]

    def onStart(self):
        if self.cfg.get('preprocessing_enabled'):
            print('[WARNING] Image preprocessing is enabled. Autopilot is fed with FILTERED image.')   # keras_pilot.py:131-133

    def onShutdown(self):
        self.on = False
        self.model.close()

    def getName(self):
        return 'Keras Pilot'
