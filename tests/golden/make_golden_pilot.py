#!/usr/bin/env python
"""Generate tests/golden/pilot.npz by running the reference's own model-building code and its own KerasPilot.step
(TritonRacerSim/components/keras_train.py:127-245 and components/keras_pilot.py:17-153, imported UNMODIFIED from /root/reference)
on top of a numpy stand-in for the handful of Keras primitives they use.

TensorFlow is not installed in this image, so the reference's models cannot be built with the real Keras.  What CAN be executed is
everything the reference itself wrote: which layers exist, their names, filter counts, kernel sizes, strides and activations, the
order of the Concatenate inputs, which tensor feeds which head, the order of the model's inputs and outputs, the `/255`, the
reshapes, the tuple `(img, spd, features)` handed to the model, the cap / smoothing / speed-control tail.  The stand-in below
(`tensorflow` injected into sys.modules) only supplies the PRIMITIVES, restated from their Keras documentation:

  Input(shape, name)                         placeholder
  Conv2D(filters, kernel_size, strides, activation, name)   padding 'valid', channels_last, kernel (kh, kw, in, out), + bias
  Dense(units, activation, name)             x @ kernel (in, out) + bias
  Dropout(rate)                              identity at inference
  Flatten(name)                              row-major reshape of (H, W, C) per sample
  Concatenate(axis=1)                        np.concatenate
  Model(inputs, outputs)                     callable on one array or a tuple in the order of `inputs` (an input one rank short of a
                                             reference input ending in 1 gets that axis, as Keras does); .numpy() on the result
  load_model(path, compile)                  returns the model registered under `path`

Arithmetic is float64 inside the stand-in (the fixture is the exact value up to float32 inputs and weights), outputs float32.
So the fixture pins the GRAPH and the pilot's glue code to the reference's executed source; it does not pin Keras' own float32
kernels (parity of the CUDA path stays tolerance-based, tests/test_pilot_gpu.py).

Weights are oracle.pilot_ref.random_weights (deterministic), assigned by the reference's layer names; frames are synth.frame_pool.
Run in the build container: ``python tests/golden/make_golden_pilot.py``.
"""
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)


# ---- the stand-in for the Keras primitives -----------------------------------------------------------------------------------------------
class Node:
    def __init__(self, layer, parents):
        self.layer, self.parents = layer, parents


class Layer:
    def __init__(self, name=None):
        self.name = name
        self.weights = None                       # [kernel, bias] float32, Keras layout

    def __call__(self, x):
        parents = list(x) if isinstance(x, (list, tuple)) else [x]
        return Node(self, parents)


def _act(name, y):
    if name == 'relu':
        return np.maximum(y, 0.0)
    assert name in (None, 'linear'), name
    return y


class Input(Layer):
    def __init__(self, shape, name=None):
        super().__init__(name)
        self.shape = tuple(shape)

    def __new__(cls, shape, name=None):            # Input(...) returns the tensor, not a layer
        layer = object.__new__(cls)
        Layer.__init__(layer, name)
        layer.shape = tuple(shape)
        return Node(layer, [])


class Conv2D(Layer):
    def __init__(self, filters, kernel_size, strides=(1, 1), activation=None, name=None, padding='valid'):
        super().__init__(name)
        assert padding == 'valid'
        self.filters, self.kernel_size, self.strides, self.activation = filters, tuple(kernel_size), tuple(strides), activation

    def run(self, xs):
        x, = xs                                    # (N, H, W, C) float64
        k, b = (w.astype(np.float64) for w in self.weights)
        kh, kw, cin, cout = k.shape
        assert (kh, kw) == self.kernel_size and cout == self.filters and cin == x.shape[3], self.name
        sh, sw = self.strides
        n, h, w, _ = x.shape
        ho, wo = (h - kh) // sh + 1, (w - kw) // sw + 1
        y = np.zeros((n, ho, wo, cout))
        for i in range(kh):
            for j in range(kw):
                y += x[:, i:i + sh * (ho - 1) + 1:sh, j:j + sw * (wo - 1) + 1:sw, :] @ k[i, j]
        return _act(self.activation, y + b)


class Dense(Layer):
    def __init__(self, units, activation=None, name=None):
        super().__init__(name)
        self.units, self.activation = units, activation

    def run(self, xs):
        x, = xs
        k, b = (w.astype(np.float64) for w in self.weights)
        assert k.shape == (x.shape[1], self.units), (self.name, k.shape, x.shape)
        return _act(self.activation, x @ k + b)


class Dropout(Layer):
    def __init__(self, rate, name=None):
        super().__init__(name)

    def run(self, xs):
        return xs[0]


class Flatten(Layer):
    def run(self, xs):
        return xs[0].reshape(xs[0].shape[0], -1)


class Concatenate(Layer):
    def __init__(self, axis=-1, name=None):
        super().__init__(name)
        self.axis = axis

    def run(self, xs):
        return np.concatenate(xs, axis=self.axis)


class Result:
    def __init__(self, arr):
        self.arr = arr

    def numpy(self):
        return self.arr


class Model:
    calls = []                                     # (model, float32 output) of every call, for the fixture

    def __init__(self, inputs, outputs):
        self.inputs, self.outputs = list(inputs), list(outputs)
        assert len(self.outputs) == 1

    def layers_by_name(self):
        out, seen, stack = {}, set(), list(self.outputs)
        while stack:
            nd = stack.pop()
            if id(nd) in seen:
                continue
            seen.add(id(nd))
            if nd.layer.name:
                out[nd.layer.name] = nd.layer
            stack.extend(nd.parents)
        return out

    def __call__(self, x):
        xs = list(x) if isinstance(x, (list, tuple)) else [x]
        assert len(xs) == len(self.inputs), "as many arrays as the model has inputs"
        env = {}
        for nd, arr in zip(self.inputs, xs):
            arr = np.asarray(arr)
            # Keras' functional models conform an input of one rank less to a reference input whose last dimension is 1
            # (functional.py, _conform_to_reference_input): keras_pilot.py:100-101 hands the speed over as shape (1,) for Input(shape=(1,))
            if arr.ndim == len(nd.layer.shape) and nd.layer.shape[-1] == 1:
                arr = arr[..., None]
            assert arr.dtype == np.float32 and arr.shape[1:] == nd.layer.shape, (nd.layer.name, arr.shape, nd.layer.shape)
            env[id(nd)] = arr.astype(np.float64)

        def ev(nd):
            if id(nd) not in env:
                env[id(nd)] = nd.layer.run([ev(p) for p in nd.parents])
            return env[id(nd)]
        out = ev(self.outputs[0]).astype(np.float32)
        Model.calls.append((self, out))
        return Result(out)

    def summary(self):
        pass


REGISTRY = {}


def load_model(path, compile=True):
    return REGISTRY[path]


def install_stand_in():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    backend = mod("tensorflow.keras.backend", set_learning_phase=lambda flag: None)
    layers = mod("tensorflow.keras.layers", Input=Input, Conv2D=Conv2D, Dense=Dense, Dropout=Dropout, Flatten=Flatten, Concatenate=Concatenate)
    mod("tensorflow.keras.layers.experimental")
    mod("tensorflow.keras.layers.experimental.preprocessing", Rescaling=object)
    optimizers, losses = mod("tensorflow.keras.optimizers"), mod("tensorflow.keras.losses")
    models = mod("tensorflow.keras.models", Model=Model, load_model=load_model)
    keras = mod("tensorflow.keras", backend=backend, layers=layers, optimizers=optimizers, losses=losses, models=models)
    mod("tensorflow.python")
    mod("tensorflow.python.keras")
    mod("tensorflow.python.keras.models", load_model=load_model)
    mod("tensorflow", keras=keras)
    sys.modules.setdefault("pygame", types.ModuleType("pygame"))      # components/controller.py imports it for the joystick classes


install_stand_in()
from TritonRacerSim.components import keras_pilot as ref_pilot  # noqa: E402
from TritonRacerSim.components import keras_train as ref_train  # noqa: E402
from TritonRacerSim.components.controller import DriveMode  # noqa: E402
from TritonRacerSim.core.config import config as REF_DEFAULTS  # noqa: E402
from TritonRacerSim.utils.types import ModelType  # noqa: E402

from oracle import pilot_ref  # noqa: E402
from triton_racer_sim_b200 import synth  # noqa: E402

MODEL_TYPES = [(pilot_ref.CNN_2D, ModelType.CNN_2D), (pilot_ref.CNN_2D_SPD_FTR, ModelType.CNN_2D_SPD_FTR),
               (pilot_ref.CNN_2D_SPD_CTL, ModelType.CNN_2D_SPD_CTL), (pilot_ref.CNN_2D_FULL_HOUSE, ModelType.CNN_2D_FULL_HOUSE)]
SIZES = [(120, 160, 6), (96, 94, 4)]               # (h, w, frames)
CFGS = {"defaults": dict(), "break_smooth": dict(spd_ctl_break=True, smooth_steering_enabled=True, smooth_steering_threshold=0.35,
                                                 spd_ctl_threshold=0.9)}


def build_reference_model(mt, h, w):
    """keras_train.py:387-398, the reference's own calls."""
    shape = (h, w, 3)
    if mt == ModelType.CNN_2D:
        return ref_train.Keras_2D_CNN.get_model(input_shape=shape, num_outputs=2, num_feature_vectors=0)
    if mt == ModelType.CNN_2D_SPD_FTR:
        return ref_train.Keras_2D_CNN.get_model(input_shape=shape, num_outputs=2, num_feature_vectors=1)
    if mt == ModelType.CNN_2D_SPD_CTL:
        return ref_train.Keras_2D_CNN.get_model(input_shape=shape, num_outputs=2, num_feature_vectors=0)
    return ref_train.Keras_2D_FULL_HOUSE.get_model(input_shape=shape)


def main():
    arrays, meta = {}, {"cases": [], "cfgs": CFGS, "note": "weights: oracle.pilot_ref.random_weights(model_type, h, w, seed); frames: "
                                                          "synth.frame_pool(n, h, w, seed=frame_seed); speed / segment arrays stored"}
    for mi, mt in MODEL_TYPES:
        for h, w, n in SIZES:
            seed = 100 * mi + h
            wts = pilot_ref.random_weights(mi, h, w, seed=seed)
            model = build_reference_model(mt, h, w)
            layers = model.layers_by_name()
            trainable = {name for name, l in layers.items() if isinstance(l, (Conv2D, Dense))}
            assert {k.split("/")[0] for k in wts} == trainable, (sorted(trainable), sorted({k.split('/')[0] for k in wts}))
            for name in trainable:
                layers[name].weights = [wts[f"{name}/kernel"], wts[f"{name}/bias"]]
            frame_seed = 7 + seed
            frames = synth.frame_pool(n, h, w, seed=frame_seed)
            rng = np.random.default_rng(seed)
            speed = rng.uniform(0, 20, n)
            speed[::3] *= 0.02                                   # some cars nearly standing: predicted speed above the real one
            segment = rng.uniform(0, 10, n)
            for cname, over in CFGS.items():
                cfg = dict(REF_DEFAULTS)
                cfg.update(over)
                path = f"{mt.value}_{h}x{w}"
                REGISTRY[path] = model
                pilot = ref_pilot.KerasPilot(cfg, path, mt)
                Model.calls.clear()
                ctl = []
                raised = None
                for i in range(n):
                    try:
                        out = pilot.step(frames[i], float(speed[i]), float(segment[i]), 0.0, DriveMode.AI)      # cam/img, gym/speed, loc/segment, gym/cte, usr/mode
                    except ValueError as e:
                        # keras_pilot.py:61 / :73 hand the whole (steering, throttle) row to __cap, whose `if val < -1.0` (:143) is ambiguous for
                        # an array of two: the reference's cnn_2d / cnn_2d_speed_as_feature pilots raise here on every frame.  The model call
                        # before it still went through the reference's code (input tuple, reshapes), so its output is recorded.
                        raised = str(e)
                        out = (np.nan, np.nan, np.nan)
                    ctl.append([float(v) for v in out])
                assert len(Model.calls) == n
                assert (raised is not None) == (mt in (ModelType.CNN_2D, ModelType.CNN_2D_SPD_FTR)), (mt, raised)
                if raised:
                    meta.setdefault("reference_raises", {})[f"{mi}"] = raised
                key = f"{mi}/{h}x{w}/{cname}"
                arrays[f"model_out/{key}"] = np.concatenate([o for _, o in Model.calls], 0)
                arrays[f"ctl/{key}"] = np.asarray(ctl, np.float64)
                assert pilot.step(None, 1.0, 1.0, 0.0, DriveMode.AI) == (0.0, 0.0, 0.0)
                assert pilot.step(frames[0], 1.0, 1.0, 0.0, DriveMode.HUMAN) == (0.0, 0.0, 0.0)
            arrays[f"speed/{mi}/{h}x{w}"] = speed
            arrays[f"segment/{mi}/{h}x{w}"] = segment
            meta["cases"].append(dict(model_type=mi, h=h, w=w, n=n, weight_seed=seed, frame_seed=frame_seed,
                                      layers=sorted(trainable), inputs=[nd.layer.name for nd in model.inputs]))
    arrays["meta_json"] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    np.savez_compressed(os.path.join(HERE, "pilot.npz"), **arrays)
    print("pilot.npz:", len(arrays), "arrays;", [c["inputs"] for c in meta["cases"][::2]])


if __name__ == "__main__":
    main()
