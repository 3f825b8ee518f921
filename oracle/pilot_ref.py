"""CPU reference of the pilots' networks — TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's cpu_baseline leg).

Restates Keras_2D_CNN.get_model (TritonRacerSim/components/keras_train.py:127-174) and Keras_2D_FULL_HOUSE.get_model
(keras_train.py:184-245) as plain PyTorch fp32 functional calls on the CPU, fed the way KerasPilot.step feeds them
(components/keras_pilot.py:49-50,59,68-71,81,98-104).  Weights are a dict of numpy arrays in Keras layout under the reference's layer
names ("conv1/kernel" (kh,kw,in,out), "dense1/kernel" (in,out), "…/bias").

PARITY: TensorFlow is not installed in this image, so Keras' own float32 kernels are unpinned.  The GRAPH is pinned: tests/golden/pilot.npz holds
the outputs of the reference's own get_model / KerasPilot.step code, imported unmodified and run over a float64 numpy stand-in for the Keras
primitives (tests/golden/make_golden_pilot.py), and tests/test_pilot_host.py holds this restatement to them within 2e-5; its Conv2D / Dense
conventions are also checked against an explicit-loop numpy restatement.  A floating-point kernel is compared with it within a stated tolerance
(tests/test_pilot_gpu.py), not bit for bit.
"""
import numpy as np
import torch
import torch.nn.functional as F

CNN_2D, CNN_2D_SPD_FTR, CNN_2D_SPD_CTL, CNN_2D_FULL_HOUSE = 0, 1, 2, 3
CONVS = [(5, 2, 24), (5, 2, 32), (5, 2, 64), (3, 1, 64), (3, 1, 64), (3, 1, 128), (3, 1, 128)]      # keras_train.py:135-152


def conv_out_hw(h, w):
    for k, s, _ in CONVS:
        h, w = (h - k) // s + 1, (w - k) // s + 1
    return h, w


def weight_shapes(model_type, h=120, w=160):
    """name -> shape of every trainable tensor of the model (Keras layout)."""
    shapes, cin = {}, 3
    for i, (k, s, f) in enumerate(CONVS):
        shapes[f"conv{i + 1}/kernel"], shapes[f"conv{i + 1}/bias"] = (k, k, cin, f), (f,)
        cin = f
    ho, wo = conv_out_hw(h, w)
    flat = ho * wo * cin

    def dense(name, a, b):
        shapes[f"{name}/kernel"], shapes[f"{name}/bias"] = (a, b), (b,)

    if model_type == CNN_2D_FULL_HOUSE:
        # keras_train.py:215: x = [image features, feature branch]; :231: the steering head sees [x, speed branch]
        for names, d, o, extra in ((("feature1", "feature2", "feature3"), ("dense1", "dense2", "dense3"), "output_speed", 64),
                                   (("current_spd_1", "current_spd_2", "current_spd_3"), ("dense4", "dense5", "dense6"), "out_steering", 128)):
            dense(names[0], 1, 16), dense(names[1], 16, 32), dense(names[2], 32, 64)
            dense(d[0], flat + extra, 100), dense(d[1], 100, 50), dense(d[2], 50, 25), dense(o, 25, 1)
    else:
        nf = 1 if model_type == CNN_2D_SPD_FTR else 0
        if nf:
            dense("feature1", nf, 4 * nf), dense("feature2", 4 * nf, 8 * nf), dense("feature3", 8 * nf, 16 * nf)
        dense("dense1", flat + 16 * nf, 100), dense("dense2", 100, 50), dense("dense3", 50, 25), dense("output_layer", 25, 2)
    return shapes


def random_weights(model_type, h=120, w=160, seed=0, bias_scale=0.05):
    """He-uniform kernels (activations stay of order one through the ReLU stack, so an absolute tolerance on the outputs means
    something) and small random biases (a trained model has non-zero ones)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in weight_shapes(model_type, h, w).items():
        if name.endswith("/kernel"):
            fan_in = int(np.prod(shape[:-1]))
            lim = np.sqrt(6.0 / fan_in)
            out[name] = rng.uniform(-lim, lim, shape).astype(np.float32)
        else:
            out[name] = rng.uniform(-bias_scale, bias_scale, shape).astype(np.float32)
    return out


def _dense(wts, name, x, relu=True):
    y = x @ torch.from_numpy(wts[f"{name}/kernel"]) + torch.from_numpy(wts[f"{name}/bias"])
    return torch.relu(y) if relu else y


def conv_stack(wts, x_nhwc, first=0, last=7):
    """Layers first..last-1 on an NHWC float tensor; returns the list of NHWC activations."""
    acts = []
    x = x_nhwc.permute(0, 3, 1, 2)
    for i in range(first, last):
        k, s, _ = CONVS[i]
        kern = torch.from_numpy(wts[f"conv{i + 1}/kernel"]).permute(3, 2, 0, 1).contiguous()       # HWIO -> OIHW
        x = torch.relu(F.conv2d(x, kern, torch.from_numpy(wts[f"conv{i + 1}/bias"]), stride=s))
        acts.append(x.permute(0, 2, 3, 1).contiguous())
    return acts


def heads(wts, model_type, flat, spd_feature=None, loc_feature=None):
    """The Dense part on the flattened conv7 output (N, flat)."""
    if model_type == CNN_2D_FULL_HOUSE:
        y = torch.from_numpy(np.asarray(loc_feature, np.float32)).reshape(-1, 1)               # feature_vec_input (keras_train.py:213)
        for n in ("feature1", "feature2", "feature3"):
            y = _dense(wts, n, y)
        x = torch.cat([flat, y], 1)
        z = x
        for n in ("dense1", "dense2", "dense3"):
            z = _dense(wts, n, z)
        speed = _dense(wts, "output_speed", z, relu=False)
        s = torch.from_numpy(np.asarray(spd_feature, np.float32)).reshape(-1, 1)               # current_spd_input (226)
        for n in ("current_spd_1", "current_spd_2", "current_spd_3"):
            s = _dense(wts, n, s)
        s = torch.cat([x, s], 1)                # keras_train.py:231 concatenates x (image features ++ feature branch) with s
        for n in ("dense4", "dense5", "dense6"):
            s = _dense(wts, n, s)
        steering = _dense(wts, "out_steering", s, relu=False)
        return torch.cat([steering, speed], 1)
    z = flat
    if model_type == CNN_2D_SPD_FTR:
        y = torch.from_numpy(np.asarray(spd_feature, np.float32)).reshape(-1, 1)
        for n in ("feature1", "feature2", "feature3"):
            y = _dense(wts, n, y)
        z = torch.cat([flat, y], 1)
    for n in ("dense1", "dense2", "dense3"):
        z = _dense(wts, n, z)
    return _dense(wts, "output_layer", z, relu=False)


def forward(wts, model_type, frames_u8, spd_feature=None, loc_feature=None, return_acts=False):
    """frames (N,H,W,3) u8 -> (N,2) float32, as KerasPilot.step calls the model (keras_pilot.py:49-50)."""
    with torch.no_grad():
        x = torch.from_numpy(np.asarray(frames_u8, dtype=np.float32) / np.float32(255))
        acts = conv_stack(wts, x)
        out = heads(wts, model_type, acts[-1].reshape(x.shape[0], -1), spd_feature, loc_feature)
    return (out.numpy(), [a.numpy() for a in acts]) if return_acts else out.numpy()
