"""The pilots' networks on the tensor cores against the fp32 CPU restatement of the reference's models
(oracle/pilot_ref.py <- keras_train.py:127-245, keras_pilot.py:49-117).

A floating-point kernel, so a tolerance instead of bit equality: operands are fp16 (10-bit mantissa, what TensorFlow's TF32 convolutions
keep on a GPU), accumulation fp32.  Two levels:
  * layer by layer, each layer's fp32 reference fed with the GPU's own fp16 input of that layer: only this layer's rounding remains
    -> |err| <= 2e-3 * max|activation| (LAYER_TOL);
  * end to end against the all-fp32 network -> |err| <= 1e-2 on outputs of order 0.1..1 (E2E_TOL).
"""
import numpy as np
import pytest
import torch

from oracle import pilot_ref as ref
from triton_racer_sim_b200 import synth
from triton_racer_sim_b200.pilot import KerasPilot, ModelType, PilotNet

pytestmark = pytest.mark.gpu

LAYER_TOL = 2e-3
E2E_TOL = 1e-2
KINDS = {ModelType.CNN_2D: ref.CNN_2D, ModelType.CNN_2D_SPD_FTR: ref.CNN_2D_SPD_FTR, ModelType.CNN_2D_SPD_CTL: ref.CNN_2D_SPD_CTL,
         ModelType.CNN_2D_FULL_HOUSE: ref.CNN_2D_FULL_HOUSE}


def features(n, seed):
    rng = np.random.default_rng(seed)
    return (rng.uniform(0, 1, n).astype(np.float32), rng.uniform(0, 10, n).astype(np.float32))


@pytest.mark.parametrize("rowconv", ["1", "0"])
@pytest.mark.parametrize("n,h,w,cap", [(9, 120, 160, 16), (7, 240, 320, 4), (5, 96, 94, 8)])
def test_layer_by_layer(monkeypatch, rowconv, n, h, w, cap):
    """rowconv: conv2 / conv3 as GEMMs per input row (k_pilot_rowconv, the default) or per (output row, kernel row) (k_pilot_gemm)."""
    monkeypatch.setenv("TRS_PILOT_ROWCONV", rowconv)
    frames = synth.frame_pool(n, h, w, seed=11)
    wts = ref.random_weights(ref.CNN_2D_FULL_HOUSE, h, w, seed=3)
    spd, loc = features(n, 1)
    net = PilotNet(ModelType.CNN_2D_FULL_HOUSE, wts, h, w, device=0, max_batch=cap)
    out = net.forward_device(torch.from_numpy(frames).cuda(), torch.from_numpy(spd).cuda(), torch.from_numpy(loc).cuda()).cpu().numpy()
    if n > cap:                                               # the activations of the last chunk are what the workspace holds
        frames, spd, loc, out = frames[-(n % cap or cap):], spd[-(n % cap or cap):], loc[-(n % cap or cap):], out[-(n % cap or cap):]
        n = frames.shape[0]
    prev = frames.astype(np.float32) / np.float32(255)       # conv1 reads the u8 frame itself (1 / 255 folded into its weights)
    report = []
    for layer in range(1, 8):
        got = net.activation(layer, n).astype(np.float32)
        want = ref.conv_stack(wts, torch.from_numpy(prev), layer - 1, layer)[0].numpy()
        assert got.shape == want.shape, (layer, got.shape, want.shape)
        err = np.abs(got - want).max()
        report.append((layer, float(err), float(np.abs(want).max())))
        assert err <= LAYER_TOL * max(1.0, np.abs(want).max()), f"conv{layer}: max|err| {err} (max|act| {np.abs(want).max()}); so far {report}"
        prev = got
    flat = torch.from_numpy(prev.reshape(n, -1))
    want_out = ref.heads(wts, ref.CNN_2D_FULL_HOUSE, flat, spd, loc).numpy()
    assert np.abs(out - want_out).max() <= 2e-3, f"heads: {np.abs(out - want_out).max()}"
    e2e = ref.forward(wts, ref.CNN_2D_FULL_HOUSE, frames, spd, loc)
    assert np.abs(out - e2e).max() <= E2E_TOL, f"end to end: {np.abs(out - e2e).max()}"
    net.close()


@pytest.mark.parametrize("mt", list(KINDS))
@pytest.mark.parametrize("n,cap", [(1, 8), (37, 64), (130, 48)])
def test_models_end_to_end(mt, n, cap):
    h, w = 120, 160
    frames = synth.frame_pool(n, h, w, seed=5 + n)
    wts = ref.random_weights(KINDS[mt], h, w, seed=n)
    spd, loc = features(n, n)
    net = PilotNet(mt, wts, h, w, device=0, max_batch=cap)
    out = net.forward_device(torch.from_numpy(frames).cuda(), torch.from_numpy(spd).cuda(), torch.from_numpy(loc).cuda()).cpu().numpy()
    want = ref.forward(wts, KINDS[mt], frames, spd, loc)
    assert out.shape == want.shape == (n, 2)
    assert np.abs(out - want).max() <= E2E_TOL, np.abs(out - want).max()
    net.close()


def test_frames_that_are_not_16_byte_aligned_take_the_per_thread_loads():
    """k_pilot_conv1r reads its patches through a TMA tensor map (16-byte aligned base); a view that starts 4 bytes into an allocation
    must fall back to k_pilot_conv1 and give the same outputs."""
    n, h, w = 6, 120, 160
    frames = synth.frame_pool(n, h, w, seed=21)
    wts = ref.random_weights(ref.CNN_2D, h, w, seed=4)
    net = PilotNet(ModelType.CNN_2D, wts, h, w, device=0, max_batch=8)
    aligned = torch.from_numpy(frames).cuda()
    raw = torch.empty(frames.size + 64, dtype=torch.uint8, device='cuda')
    shifted = raw[4:4 + frames.size].view(n, h, w, 3)
    shifted.copy_(aligned)
    assert shifted.data_ptr() % 16 == 4
    a = net.forward_device(aligned).cpu().numpy()
    b = net.forward_device(shifted).cpu().numpy()
    want = ref.forward(wts, ref.CNN_2D, frames)
    assert np.abs(a - want).max() <= E2E_TOL and np.abs(b - want).max() <= E2E_TOL
    assert np.abs(a - b).max() <= 2e-3, np.abs(a - b).max()
    net.close()


@pytest.mark.parametrize("h,w", [(240, 320), (122, 166), (96, 94)])
def test_other_frame_sizes(h, w):
    n = 5
    frames = synth.frame_pool(n, h, w, seed=h)
    wts = ref.random_weights(ref.CNN_2D, h, w, seed=h + w)
    net = PilotNet(ModelType.CNN_2D, wts, h, w, device=0, max_batch=4)
    out = net.forward_device(torch.from_numpy(frames).cuda()).cpu().numpy()
    want = ref.forward(wts, ref.CNN_2D, frames)
    assert np.abs(out - want).max() <= E2E_TOL, np.abs(out - want).max()
    net.close()


def test_keras_pilot_component_matches_the_reference_glue():
    """KerasPilot.step: model -> cap / speed control exactly as keras_pilot.py:56-117, given the model outputs."""
    import oracle
    n, h, w = 64, 120, 160
    frames = synth.frame_pool(n, h, w, seed=2)
    rng = np.random.default_rng(0)
    speed = rng.uniform(0, 20, n)
    seg = rng.uniform(0, 10, n)
    for mt in KINDS:
        wts = ref.random_weights(KINDS[mt], h, w, seed=1)
        cfg = dict(spd_ctl_break=True)
        pilot = KerasPilot(cfg, wts, mt, device=0, max_batch=64)
        assert pilot.step_inputs == ['cam/img', 'gym/speed', 'loc/segment', 'gym/cte', 'usr/mode']
        assert pilot.step_outputs == ['ai/steering', 'ai/throttle', 'ai/breaking']
        assert pilot.step(None, 0, 0, 0, 'ai') == (0.0, 0.0, 0.0)
        assert pilot.step(frames[0], 1.0, 1.0, 0.0, 'user') == (0.0, 0.0, 0.0)
        fr = torch.from_numpy(frames).cuda()
        s, t, b = pilot.step(fr, torch.from_numpy(speed).cuda(), torch.from_numpy(seg).cuda(), None, 'ai')
        spd_feat = (speed / 20).astype(np.float32)
        raw = pilot.model.forward_device(fr, torch.from_numpy(spd_feat).cuda(), torch.from_numpy(seg.astype(np.float32)).cuda()).cpu().numpy()
        if mt in (ModelType.CNN_2D, ModelType.CNN_2D_SPD_FTR):
            assert np.array_equal(s.cpu().numpy(), np.clip(raw[:, 0].astype(np.float64), -1, 1))
            assert np.array_equal(t.cpu().numpy(), np.clip(raw[:, 1].astype(np.float64), -1, 1))
            assert not b.cpu().numpy().any()
        else:
            so, th, br, _ = oracle.speed_control(speed, raw[:, 1], raw[:, 0], pilot.cfg)
            assert np.array_equal(s.cpu().numpy(), so)
            assert np.allclose(t.cpu().numpy(), th, rtol=1e-5, atol=0) and np.allclose(b.cpu().numpy(), br, rtol=1e-5, atol=0)
        one = pilot.step(frames[3], float(speed[3]), float(seg[3]), 0.0, 'ai')
        assert all(isinstance(v, float) for v in one)
        assert abs(one[0] - float(s[3])) <= 1e-6
        pilot.onShutdown()


def test_components_chain_like_the_readme(tmp_path):
    """ImgPreprocessing -> LocationTracker -> KerasPilot with CUDA tensors, weights from an .npz file, against the oracles end to end."""
    import oracle
    from triton_racer_sim_b200 import ImgPreprocessing, LocationTracker
    from triton_racer_sim_b200.config import full_house_config
    n, h, w = 48, 120, 160
    cfg = full_house_config()
    cfg['spd_ctl_break'] = True
    frames = synth.frame_pool(n, h, w, seed=21)
    wp = synth.synthetic_track(300)
    xyz, cur, _, _ = synth.car_states(wp, n, seed=3)
    wts = ref.random_weights(ref.CNN_2D_FULL_HOUSE, h, w, seed=6)
    np.savez(tmp_path / "model.npz", **wts)
    pre, trk = ImgPreprocessing(cfg, device=0), LocationTracker(wp, device=0)
    pilot = KerasPilot(cfg, str(tmp_path / "model.npz"), 'cnn_2d_full_house', device=0)
    pilot.step_inputs[0] = 'cam/processed_img'                                       # manage.py:49-50
    processed, = pre.step(torch.from_numpy(frames).cuda())
    t_xyz = torch.from_numpy(xyz).cuda()
    segment, = trk.step(t_xyz[:, 0], t_xyz[:, 1], t_xyz[:, 2])
    s, t, b = pilot.step(processed, torch.from_numpy(cur).cuda(), segment, None, 'ai')
    # the same through the CPU checkers
    want_img = oracle.process_batch(frames, cfg)
    assert np.array_equal(processed.cpu().numpy(), want_img)
    _, want_seg = oracle.locate(wp, xyz)
    assert np.array_equal(segment.cpu().numpy(), want_seg)
    model = ref.forward(wts, ref.CNN_2D_FULL_HOUSE, want_img, (cur / 20).astype(np.float32), want_seg.astype(np.float32))
    got_model = pilot.model.forward_device(processed, torch.from_numpy((cur / 20).astype(np.float32)).cuda(), segment).cpu().numpy()
    assert np.abs(got_model - model).max() <= E2E_TOL
    so, th, br, _ = oracle.speed_control(cur, got_model[:, 1], got_model[:, 0], pilot.cfg)
    assert np.array_equal(s.cpu().numpy(), so) and np.allclose(t.cpu().numpy(), th, rtol=1e-5, atol=0)
    assert np.allclose(b.cpu().numpy(), br, rtol=1e-5, atol=0)
    for c in (pre, trk, pilot):
        c.onShutdown()


def test_outputs_do_not_depend_on_batch_or_shard():
    """Size-independent property at a larger size: a frame's outputs are bit-identical whatever batch, chunk or shard it sits in
    (rows of an MMA tile are independent; tiles stack frames in conv3..conv7 and in the Dense GEMM)."""
    h, w, n = 120, 160, 5000
    pool = torch.from_numpy(synth.frame_pool(64, h, w, seed=31)).cuda()
    frames = synth.expand_torch(pool, n)
    wts = ref.random_weights(ref.CNN_2D_FULL_HOUSE, h, w, seed=13)
    g = torch.Generator(device='cuda').manual_seed(5)
    spd = torch.rand(n, device='cuda', generator=g)
    loc = torch.rand(n, device='cuda', generator=g) * 10
    big = PilotNet(ModelType.CNN_2D_FULL_HOUSE, wts, h, w, device=0, max_batch=2048)           # three chunks, the last one ragged
    whole = big.forward_device(frames, spd, loc).clone()
    small = PilotNet(ModelType.CNN_2D_FULL_HOUSE, wts, h, w, device=0, max_batch=64)
    for lo, hi in ((0, 1), (7, 20), (2040, 2060), (4990, 5000), (1234, 1234 + 333)):             # shards, incl. across chunk borders
        part = small.forward_device(frames[lo:hi], spd[lo:hi], loc[lo:hi])
        assert torch.equal(part, whole[lo:hi]), (lo, hi)
    # and against the fp32 reference on a sample
    idx = [0, 63, 2047, 2048, 4999]
    want = ref.forward(wts, ref.CNN_2D_FULL_HOUSE, frames[idx].cpu().numpy(), spd[idx].cpu().numpy(), loc[idx].cpu().numpy())
    assert np.abs(whole[idx].cpu().numpy() - want).max() <= E2E_TOL
    big.close()
    small.close()


def test_workspace_grows_with_the_batch():
    h, w = 120, 160
    wts = ref.random_weights(ref.CNN_2D, h, w, seed=12)
    net = PilotNet(ModelType.CNN_2D, wts, h, w, device=0, max_batch=64, initial_batch=2)
    assert net.capacity == 2
    expect = 2
    for n in (1, 3, 40, 5, 100):
        frames = synth.frame_pool(n, h, w, seed=n)
        out = net.forward_device(torch.from_numpy(frames).cuda()).cpu().numpy()
        expect = max(expect, min(n, 64))                                  # grows to the largest batch seen, never past max_batch
        assert net.capacity == expect
        assert np.abs(out - ref.forward(wts, ref.CNN_2D, frames)).max() <= E2E_TOL
    net.close()


def test_cap_and_smooth_steering_like_the_reference():
    """keras_pilot.py:142-153 on the model outputs of ModelType.CNN_2D: cap to [-1, 1], snap beyond the threshold, breaking 0."""
    n, h, w = 96, 120, 160
    frames = torch.from_numpy(synth.frame_pool(n, h, w, seed=4)).cuda()
    wts = ref.random_weights(ref.CNN_2D, h, w, seed=8)
    wts["output_layer/kernel"] = wts["output_layer/kernel"] * 6          # push outputs beyond +-1 so the cap matters
    pilot = KerasPilot(dict(smooth_steering_enabled=True, smooth_steering_threshold=0.3), wts, ModelType.CNN_2D, device=0, max_batch=128)
    s, t, b = (v.cpu().numpy() for v in pilot.step(frames, None, None, None, 'ai_steering'))
    raw = pilot.model.forward_device(frames).cpu().numpy().astype(np.float64)
    cap = np.clip(raw, -1.0, 1.0)
    want_s = np.where(cap[:, 0] > 0.3, 1.0, np.where(cap[:, 0] < -0.3, -1.0, cap[:, 0]))
    assert (np.abs(raw) > 1).any() and (np.abs(cap[:, 0]) <= 0.3).any()
    assert np.array_equal(s, want_s) and np.array_equal(t, cap[:, 1]) and not b.any()
    pilot.onShutdown()


def test_missing_or_misshapen_weights_are_refused():
    wts = ref.random_weights(ref.CNN_2D, 120, 160, seed=0)
    bad = dict(wts)
    del bad["conv3/bias"]
    with pytest.raises(ValueError, match="conv3/bias"):
        PilotNet(ModelType.CNN_2D, bad, 120, 160, device=0, max_batch=4)
    bad = dict(wts)
    bad["dense1/kernel"] = bad["dense1/kernel"][:-1]
    with pytest.raises(ValueError, match="dense1/kernel"):
        PilotNet(ModelType.CNN_2D, bad, 120, 160, device=0, max_batch=4)
    with pytest.raises(ValueError):
        PilotNet(ModelType.CNN_2D, wts, 120, 161, device=0, max_batch=4)
