"""Tub ingestion: the CUDA JPEG decoder against Pillow (the decoder the reference's loaders call, keras_train.py:41), bit for bit."""
import io
import json
import os

import numpy as np
import pytest
import torch
from PIL import Image

from triton_racer_sim_b200 import synth, tub

pytestmark = pytest.mark.gpu


def encode(frames, **kw):
    files = []
    for f in frames:
        buf = io.BytesIO()
        Image.fromarray(f).save(buf, format="JPEG", **kw)            # datastorage.py:78 (defaults: quality 75, 4:2:0)
        files.append(buf.getvalue())
    return files


def pil_decode(files):
    return np.stack([np.asarray(Image.open(io.BytesIO(b))) for b in files])


@pytest.mark.parametrize("h,w,n,kw", [(120, 160, 300, {}), (240, 320, 40, {}), (120, 160, 60, dict(quality=95)), (120, 160, 60, dict(quality=20)),
                                     (121, 163, 9, {}), (7, 9, 5, {}), (64, 96, 7, dict(quality=100)), (2, 3, 4, {}), (1, 1, 3, {})])
def test_decoder_matches_pillow(h, w, n, kw):
    frames = synth.frame_pool(n, h, w, seed=7 * h + w)
    files = encode(frames, **kw)
    got = tub.decode_jpeg_batch(files, device=0)
    assert got.shape == (n, h, w, 3) and got.dtype == torch.uint8
    want = pil_decode(files)
    assert np.array_equal(got.cpu().numpy(), want)


@pytest.mark.parametrize("sub,h,w", [(0, 120, 160), (1, 120, 160), (0, 121, 163), (1, 121, 163), (1, 7, 9), (0, 3, 5), (2, 33, 70)])
def test_other_chroma_subsamplings(sub, h, w):
    frames = synth.frame_pool(20, h, w, seed=sub + h + w)
    files = encode(frames, quality=85, subsampling=sub)                # 0: 4:4:4, 1: 4:2:2, 2: 4:2:0
    assert np.array_equal(tub.decode_jpeg_batch(files, device=0).cpu().numpy(), pil_decode(files))


def test_mixed_table_sets_and_golden_fixture():
    frames = synth.frame_pool(30, 120, 160, seed=3)
    files = [encode(frames[k:k + 1], quality=q)[0] for k, q in enumerate([75, 50, 90] * 10)]      # three table sets in one batch
    assert np.array_equal(tub.decode_jpeg_batch(files, device=0).cpu().numpy(), pil_decode(files))
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jpeg.npz"))
    blob, offsets = g["blob"], g["offsets"]
    got = tub.decode_jpeg_batch((blob, offsets), hw=(120, 160), device=0)
    assert np.array_equal(got.cpu().numpy(), g["decoded"])                                         # Pillow 12.2 / libjpeg-turbo 3.x in the build container


def test_unsupported_and_corrupt_files_raise():
    frames = synth.frame_pool(2, 120, 160, seed=5)
    with pytest.raises(ValueError):
        tub.decode_jpeg_batch(encode(frames, progressive=True), device=0)
    with pytest.raises(ValueError):
        tub.decode_jpeg_batch(encode(frames[:1], subsampling=0) + encode(frames[1:], subsampling=2), device=0)   # two samplings in one batch
    grey = io.BytesIO()
    Image.fromarray(frames[0][..., 0]).save(grey, format="JPEG")
    with pytest.raises(ValueError):
        tub.decode_jpeg_batch([grey.getvalue()], hw=(120, 160), device=0)
    ok = encode(frames)
    with pytest.raises(ValueError):
        tub.decode_jpeg_batch([ok[0], ok[1][:200]], device=0)                                       # truncated header
    with pytest.raises(ValueError):
        tub.decode_jpeg_batch(ok, hw=(240, 320), device=0)                                          # wrong stated size


def test_large_batch_of_mixed_records():
    """20,003 records of very different sizes and two table sets in random order: every copy of a file decodes to Pillow's pixels."""
    uniq = synth.frame_pool(96, 120, 160, seed=23)
    files = encode(uniq[:48]) + encode(uniq[48:], quality=92)                 # two table sets, sizes from ~3 KB to ~25 KB
    want_u = pil_decode(files)
    n = 20003
    order = np.random.default_rng(2).integers(0, len(files), n)
    got = tub.decode_jpeg_batch([files[i] for i in order], device=0).cpu().numpy()
    assert got.shape == (n, 120, 160, 3)
    for k in range(len(files)):
        sel = np.nonzero(order == k)[0]
        assert (got[sel] == want_u[k][None]).all(), k


def test_tub_reader_round_trip(tmp_path):
    """A tub written the way the reference's recorder writes it (datastorage.py:67-79), read back through the GPU decoder and through
    the full chain: the same as the reference loader's `Image.open` followed by the oracle chain."""
    import oracle
    from triton_racer_sim_b200 import ImgPreprocessing
    from triton_racer_sim_b200.config import full_house_config
    frames = synth.frame_pool(12, 120, 160, seed=11)
    for i, f in enumerate(frames, start=1):
        Image.fromarray(f).save(os.path.join(tmp_path, f"img_{i}.jpg"))
        with open(os.path.join(tmp_path, f"record_{i}.json"), "w") as fp:
            json.dump({"cam/img": f"img_{i}.jpg", "mux/steering": 0.1 * i, "gym/speed": 2.0 * i}, fp)
    rd = tub.TubReader(str(tmp_path), device=0)
    assert rd.count() == 12
    dev_frames, records = rd.load(range(1, 13))
    want = np.stack([np.asarray(Image.open(os.path.join(tmp_path, f"img_{i}.jpg"))) for i in range(1, 13)])
    assert np.array_equal(dev_frames.cpu().numpy(), want)
    assert [r["gym/speed"] for r in records] == [2.0 * i for i in range(1, 13)]
    cfg = full_house_config()
    comp = ImgPreprocessing(cfg, device=0)
    u8, _ = comp.process_device(dev_frames)
    assert np.array_equal(u8.cpu().numpy(), oracle.process_batch(want, cfg))
    comp.onShutdown()
    # the trainer's view of the same batch (keras_train.py:33-57 with SpeedCtlDataLoader, :271-276): labels (steering, speed / 20)
    fr2, labels, feats = rd.load_examples(range(1, 13), "cnn_2d_speed_control")
    assert torch.equal(fr2, dev_frames) and feats is None
    assert np.array_equal(labels.cpu().numpy(), np.float32([[0.1 * i, 2.0 * i / 20] for i in range(1, 13)]))
    rd.close()


def test_telemetry_packets_like_the_gym_interface():
    """gyminterface.py:95-104: Image.open(BytesIO(base64.b64decode(packet["image"]))) and five floats per packet."""
    import base64
    frames = synth.frame_pool(150, 120, 160, seed=17)
    files = encode(frames)
    rng = np.random.default_rng(3)
    packets = []
    for k, f in enumerate(files):
        txt = base64.b64encode(f).decode()
        if k % 7 == 0:
            txt = "\n".join(txt[i:i + 76] for i in range(0, len(txt), 76))            # MIME-style line breaks are ignored
        if k % 5 == 0:
            txt = txt.rstrip("=")                                                       # padding is optional
        packets.append({"msg_type": "telemetry", "image": txt, "pos_x": str(rng.normal()), "pos_y": float(rng.normal()), "pos_z": 1.5 + k,
                        "speed": rng.uniform(0, 20), "cte": "0.25"})
    got = tub.decode_telemetry_batch(packets, device=0)
    want = np.stack([np.asarray(Image.open(io.BytesIO(f))) for f in files])           # == b64decode of the (re-padded) strings
    assert np.array_equal(got["cam/img"].cpu().numpy(), want)
    assert np.array_equal(got["gym/x"].cpu().numpy(), np.asarray([float(p["pos_x"]) for p in packets]))
    assert np.array_equal(got["gym/cte"].cpu().numpy(), np.full(150, 0.25)) and got["gym/speed"].dtype == torch.float64
    packets[3]["image"] = packets[3]["image"][:50] + "!" + packets[3]["image"][51:]
    with pytest.raises(ValueError):
        tub.decode_telemetry_batch(packets, hw=(120, 160), device=0)


def test_telemetry_fixture_from_the_reference_run():
    """tests/golden/telemetry.npz: what the reference's own GymInterface.on_msg_recv / step publish for six simulator packets
    (tests/golden/make_golden_telemetry.py runs that code unmodified with a socket-less SDClient): image and the five floats, bit for bit."""
    import json
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "telemetry.npz"))
    packets = json.loads(bytes(g["packets_json"]).decode())
    got = tub.decode_telemetry_batch(packets, device=0)
    assert np.array_equal(got["cam/img"].cpu().numpy(), g["images"])
    for j, key in enumerate(("gym/x", "gym/y", "gym/z", "gym/speed", "gym/cte")):
        assert np.array_equal(got[key].cpu().numpy(), g["floats"][:, j]), key
