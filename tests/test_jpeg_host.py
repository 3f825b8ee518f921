"""CPU-side check of the tub-ingestion arithmetic: csrc/jpeg_core.cuh + csrc/jpeg_host.h (pure host/device functions, the same source the
kernels compile) built with g++ and compared with Pillow — the decoder the reference's loaders call (keras_train.py:41) — bit for bit,
plus the committed fixture (tests/golden/jpeg.npz: files and Pillow's decode from the build container)."""
import ctypes as C
import io
import os
import subprocess

import numpy as np
import pytest
from PIL import Image

from tests.conftest import ROOT
from triton_racer_sim_b200 import synth


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("jpg") / "libjpgcheck.so"
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", str(out), os.path.join(ROOT, "tests", "host_jpeg_check.cpp")])
    lib = C.CDLL(str(out))
    lib.jpg_decode_rgb_host.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
    return lib


def host_decode(lib, data, h, w):
    arr = np.frombuffer(data, np.uint8).copy()
    out = np.zeros((h, w, 3), np.uint8)
    rc = lib.jpg_decode_rgb_host(arr.ctypes.data, len(data), h, w, out.ctypes.data)
    return rc, out


@pytest.mark.parametrize("h,w", [(120, 160), (240, 320), (121, 163), (7, 9), (17, 33), (2, 3), (1, 1), (5, 2)])
def test_host_decoder_matches_pillow(host_lib, h, w):
    frames = synth.frame_pool(4, h, w, seed=h * 3 + w)
    for q in (75, 30, 95, 100):
        for f in frames:
            buf = io.BytesIO()
            Image.fromarray(f).save(buf, format="JPEG", quality=q)
            data = buf.getvalue()
            rc, got = host_decode(host_lib, data, h, w)
            assert rc == 0
            assert np.array_equal(got, np.asarray(Image.open(io.BytesIO(data)))), (h, w, q)


@pytest.mark.parametrize("sub", [0, 1, 2])
def test_host_decoder_other_chroma_subsamplings(host_lib, sub):
    """4:4:4 (no upsampling), 4:2:2 (h2v1 fancy upsampling) and 4:2:0 on odd sizes."""
    for h, w in ((120, 160), (121, 163), (7, 9), (3, 5), (17, 4)):
        for f in synth.frame_pool(3, h, w, seed=31 * h + w):
            buf = io.BytesIO()
            Image.fromarray(f).save(buf, format="JPEG", quality=85, subsampling=sub)
            data = buf.getvalue()
            rc, got = host_decode(host_lib, data, h, w)
            assert rc == 0 and np.array_equal(got, np.asarray(Image.open(io.BytesIO(data)))), (h, w, sub)


def test_host_decoder_matches_committed_fixture(host_lib):
    g = np.load(os.path.join(ROOT, "tests", "golden", "jpeg.npz"))
    blob, offsets = g["blob"], g["offsets"]
    for k in range(len(offsets) - 1):
        rc, got = host_decode(host_lib, blob[int(offsets[k]):int(offsets[k + 1])].tobytes(), 120, 160)
        assert rc == 0 and np.array_equal(got, g["decoded"][k])


def test_parser_rejects_what_the_kernels_do_not_decode(host_lib):
    f = synth.frame_pool(1, 32, 32, seed=1)[0]
    for kw in (dict(progressive=True),):
        buf = io.BytesIO()
        Image.fromarray(f).save(buf, format="JPEG", **kw)
        rc, _ = host_decode(host_lib, buf.getvalue(), 32, 32)
        assert rc == 111, kw                       # 100 + JPG_E_UNSUPPORTED
    buf = io.BytesIO()
    Image.fromarray(f[..., 0]).save(buf, format="JPEG")          # greyscale
    assert host_decode(host_lib, buf.getvalue(), 32, 32)[0] == 111
    assert host_decode(host_lib, b"not a jpeg at all", 32, 32)[0] == 110


def test_host_decoder_survives_corrupted_files_under_the_address_sanitizer(tmp_path):
    """Telemetry JPEGs are external input: 3,000 corrupted files (random bytes, truncation, runs of 0xff, damaged tables, inserted bytes; all three
    chroma subsamplings) through the product's parser / entropy decoder / IDCT built with -fsanitize=address.  A malformed file may give an error
    code or garbage pixels, never an access outside its buffers."""
    exe = tmp_path / "jpeg_fuzz"
    build = subprocess.run(["g++", "-O1", "-g", "-fsanitize=address", "-fno-omit-frame-pointer", "-o", str(exe),
                            os.path.join(ROOT, "tests", "host_jpeg_fuzz.cpp")], capture_output=True, text=True)
    if build.returncode != 0:
        pytest.skip("no AddressSanitizer runtime in this toolchain: " + build.stderr[-200:])
    files = []
    for i, (q, sub) in enumerate(((75, 2), (90, 1), (60, 0))):
        buf = io.BytesIO()
        Image.fromarray(synth.frame_pool(1, 120, 160, seed=50 + i)[0]).save(buf, format="JPEG", quality=q, subsampling=sub)
        p = tmp_path / f"f{i}.jpg"
        p.write_bytes(buf.getvalue())
        files.append(str(p))
    run = subprocess.run([str(exe), "120", "160", "1000"] + files, capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stderr[-3000:]
    assert run.stdout.startswith("runs 3000 decoded")
