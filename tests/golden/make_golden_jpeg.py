#!/usr/bin/env python
"""Generator of tests/golden/jpeg.npz: tub images as the reference's recorder writes them, and what the reference's loaders read back.

  writer  components/datastorage.py:78   Image.fromarray(img).save(path)            (Pillow defaults: baseline, quality 75, 4:2:0)
  reader  components/keras_train.py:41   np.asarray(Image.open(path))               (Pillow -> libjpeg-turbo default decompression)

Six 120x160 frames of the seeded synthetic pool (one noise frame, smooth frames, track-like scenes) are encoded with the writer's call and
decoded with the reader's call in THIS container (Pillow / libjpeg-turbo versions printed below); the files (`blob`, `offsets`) and the
decoded pixels (`decoded`) are stored.  tests/test_jpeg_host.py and tests/test_jpeg_gpu.py decode the stored files with the CUDA
decoder's host twin and with the kernels and require the stored pixels bit for bit, so the fixture travels to the GPU box without Pillow
having to produce the same bytes there.

Run from the repo root:  python tests/golden/make_golden_jpeg.py
"""
import io
import os
import sys

import numpy as np
from PIL import Image, features

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from triton_racer_sim_b200 import synth  # noqa: E402


def main():
    pool = synth.frame_pool(8, 120, 160, seed=4242)[[0, 1, 2, 3, 5, 7]]       # noise, smooth, smooth, track, smooth, track
    files, decoded = [], []
    for img in pool:
        buf = io.BytesIO()
        Image.fromarray(img).save(buf, format="JPEG")                        # datastorage.py:78 (the file name ends in .jpg there)
        data = buf.getvalue()
        files.append(data)
        decoded.append(np.asarray(Image.open(io.BytesIO(data))))             # keras_train.py:41 (before the float32 cast)
    sizes = np.fromiter((len(f) for f in files), dtype=np.uint64, count=len(files))
    offsets = np.zeros(len(files) + 1, np.uint64)
    np.cumsum(sizes, out=offsets[1:])
    np.savez_compressed(os.path.join(HERE, "jpeg.npz"), blob=np.frombuffer(b"".join(files), np.uint8), offsets=offsets,
                        decoded=np.stack(decoded).astype(np.uint8))
    print("jpeg.npz:", [len(f) for f in files], "Pillow", Image.__version__, "libjpeg-turbo", features.version("libjpeg_turbo"))


if __name__ == "__main__":
    main()
