#!/usr/bin/env python
"""Generate tests/golden/tub.npz by running the reference's own tub loaders (TritonRacerSim/components/keras_train.py:22-119, 264-325:
DataLoader, SpeedFeatureDataLoader, SpeedCtlDataLoader, FullHouseDataLoader, imported UNMODIFIED from /root/reference) over a small tub
written by the reference's own recorder (components/datastorage.py, DataStorage.step + its file thread, also imported unmodified).

TensorFlow is not installed in this image; the loaders only touch it after the per-record loop (`tf.data.Dataset.from_tensors(...)`), so the
stand-in of make_golden_pilot.py plus a do-nothing `tf.data.Dataset` is enough to run `load()`; what is stored is `loader.dataset`, the list
of (image float32 / 255, feature vector, labels) the reference builds record by record.  A second tub with record 4 missing pins where the
reference stops counting (the first missing file ends the folder, keras_train.py:54-56).
Run in the build container: ``python tests/golden/make_golden_tub.py``.
"""
import io
import json
import os
import sys
import shutil
import tempfile
import time

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_pilot as stand_in  # noqa: E402  (installs the `tensorflow` stand-in and puts /root/reference and the repo on sys.path)

import tensorflow as tf  # noqa: E402  (the stand-in)


class _Dataset:
    @staticmethod
    def from_tensors(t):
        return _Dataset()

    def unbatch(self):
        return self

    def shuffle(self, n):
        return self

    def batch(self, n, drop_remainder=False):
        return self


tf.data = type(sys)("tensorflow.data")
tf.data.Dataset = _Dataset

from TritonRacerSim.components import keras_train as ref_train  # noqa: E402

from triton_racer_sim_b200 import synth  # noqa: E402

LOADER_CLASSES = ["DataLoader", "SpeedFeatureDataLoader", "SpeedCtlDataLoader", "FullHouseDataLoader"]


def write_tub(folder, frames, values):
    """The tub as the reference's recorder writes it: DataStorage.step once per frame with recording on (datastorage.py:25-34), its file thread
    stores `img_k.jpg` + `record_k.json` (:67-79, :98-112).  The thread numbers the records from 0 while the loaders read from 1
    (keras_train.py:36): the first recorded frame is never loaded."""
    from TritonRacerSim.components.datastorage import DataStorage
    ds = DataStorage(storage_path=folder)
    keys = ds.step_inputs[:-2]
    for f, v in zip(frames, values):
        row = dict(v)
        row['cam/img'] = f
        ds.step(*[row[k] for k in keys], False, True)                            # usr/del_record, usr/toggle_record
    last = os.path.join(folder, f"record_{len(frames) - 1}.json")
    for _ in range(2000):
        if os.path.exists(last) and os.path.getsize(last) > 0:
            break
        time.sleep(0.005)
    time.sleep(0.05)
    ds.on = False
    return keys


def main():
    n, h, w = 8, 24, 32
    frames = synth.frame_pool(n, h, w, seed=61)
    rng = np.random.default_rng(61)
    values = [{"mux/steering": float(rng.uniform(-1, 1)), "mux/throttle": float(rng.uniform(-1, 1)), "mux/break": 0.0,
               "gym/speed": float(rng.uniform(0, 25)), "gym/cte": float(rng.normal()), "loc/segment": float(rng.uniform(0, 10)),
               "gym/x": float(rng.normal(50, 20)), "gym/y": 0.56, "gym/z": float(rng.normal(30, 20))} for _ in range(n)]
    arrays = {}
    with tempfile.TemporaryDirectory() as tmp:
        full, gap = os.path.join(tmp, "records_1"), os.path.join(tmp, "records_2")
        write_tub(full, frames, values)
        shutil.copytree(full, gap)
        os.remove(os.path.join(gap, "record_4.json"))
        names = sorted(os.listdir(full))
        assert names == sorted([f"img_{i}.jpg" for i in range(n)] + [f"record_{i}.json" for i in range(n)]), names
        files, records = [], []
        for i in range(n):                                                        # every file the recorder wrote, index 0 included
            with open(os.path.join(full, f"img_{i}.jpg"), "rb") as fh:
                files.append(fh.read())
            with open(os.path.join(full, f"record_{i}.json")) as fh:
                records.append(json.load(fh))
        arrays["records_json"] = np.frombuffer(json.dumps(records).encode(), np.uint8)
        arrays["jpeg_blob"] = np.frombuffer(b"".join(files), np.uint8)
        arrays["jpeg_sizes"] = np.asarray([len(f) for f in files], np.int64)
        arrays["recorded_frames_u8"] = frames
        for cname in LOADER_CLASSES:
            loader = getattr(ref_train, cname)(full)
            loader.load(train_val_split=0.8, batch_size=2)
            assert len(loader.dataset) == n - 1                                   # records 1 .. n-1
            imgs = np.stack([d[0] for d in loader.dataset])
            assert imgs.dtype == np.float32
            u8 = np.rint(imgs * 255).astype(np.uint8)
            assert np.array_equal(u8.astype(np.float32) / np.float32(255), imgs)     # the stored bytes reproduce the float images exactly
            arrays.setdefault("frames_u8", u8)
            assert np.array_equal(arrays["frames_u8"], u8)
            arrays[f"labels/{cname}"] = np.stack([d[2] for d in loader.dataset])
            feats = [d[1] for d in loader.dataset]
            arrays[f"features/{cname}"] = np.stack(feats)
            assert arrays[f"labels/{cname}"].dtype == np.float32
            short = getattr(ref_train, cname)(gap)
            short.load(train_val_split=0.8, batch_size=1)
            arrays.setdefault("count_with_record_4_missing", np.asarray(len(short.dataset)))
            assert int(arrays["count_with_record_4_missing"]) == len(short.dataset) == 3
    np.savez_compressed(os.path.join(HERE, "tub.npz"), **arrays)
    print("tub.npz:", {k: v.shape for k, v in arrays.items()})


if __name__ == "__main__":
    main()
