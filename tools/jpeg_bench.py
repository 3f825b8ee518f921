"""Throughput of the tub-ingestion path: N JPEG records (host bytes) -> GPU decode [-> full observation chain]."""
import io, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from PIL import Image
from triton_racer_sim_b200 import ImgPreprocessing, synth, tub
from triton_racer_sim_b200 import _native as nat
from triton_racer_sim_b200.config import full_house_config
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
pool = synth.frame_pool(256, 120, 160)
files = []
for f in pool:
    b = io.BytesIO(); Image.fromarray(f).save(b, format='JPEG'); files.append(b.getvalue())
files = [files[i % 256] for i in range(n)]
blob, offsets = tub.pack_files(files)
pinned = torch.empty(len(blob), dtype=torch.uint8, pin_memory=True)
pinned.numpy()[:] = blob
blob = pinned.numpy()                          # the packed tub in pinned host memory
print(f"{n} records, {len(blob) / n:.0f} B per file")
ctx = nat.Context(0)
out = torch.empty((n, 120, 160, 3), dtype=torch.uint8, device='cuda')
for _ in range(2):
    tub.decode_jpeg_batch((blob, offsets), hw=(120, 160), ctx=ctx, out=out)
torch.cuda.synchronize(); t0 = time.perf_counter()
reps = 5
for _ in range(reps):
    tub.decode_jpeg_batch((blob, offsets), hw=(120, 160), ctx=ctx, out=out)
torch.cuda.synchronize(); t = (time.perf_counter() - t0) / reps
print(f"decode (host parse + H2D + kernels): {n / t / 1e6:.3f} M records/s ({t * 1e3:.1f} ms)")
# CPU side: Pillow single thread
t0 = time.perf_counter()
for b in files[:512]:
    np.asarray(Image.open(io.BytesIO(b)))
print(f"Pillow, 1 thread: {512 / (time.perf_counter() - t0):.0f} records/s")
