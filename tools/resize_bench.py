import sys, torch
sys.path.insert(0, '.')
from triton_racer_sim_b200 import FrameNormalise, synth
pool = torch.from_numpy(synth.frame_pool(256, 240, 320)).cuda()
src = synth.expand_torch(pool, 4096)
cam = FrameNormalise(device=0, out_hw=(120, 160))
for _ in range(3): cam.normalise_device(src)
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): cam.normalise_device(src)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 20 * 1e-3
print(f"{4096 / t / 1e6:.2f} M frames/s, {4096 * 345600 / t / 1e9:.0f} GB/s physical")
