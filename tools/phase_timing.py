"""Per-phase cycles per frame per CTA of the resident fused kernel (thread 0's clock64 accounting, statistics run)."""
import os, sys, torch
sys.path.insert(0, '.')
from triton_racer_sim_b200 import ImgPreprocessing, synth
from triton_racer_sim_b200.config import full_house_config
h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (120, 160)
n = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
want_f32 = (sys.argv[4] != 'u8') if len(sys.argv) > 4 else True
pool = torch.from_numpy(synth.frame_pool(256, h, w)).cuda()
batch = synth.expand_torch(pool, n)
comp = ImgPreprocessing(full_house_config(), device=0, collect_stats=True)
out_u8 = torch.empty_like(batch); out_f32 = torch.empty(batch.shape, dtype=torch.float32, device='cuda') if want_f32 else None
comp.process_device(batch, out_u8=out_u8, out_f32=out_f32, want_f32=want_f32)
st = comp.stats()
names = (['wait_frame', 'p1_strip_walk', 'p2_nms', 'p3_hysteresis', 'p4_output', 'total'] if os.environ.get('TRS_NO_STORE_WARP') else
         ['wait_frame', 'p1_strip_walk', 'wait_store_warps', 'p2_nms', 'p3_hysteresis', 'total'])
keys = ["t_wait_frame", "t_strip_walk", "t_phase_a", "t_phase_b", "t_phase_c", "t_total"]
print({nm: round(st[k] / st['frames']) for nm, k in zip(names, keys)}, 'cycles per frame per CTA; sweeps/frame', st['hyst_sweeps'] / st['frames'])
