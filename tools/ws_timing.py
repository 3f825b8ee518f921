import sys, torch
sys.path.insert(0, '.')
from triton_racer_sim_b200 import ImgPreprocessing, synth
from triton_racer_sim_b200.config import full_house_config
pool = torch.from_numpy(synth.frame_pool(1024, 120, 160)).cuda()
batch = synth.expand_torch(pool, 16384)
comp = ImgPreprocessing(full_house_config(), device=0, collect_stats=True)
out_u8 = torch.empty_like(batch); out_f32 = torch.empty(batch.shape, dtype=torch.float32, device='cuda')
for _ in range(3):
    comp.process_device(batch, out_u8=out_u8, out_f32=out_f32, want_f32=True)
st = comp.stats()
import os
ctas = 148 if os.environ.get('TRS_WS') else 296
per = {k: v / ctas / (16384 / ctas) for k, v in st.items() if k.startswith('t_')}
print({k: round(v) for k, v in per.items()}, 'cycles per frame per CTA; sweeps/frame', st['hyst_sweeps'] / st['frames'])
