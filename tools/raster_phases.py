"""Per-phase cycle shares of k_raster (build with MG_EXTRA_NVCC_FLAGS=-DRASTER_PROF)."""
import ctypes, os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import magical_b200 as magical
from magical_b200 import _native
env_id = sys.argv[1] if len(sys.argv) > 1 else 'ClusterColour-Demo-LoRes4E-v0'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
venv = magical.make_vec(env_id, B, auto_reset=True, seed=1)
venv.reset()
g = torch.Generator(device='cuda'); g.manual_seed(0)
ids = np.arange(B)
for t in range(60):
    venv.step(torch.randint(0, 18, (B,), dtype=torch.int32, device='cuda', generator=g))
    venv.reset(env_ids=ids[ids % 60 == t])
torch.cuda.synchronize()
lib = _native.load()
buf = (ctypes.c_uint64 * 16)()
lib.mg_raster_prof_read(buf, 1)
n = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(n):
    venv.step_render()
e1.record()
torch.cuda.synchronize()
lib.mg_raster_prof_read(buf, 0)
raw = np.array(list(buf), dtype=np.float64) / (n * B)
v = np.concatenate([raw[:7], raw[8:10], raw[7:8]])
names = ['A (static now)', 'B verts', 'C prim records', 'D edge eq', 'E0 span init', 'E1 fold', 'F bins', 'G0 tile desc', 'G1 flat tiles', 'G2 busy tiles']
print(f'{env_id} B={B}: k_raster {e0.elapsed_time(e1)/n:.3f} ms per launch (instrumented)')
for nm, c in zip(names, v):
    print(f'  {nm:16s} {c:9.0f} cycles/env  {100*c/v.sum():5.1f} %')
print(f'  total            {v.sum():9.0f} cycles/env')
c = np.zeros(4)
print(f'  per env: flat tiles {c[0]:.1f}, non-flat tiles {c[1]:.1f}, prim visits (per 4-px group) {c[2]:.0f} = {c[2]/max(c[1]*16,1):.2f} per group of a non-flat tile')
