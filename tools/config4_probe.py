"""Kernel times of BASELINE config 4 (MatchRegions-TestAll-LoResStack) on one GPU: device-sampled resets vs a
host-sampled scene pool vs the Demo layout.  usage: config4_probe.py [batch]"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import magical_b200 as magical
from magical_b200 import _native

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
lib = _native.load()


def probe(env_id, stagger=1, **kw):
    venv = magical.make_vec(env_id, B, auto_reset=True, seed=1, **kw)
    venv.reset()
    g = torch.Generator(device='cuda'); g.manual_seed(0)
    ids = np.arange(B)
    for t in range(60):
        venv.step(torch.randint(0, 18, (B,), dtype=torch.int32, device='cuda', generator=g))
        venv.reset(env_ids=ids[(ids // stagger) % 60 == t])
    torch.cuda.synchronize()
    acts = torch.randint(0, 18, (B,), dtype=torch.int32, device='cuda', generator=g)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    tp = tr = ts = 0.0
    n = 10
    for i in range(n):
        ev[0].record(); venv.step_physics(acts); ev[1].record(); venv.step_render(); ev[2].record()
        torch.cuda.synchronize()
        tp += ev[0].elapsed_time(ev[1]); tr += ev[1].elapsed_time(ev[2])
    print(f'{env_id} {kw} B={B}: physics+finish {tp/n:.3f} ms, render {tr/n:.3f} ms, '
          f'{(tp+tr)/n*65536/B:.2f} ms per 65536 envs')
    venv.close()


probe('MatchRegions-Demo-LoRes4E-v0')
probe('MatchRegions-TestAll-LoRes4E-v0', n_scenes=1)
probe('MatchRegions-TestAll-LoRes4E-v0', n_scenes=64)
probe('MatchRegions-Demo-LoResStack-v0')
probe('MatchRegions-Demo-LoRes3EA-v0')
probe('MatchRegions-TestAll-LoResStack-v0', device_sampling=False)
probe('MatchRegions-TestAll-LoResStack-v0', device_sampling=True)
