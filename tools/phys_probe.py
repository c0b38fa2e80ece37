"""k_physics_tpe alone at 65 536 envs (steady-state mix) under the environment's current MG_TPE_* knobs."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import magical_b200 as magical
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
venv = magical.make_vec('ClusterColour-Demo-LoRes4E-v0', B, auto_reset=True, alloc_obs=False)
venv.reset()
g = torch.Generator(device='cuda'); g.manual_seed(0)
acts = [torch.randint(0, 18, (B,), dtype=torch.int32, device='cuda', generator=g) for _ in range(8)]
ids = np.arange(B)
for t in range(240):
    venv.step_physics(acts[t % 8])
    venv.reset(env_ids=ids[ids % 240 == t])
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(20):
    venv.step_physics(acts[i % 8])
b.record(); torch.cuda.synchronize()
print('knobs', {k: v for k, v in os.environ.items() if k.startswith('MG_')}, 'physics ms %.3f' % (a.elapsed_time(b) / 20),
      'overflow', venv.overflow_count(), flush=True)
