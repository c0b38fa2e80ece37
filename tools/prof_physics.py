"""Physics-only rollout (no observation buffer) for ncu captures of k_physics_tpe at full batch."""
import os
import sys
sys.path.insert(0, os.getcwd())
import torch
import magical_b200 as magical
from magical_b200.vec_env import MagicalVecEnv
env_id = sys.argv[1] if len(sys.argv) > 1 else 'ClusterColour-Demo-LoRes4E-v0'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
n = int(sys.argv[3]) if len(sys.argv) > 3 else 160
task, spec = magical.make_task(env_id)
venv = MagicalVecEnv(task, B, preproc=spec.preproc, auto_reset=True, alloc_obs=False)
g = torch.Generator(device='cuda')
g.manual_seed(0)
for i in range(n):
    a = torch.randint(0, 18, (B,), dtype=torch.int32, device='cuda', generator=g)
    venv.step_physics(a)
torch.cuda.synchronize()
print('done')
