import sys, os
sys.path.insert(0, os.getcwd())
import torch, magical_b200 as magical
env_id = sys.argv[1] if len(sys.argv) > 1 else 'ClusterColour-Demo-LoRes4E-v0'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
venv = magical.make_vec(env_id, B, auto_reset=True)
venv.reset()
g = torch.Generator(device='cuda'); g.manual_seed(0)
for i in range(int(sys.argv[3]) if len(sys.argv) > 3 else 40):
    a = torch.randint(0, 18, (B,), dtype=torch.int32, device='cuda', generator=g)
    venv.step(a)
torch.cuda.synchronize()
print('done')
