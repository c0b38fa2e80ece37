import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import magical_b200 as magical
B = 16384
def probe(env_id, seed):
    venv = magical.make_vec(env_id, B, auto_reset=True, seed=seed, n_scenes=1, alloc_obs=False)
    venv.reset()
    g = torch.Generator(device='cuda'); g.manual_seed(0)
    ids = np.arange(B)
    for t in range(60):
        venv.step_physics(torch.randint(0, 18, (B,), dtype=torch.int32, device='cuda', generator=g))
        venv.reset(env_ids=ids[ids % 60 == t])
    torch.cuda.synchronize()
    acts = torch.randint(0, 18, (B,), dtype=torch.int32, device='cuda', generator=g)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10): venv.step_physics(acts)
    e1.record(); torch.cuda.synchronize()
    sc = venv.scenes[0]
    lc = [int(venv.get_state(i)['n_cache']) for i in range(0, 2048, 64)]
    kinds = [int(sc['shapes'][k]['kind']) if 'kind' in sc['shapes'].dtype.names else -1 for k in range(int(sc['n_shapes']))]
    print(f"{env_id} seed {seed}: bodies {int(sc['n_bodies'])} shapes {int(sc['n_shapes'])} bpairs {int(sc['n_bpairs']) if 'n_bpairs' in sc.dtype.names else '?'} physics {e0.elapsed_time(e1)/10:.3f} ms  cache entries mean {np.mean(lc):.1f} max {max(lc)}", flush=True)
    venv.close()
probe('MatchRegions-Demo-LoRes4E-v0', 0)
for s in range(6):
    probe('MatchRegions-TestAll-LoRes4E-v0', s)
probe('ClusterColour-Demo-LoRes4E-v0', 0)
probe('ClusterColour-TestAll-LoRes4E-v0', 0)
probe('ClusterColour-TestAll-LoRes4E-v0', 1)
