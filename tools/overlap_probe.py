"""How do k_stack_push and k_physics_tpe / k_raster share one GPU?  (single GPU, local buffers)"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import magical_b200 as magical
from magical_b200 import dist as mdist
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
venv = magical.make_vec('ClusterColour-Demo-LoRes4E-v0', B, auto_reset=True)
venv.reset()
g = torch.Generator(device='cuda'); g.manual_seed(0)
acts = [torch.randint(0, 18, (B,), dtype=torch.int32, device='cuda', generator=g) for _ in range(8)]
ids = np.arange(B)
for t in range(120):
    venv.step(acts[t % 8])
    venv.reset(env_ids=ids[ids % 120 == t])
stacks = torch.zeros((B, 96, 96, 12), dtype=torch.uint8, device='cuda')
newest = torch.zeros((B, 96, 96, 3), dtype=torch.uint8, device='cuda')
cur = torch.cuda.current_stream()
def ev(): return torch.cuda.Event(enable_timing=True)
def run(label, main, side, prio, order):
    s2 = torch.cuda.Stream(priority=prio)
    res = []
    for rep in range(4):
        torch.cuda.synchronize()
        a0, a1, b0, b1, go = ev(), ev(), ev(), ev(), torch.cuda.Event()
        a0.record(cur); go.record(cur)
        def do_main():
            if main: main()
            a1.record(cur)
        def do_side():
            with torch.cuda.stream(s2):
                s2.wait_event(go)
                b0.record(s2)
                if side: side(s2)
                b1.record(s2)
        if order == 'main_first':
            do_main(); do_side()
        else:
            do_side(); do_main()
        torch.cuda.synchronize()
        res.append((a0.elapsed_time(a1), a0.elapsed_time(b1)))
    m = np.median(np.array(res), axis=0)
    print(f'{label:60s} main {m[0]:6.2f} ms   both done {m[1]:6.2f} ms', flush=True)
phys = lambda: venv.step_physics(acts[0])
rast = lambda: venv.step_render()
push = lambda s: mdist.cuda_stack_push(stacks, newest.view(-1), None, 0, B, B, 0, stream=s)
run('physics alone', phys, None, 0, 'main_first')
run('raster alone', rast, None, 0, 'main_first')
run('push alone', None, push, 0, 'main_first')
for prio in (0, -1):
    for order in ('main_first', 'side_first'):
        run(f'physics || push  prio {prio} {order}', phys, push, prio, order)
        run(f'raster  || push  prio {prio} {order}', rast, push, prio, order)
