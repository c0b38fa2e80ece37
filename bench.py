#!/usr/bin/env python3
"""Benchmark of the MAGICAL hot path (physics x10 + score + render + stack).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload cluster65536|mtc4096] [--batch B] [--gather-obs]

One "step" = one env-step of every environment in the batch (synthetic random
actions, Demo scene, auto-reset).  Prints ONE JSON line (rank 0).

  value : env-steps/s, whole job, actions already resident in HBM, timed with
          CUDA events on the launching stream, max over ranks.
  e2e   : the same metric through the public API with HOST action buffers:
          every step copies the actions H2D from pinned memory and reads the
          step's result (reward, done, eval_score) back D2H.  The observation
          stays in the device tensor, which is what the API returns.
  roofline : the dominant kernel (k_physics or k_raster, whichever takes
          longer per step) timed alone with CUDA events; achieved =
          algorithmic bytes per launch / launch duration.
  cpu_baseline : the CPU oracle (a C restatement of the reference's
          pymunk+render path; the reference itself cannot be installed here,
          see DESIGN.md) on a bounded sample, rank 0, N=1 only.

--impl reference times that CPU oracle on all host cores instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[2]: the north-star target (many-body contact stress)
    'cluster65536': ('ClusterColour-Demo-LoRes4E-v0', 65536),
    # BASELINE.json configs[1]
    'mtc4096': ('MoveToCorner-Demo-LoRes4E-v0', 4096),
    # BASELINE.json configs[3], one GPU's share (65536 / 8): randomised scenes from a pool of 64
    'mr_testall8192': ('MatchRegions-TestAll-LoResStack-v0', 8192),
}
# algorithmic HBM bytes per env-step, LoRes4E
OBS_WRITE = 96 * 96 * 12      # 110 592 B: the stacked observation written
STACK_READ = 96 * 96 * 9      # 82 944 B: the three surviving frames read back
STATE_RW = 6144               # survey's figure for the physics state read+write
SCALARS = 13                  # action i32 + reward f32 + done u8 + score f32
BYTES_PER_STEP_TOTAL = OBS_WRITE + STACK_READ + STATE_RW + SCALARS  # 199 693


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured'
    return 6650.0, 'fallback'  # B200_PROFILING.md fallback


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    QUERY = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ['nvidia-smi', f'--query-gpu={self.QUERY}',
                     '--format=csv,noheader,nounits', '-i', str(self.index)],
                    capture_output=True, text=True, timeout=5).stdout.strip()
                parts = [p.strip() for p in out.split(',')]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                 'sw_power_cap']
        reasons = [n for i, n in enumerate(names)
                   if any(s[2 + i].lower().startswith('active')
                          for s in self.samples)]
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(self.samples[0][1]),
                'reasons': reasons, 'samples': len(sm)}


# ----------------------------------------------------------------- CPU arm
def _oracle_worker(args):
    env_id, seconds, seed = args
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import magical_b200 as magical
    from oracle_lib import OracleEnv
    task, spec = magical.make_task(env_id)
    orc = OracleEnv(task.build_scene())
    rng = np.random.RandomState(seed)
    view = 0 if spec.preproc == 'LoRes4A' else 1
    n = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        for _ in range(8):
            _, done, _ = orc.step(int(rng.randint(18)))
            orc.render_lores(view)  # 384x384 render + 4x4 area mean
            if done:
                orc.reset()
            n += 1
    return n, time.perf_counter() - t0


def cpu_baseline(env_id, seconds, cores):
    """CPU oracle (physics + full-res render + downsample) on `cores`
    processes, one environment each, for ~`seconds` seconds."""
    import __graft_entry__ as entry
    entry.build_oracle()
    if cores == 1:
        n, dt = _oracle_worker((env_id, seconds, 7))
        return n / dt
    import multiprocessing as mp
    ctx = mp.get_context('spawn')
    with ctx.Pool(cores) as pool:
        res = pool.map(_oracle_worker,
                       [(env_id, seconds, 7 + i) for i in range(cores)])
    return sum(n / dt for n, dt in res)


def run_reference(args, env_id, batch):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # each "step" is a bounded sample of the workload; keep the whole run within ~2 minutes
    per_step_s = max(1.0, min(5.0, 120.0 / max(args.steps + args.warmup, 1)))
    vals = []
    for i in range(args.warmup + args.steps):
        v = cpu_baseline(env_id, per_step_s if i >= args.warmup else min(2.0, per_step_s), cores)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    sample = (f'{cores} processes x 1 env each, {per_step_s:.0f} s per step, '
              f'{env_id} physics + 384x384 ego render + 4x4 mean')
    line = {
        'impl': 'reference', 'metric': 'env_steps_per_s', 'value': value,
        'unit': 'env-steps/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1000.0 * batch / value,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': f'{env_id} batch {batch} per GPU, random '
                               'actions, auto-reset'},
        'cpu_baseline': {'value': value, 'unit': 'env-steps/s', 'cores': cores,
                         'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'env-steps/s',
                'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'note': 'CPU oracle (C restatement of the pymunk + GL + cv2 path); the '
                'reference itself is not installable offline',
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------- GPU arm
def run_b200(args, env_id, batch):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.gpus > 1 and world == 1:
        raise SystemExit('for --gpus N>1 launch with torch.distributed.run '
                         '(one rank per GPU)')
    if rank == 0:
        entry.build_cuda()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        dist.barrier()
    import magical_b200 as magical
    from magical_b200 import dist as mdist
    dev = torch.device('cuda', local_rank)
    K, W = args.steps, args.warmup

    # fixed seed: the randomised variants sample their scene pool from it (reproducible workload)
    venv = magical.make_vec(env_id, batch, device=local_rank, auto_reset=True, seed=1234 + rank)
    venv.reset()
    gen = torch.Generator(device=dev)
    gen.manual_seed(42 + rank)
    n_pool = 16
    act_pool = torch.randint(0, 18, (n_pool, batch), dtype=torch.int32,
                             device=dev, generator=gen)
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_scalars(rew, done, info):
        if world > 1:
            packed = mdist.pack_scalars(rew, done, info['eval_score'])
            out = torch.empty((world * batch, 3), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(out, packed)
            if args.gather_obs:
                g = torch.empty((world,) + tuple(venv.obs.shape),
                                dtype=torch.uint8, device=dev)
                dist.all_gather_into_tensor(g, venv.obs)

    # ---- prelude (untimed): spread the episode phases uniformly over the batch.  All envs start at
    # step 0 with the blocks apart; contact load grows over an episode, so timing the first steps of
    # synchronised episodes would flatter the physics.  One episode's worth of steps, resetting
    # 1/max_steps of the envs after each, leaves env i at phase (i mod max_steps): the steady-state mix
    # of a long auto-resetting rollout.
    prelude = venv.max_episode_steps if args.prelude < 0 else args.prelude
    if prelude > 0:
        ids = np.arange(batch)
        for t in range(prelude):
            venv.step(act_pool[t % n_pool])
            sel = ids[ids % prelude == t]
            if len(sel):
                venv.reset(env_ids=sel)
    # ---- value: device-resident actions
    for i in range(W):
        obs, rew, done, info = venv.step(act_pool[i % n_pool])
        gather_scalars(rew, done, info)
    launches0 = venv.launch_count()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        ev0.record(stream)
        for i in range(K):
            obs, rew, done, info = venv.step(act_pool[i % n_pool])
            gather_scalars(rew, done, info)
        ev1.record(stream)
        barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    launches = venv.launch_count() - launches0
    value = world * batch * K / (ms_total / 1000.0)

    # ---- e2e: host actions in, host results out, every step
    host_actions = torch.randint(0, 18, (max(K, 1), batch), dtype=torch.int32).pin_memory()
    dev_actions = torch.empty(batch, dtype=torch.int32, device=dev)
    host_rew = torch.empty(batch, dtype=torch.float32).pin_memory()
    host_done = torch.empty(batch, dtype=torch.uint8).pin_memory()
    host_score = torch.empty(batch, dtype=torch.float32).pin_memory()

    def e2e_step(i):
        dev_actions.copy_(host_actions[i % len(host_actions)], non_blocking=True)
        obs, rew, done, info = venv.step(dev_actions)
        host_rew.copy_(rew, non_blocking=True)
        host_done.copy_(done, non_blocking=True)
        host_score.copy_(info['eval_score'], non_blocking=True)
        stream.synchronize()  # the caller needs done/score before choosing the next action
        return obs

    for i in range(W):
        e2e_step(i)
    barrier()
    ev0.record(stream)
    for i in range(K):
        e2e_step(i)
    ev1.record(stream)
    barrier()
    e2e_ms = max_over_ranks(ev0.elapsed_time(ev1))
    e2e_value = world * batch * K / (e2e_ms / 1000.0)

    # ---- per-kernel timing for the roofline (this rank's GPU, kernels alone)
    def time_loop(fn, n):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for i in range(n):
            fn(i)
        b.record(stream)
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    n_views = 2 if 'LoResStack' in env_id else 1  # LoResStack moves both views
    n_k = max(K, 3)
    phys_ms = time_loop(lambda i: venv.step_physics(act_pool[i % n_pool]), n_k)
    rast_ms = time_loop(lambda i: venv.render(), n_k)
    peak, peak_kind = peaks()
    if phys_ms >= rast_ms:
        kname, kms = 'k_physics_tpe (+k_finish)', phys_ms
        alg = (STATE_RW + SCALARS) * batch
    else:
        kname, kms = 'k_raster', rast_ms
        alg = n_views * (OBS_WRITE + STACK_READ) * batch
    achieved = alg / (kms / 1000.0) / 1e9
    roofline = {
        'bound': 'hbm', 'kernel': kname, 'achieved': achieved, 'peak': peak,
        'unit': 'GB/s', 'frac': achieved / peak, 'traffic': None,
        'peak_source': peak_kind,
        'kernel_ms': {'k_physics_tpe+k_finish': phys_ms, 'k_raster': rast_ms},
        'algorithmic_bytes_per_launch': alg,
        'physics_achieved_gbs': (STATE_RW + SCALARS) * batch / (phys_ms / 1000.0) / 1e9,
        'whole_step_achieved_gbs': (n_views * (OBS_WRITE + STACK_READ) + STATE_RW + SCALARS) * batch
        / ((ms_total / K) / 1000.0) / 1e9,
        'raster_achieved_gbs': n_views * (OBS_WRITE + STACK_READ) * batch / (rast_ms / 1000.0) / 1e9,
        'note': 'the path is instruction-issue / dependent-latency bound (fp64 '
                'sequential-impulse solver), not HBM bound; see DESIGN.md',
    }
    traffic_file = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as fh:
                tr = json.load(fh)
            # measured at the default workload's batch; scale per environment for other batches
            per_env = tr.get(kname.split(' ')[0])
            if per_env is not None and tr.get('_env_id', env_id) == env_id:
                roofline['traffic'] = per_env / tr.get('_batch', batch) * batch
                roofline['traffic_source'] = tr.get('_source')
        except Exception:  # noqa: BLE001
            pass

    line = {
        'metric': 'env_steps_per_s', 'value': value, 'unit': 'env-steps/s',
        'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': ms_total / K,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {
            'workload': f'{env_id} batch {batch} per GPU, random actions, auto-reset',
            'env_id': env_id, 'batch_per_gpu': batch, 'global_batch': batch * world,
            'parallelism': f'env-sharded dp{world}; all-gather of reward/done/score'
                           + (' + obs' if args.gather_obs else ''),
            'episode_phase': (f'uniform over the {prelude}-step episode (untimed prelude of {prelude} steps '
                              'with staggered resets)') if prelude > 0 else 'all envs at episode start',
            'l2': 'state (254 MB) + obs (7.2 GB) per step exceed the 126 MB L2'
                  if batch >= 65536 else 'inputs smaller than L2 (small-batch config)',
        },
        'e2e': {'value': e2e_value, 'unit': 'env-steps/s', 'ms_per_step': e2e_ms / K,
                'h2d_bytes_per_step': 4 * batch, 'd2h_bytes_per_step': 9 * batch},
        'gpu_launches': int(launches),
        # environments x episodes that hit a physics capacity limit on this rank since the handle was created
        # (prelude + warm-up + all timed loops); anything but 0 would mean results that differ from the reference's
        'overflow_envs': int(venv.overflow_count()),
        'clocks': clocks.summary(),
        'roofline': roofline,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v = cpu_baseline(env_id, 12.0, 1)
        line['cpu_baseline'] = {
            'value': v, 'unit': 'env-steps/s', 'cores': 1, 'kind': 'port',
            'sample': f'1 env, 12 s of {env_id} random-action steps: physics + '
                      '384x384 ego render + 4x4 mean (CPU oracle, 1 thread)'}
    venv.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=40)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='cluster65536', choices=list(WORKLOADS))
    ap.add_argument('--batch', type=int, default=None)
    ap.add_argument('--gather-obs', action='store_true')
    ap.add_argument('--prelude', type=int, default=-1,
                    help='untimed steps that stagger the episode phases (-1: one episode, 0: none)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    env_id, batch = WORKLOADS[args.workload]
    if args.batch:
        batch = args.batch
    if args.impl == 'reference':
        run_reference(args, env_id, batch)
    else:
        run_b200(args, env_id, batch)


if __name__ == '__main__':
    main()
