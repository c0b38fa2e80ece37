#!/usr/bin/env python3
"""Benchmark of the MAGICAL hot path (physics x10 + score + render + stack).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload NAME] [--batch B] [--no-records]

One "step" = one env-step of every environment of the (global) batch:
synthetic random actions, auto-reset.  Prints ONE JSON line (rank 0).

Workloads (BASELINE.json configs):
  cluster65536    configs[2]  ClusterColour-Demo-LoRes4E, 65 536 envs PER GPU (weak scaling); N=1 default
  mtc4096         configs[1]  MoveToCorner-Demo-LoRes4E, 4 096 envs per GPU
  config4         configs[3]  MatchRegions-TestAll-LoResStack, 65 536 envs GLOBAL, sharded, observation
                              all-gather (newest frame over NVLink + k_stack_push); N>1 default
  config5         configs[4]  all 8 Demo tasks mixed (LoRes4E), 131 072 envs GLOBAL, sharded, obs all-gather
  mr_testall8192              config 4's per-GPU share at N=8 on one GPU (no gather)
At N>1 the line's `records` additionally hold config5 and the weak-scaling cluster65536 workload with and
without the observation gather (--no-records skips them).

  value : env-steps/s, whole job, actions already resident in HBM, CUDA events on the launching stream,
          median of 3 repeats of K steps, max over ranks.  At N>1 every rank ends each step with the global
          observation / reward / done / score batch (ShardedVecEnv, gather_obs='newest': flag barrier, then
          k_stack_push_p2p reads the other ranks' newest frames over NVLink and rebuilds their stacks).  The
          exchange is NOT overlapped with the next step's physics by default: measured on B200, the
          bandwidth-saturating stack rebuild slows the latency-bound physics kernel by more than its own
          duration when they run together (profiles/r02_bench_log.md); --pipeline turns the overlap on.
  e2e   : the same metric through the public API with HOST action buffers, closed loop: every step copies the
          (global) actions H2D from pinned memory, waits for the gather, and reads reward/done/eval_score of
          the global batch back D2H before the next step.  The observation stays in the device tensor the
          API returns.
  roofline : the dominant kernel timed alone with CUDA events; achieved = algorithmic bytes per launch /
          launch duration (SURVEY 8(d) figures).
  cpu_baseline : the CPU oracle (a C restatement of the reference's pymunk + GL + cv2 path; the reference
          itself cannot be installed here) on a bounded sample, rank 0, N=1 only.

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

DEMO_MIX = ['MoveToCorner', 'MoveToRegion', 'MatchRegions', 'MakeLine', 'FindDupe', 'FixColour',
            'ClusterColour', 'ClusterShape']
# name -> (env id or list of ids, batch, 'per_gpu' | 'global')
WORKLOADS = {
    'cluster65536': ('ClusterColour-Demo-LoRes4E-v0', 65536, 'per_gpu'),
    'mtc4096': ('MoveToCorner-Demo-LoRes4E-v0', 4096, 'per_gpu'),
    'mr_testall8192': ('MatchRegions-TestAll-LoResStack-v0', 8192, 'per_gpu'),
    'config4': ('MatchRegions-TestAll-LoResStack-v0', 65536, 'global'),
    'config5': ([f'{t}-Demo-LoRes4E-v0' for t in DEMO_MIX], 131072, 'global'),
}
# algorithmic HBM bytes per env-step and view
OBS_WRITE = 96 * 96 * 12      # 110 592 B: the stacked observation written
STACK_READ = 96 * 96 * 9      # 82 944 B: the three surviving frames read back
FRAME = 96 * 96 * 3           # 27 648 B: one frame
STATE_RW = 6144               # survey's figure for the physics state read+write
SCALARS = 13                  # action i32 + reward f32 + done u8 + score f32


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
            self._stop.wait(0.1)

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
    from oracle_lib import OracleEnv, downsample4
    task, spec = magical.make_task(env_id)
    orc = OracleEnv(task.build_scene())
    rng = np.random.RandomState(seed)
    n = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        for _ in range(8):
            _, done, _ = orc.step(int(rng.randint(18)))
            # BaseEnv.render always draws BOTH 384x384 views (base_env.py:326-336); the LoRes
            # preprocessors then downsample the view(s) they keep (4x4 area mean)
            allo = orc.render_view(0, 384)
            ego = orc.render_view(1, 384)
            if spec.preproc in ('LoRes4A', 'LoRes3EA', 'LoResStack'):
                downsample4(allo)
            if spec.preproc != 'LoRes4A':
                downsample4(ego)
            if done:
                if magical.EnvName(env_id).is_test:
                    # the reference samples a fresh layout at every reset (base_env.py:177-234)
                    orc.close()
                    orc = OracleEnv(task.build_scene())
                else:
                    orc.reset()
            n += 1
    return n, time.perf_counter() - t0


def cpu_baseline(env_id, seconds, cores):
    """CPU oracle (physics + both full-res views + downsample) on `cores` processes, one environment
    each, for ~`seconds` seconds.  Returns env-steps/s summed over the processes."""
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


def run_reference(args, env_id, batch, kind='per_gpu'):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    if isinstance(env_id, list):
        env_id = env_id[-2]  # the mix's heaviest member stands in (ClusterColour)
    cores = os.cpu_count() or 1
    # each "step" is a bounded sample of the workload; keep the whole run within ~2 minutes
    per_step_s = max(1.0, min(5.0, 120.0 / max(args.steps + args.warmup, 1)))
    vals, walls = [], []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        v = cpu_baseline(env_id, per_step_s if i >= args.warmup else min(2.0, per_step_s), cores)
        if i >= args.warmup:
            vals.append(v)
            walls.append(time.perf_counter() - t0)
    value = float(np.median(vals))
    sample = (f'{cores} processes x 1 env each, {per_step_s:.0f} s of random-action env-steps per bench step '
              f'({env_id}: physics + allo and ego 384x384 renders + 4x4 mean of the kept view); '
              f'value = env-steps/s summed over the processes, median of {len(vals)} samples')
    line = {
        'impl': 'reference', 'metric': 'env_steps_per_s', 'value': value,
        'unit': 'env-steps/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup,
        # wall time of one bench step = one bounded sample (process start-up included), NOT the time the CPU
        # would need for one env-step of the whole batch (that is batch / value seconds)
        'ms_per_step': 1000.0 * float(np.median(walls)),
        'seconds_per_batch_step_extrapolated': batch / value,
        'spread': {'min': float(np.min(vals)), 'max': float(np.max(vals)), 'n': len(vals)},
        'higher_is_better': True, 'scaling': 'weak' if kind == 'per_gpu' else 'strong', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': f'{env_id}, random actions, auto-reset; bounded sample of the batch-{batch} '
                               f'({"per GPU" if kind == "per_gpu" else "global"}) workload, one env per host core'},
        'cpu_baseline': {'value': value, 'unit': 'env-steps/s', 'cores': cores,
                         'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'env-steps/s',
                'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'note': 'CPU oracle (C restatement of the pymunk + GL + cv2 path); the '
                'reference itself is not installable offline',
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------- GPU arm
class Runner:
    """One workload on this rank's GPU (and its peers): builds the env, runs the untimed prelude, and
    measures value / e2e / per-kernel times."""

    def __init__(self, args, name, gather_obs, batch_override=None):
        import torch
        import torch.distributed as dist
        import magical_b200 as magical
        from magical_b200 import dist as mdist
        self.torch, self.dist, self.args, self.name = torch, dist, args, name
        self.rank = int(os.environ.get('RANK', '0'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.dev = torch.device('cuda', self.local_rank)
        env_id, batch, kind = WORKLOADS[name]
        if batch_override:
            batch = batch_override
        self.env_id = env_id
        self.label = env_id if isinstance(env_id, str) else '8-task Demo mix (' + ', '.join(DEMO_MIX) + ') LoRes4E'
        self.total = batch * self.world if kind == 'per_gpu' else batch
        assert self.total % self.world == 0
        self.n_local = self.total // self.world
        self.kind = kind
        self.gather_obs = gather_obs if self.world > 1 else False
        rank, world, total = self.rank, self.world, self.total
        sharded_obs = self.gather_obs == 'newest'

        def make_local(n):
            start = rank * n
            if isinstance(env_id, list):
                return magical.make_vec_mixed(env_id, n, device=self.local_rank, alloc_obs=not sharded_obs,
                                              first_env=start, total=total)
            # fixed seed: the randomised variants sample their scene pool from it (reproducible workload)
            # randomised (Test*) variants: the pool holds structure TEMPLATES and every reset samples a fresh
            # layout on the device, as the reference re-randomises on every reset
            return magical.make_vec(env_id, n, device=self.local_rank, auto_reset=True, seed=1234 + rank,
                                    alloc_obs=not sharded_obs,
                                    device_sampling=magical.EnvName(env_id).is_test)

        self.env = mdist.ShardedVecEnv(make_local, total, rank, world, gather_obs=self.gather_obs,
                                       pipeline=args.pipeline, transport=args.transport)
        self.venv = self.env.local
        self.views = 2 if self.venv.preproc == 'LoResStack' else 1
        self.env.reset()
        gen = torch.Generator(device=self.dev)
        gen.manual_seed(42)  # the GLOBAL action stream is the same on every rank
        self.n_pool = 16
        self.act_pool = torch.randint(0, 18, (self.n_pool, total), dtype=torch.int32, device=self.dev,
                                      generator=gen)
        self.stream = torch.cuda.current_stream()

    # -- helpers
    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        if self.world == 1:
            return ms
        t = self.torch.tensor([ms], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def prelude(self):
        """Untimed: spread the episode phases uniformly over the batch.  All envs start at step 0 with the
        blocks apart; contact load grows over an episode, so timing the first steps of synchronised episodes
        would flatter the physics.  One episode's worth of steps, resetting 1/max_steps of the envs after
        each, leaves env i at phase (i mod max_steps): the steady-state mix of a long auto-resetting rollout."""
        venv, args = self.venv, self.args
        n = venv.max_episode_steps if args.prelude < 0 else args.prelude
        self.prelude_steps = n
        if n <= 0:
            return
        ids = np.arange(self.n_local)
        lo = self.rank * self.n_local
        for t in range(n):
            venv.step(self.act_pool[t % self.n_pool][lo:lo + self.n_local])
            sel = ids[ids % n == t]
            if len(sel):
                venv.reset(env_ids=sel)
        if self.env._newest_ready:
            self.env._full_gather()   # the remote shards' stacks, once, after the local-only prelude
        self.torch.cuda.synchronize()

    def timed(self, step_fn, K, W, clocks=False):
        """W warm-up + 3 x K timed steps; returns (median ms per K steps, [ms], launches, clock summary)."""
        torch = self.torch
        for i in range(W):
            step_fn(i)
        self.env.wait_obs()
        reps, launches, summary = [], 0, None
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for r in range(3):
            self.barrier()
            l0 = self.venv.launch_count() + self.env.push_launches
            sampler = ClockSampler(self.local_rank) if (clocks and r == 0) else None
            if sampler:
                sampler.__enter__()
            ev0.record(self.stream)
            for i in range(K):
                step_fn(i)
            self.env.wait_obs()   # the last gather + stack push are inside the timed region
            ev1.record(self.stream)
            self.barrier()
            if sampler:
                sampler.__exit__()
                summary = sampler.summary()
            reps.append(self.max_over_ranks(ev0.elapsed_time(ev1)))
            launches = self.venv.launch_count() + self.env.push_launches - l0
        return float(np.median(reps)), reps, launches, summary

    def time_loop(self, fn, n):
        torch = self.torch
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(self.stream)
        for i in range(n):
            fn(i)
        b.record(self.stream)
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    # -- measurements
    def measure(self, K, W, clocks=False, e2e=True, kernels=True):
        torch = self.torch
        env, venv, total = self.env, self.venv, self.total
        self.prelude()
        res = {}
        ms, reps, launches, clk = self.timed(lambda i: env.step(self.act_pool[i % self.n_pool]), K, W, clocks)
        res.update(ms_total=ms, reps=reps, launches=launches, clocks=clk, value=total * K / (ms / 1000.0))
        if e2e:
            host_actions = torch.randint(0, 18, (max(K, 1), total), dtype=torch.int32).pin_memory()
            dev_actions = torch.empty(total, dtype=torch.int32, device=self.dev)
            host_rew = torch.empty(total, dtype=torch.float32).pin_memory()
            host_done = torch.empty(total, dtype=torch.uint8).pin_memory()
            host_score = torch.empty(total, dtype=torch.float32).pin_memory()

            def e2e_step(i):
                dev_actions.copy_(host_actions[i % len(host_actions)], non_blocking=True)
                obs, rew, done, info = env.step(dev_actions)
                env.wait_obs()   # closed loop: the global obs / scalars are complete before the read-back
                host_rew.copy_(rew, non_blocking=True)
                host_done.copy_(done, non_blocking=True)
                host_score.copy_(info['eval_score'], non_blocking=True)
                self.stream.synchronize()  # the caller needs done/score before choosing the next action
                return obs

            ms2, reps2, _, _ = self.timed(e2e_step, K, W)
            res['e2e'] = {'value': total * K / (ms2 / 1000.0), 'unit': 'env-steps/s', 'ms_per_step': ms2 / K,
                          'repeats_ms_per_step': [m / K for m in reps2],
                          'h2d_bytes_per_step': 4 * total, 'd2h_bytes_per_step': 9 * total}
        if kernels:
            lo = self.rank * self.n_local
            acts = [self.act_pool[i][lo:lo + self.n_local].contiguous() for i in range(self.n_pool)]
            n_k = max(K, 3)
            km = {'k_physics_tpe+k_finish': self.time_loop(lambda i: venv.step_physics(acts[i % self.n_pool]), n_k),
                  'k_raster': self.time_loop(lambda i: venv.step_render(), n_k)}
            if env._newest_ready:
                # p2p transport: k_stack_push_p2p reads the frames over NVLink itself and `exchange` is the flag
                # barrier + the 12 B/env scalars; nccl transport: `exchange` is the two ncclAllGathers
                # the push is timed behind the flag barrier, as in a step: ranks that start together walk
                # their peers in the staggered ring order (no NVLink egress hot spot)
                km['exchange'] = self.time_loop(lambda i: env.gather_only(), n_k)
                km['k_stack_push'] = self.time_loop(lambda i: (env.gather_only(), env.push_only()), n_k) \
                    - km['exchange']
            res['kernel_ms'] = km
        res['overflow_envs'] = int(venv.overflow_count())
        return res

    def roofline(self, res):
        """Roofline of the dominant kernel of this workload on this rank."""
        peak, peak_kind = peaks()
        km = res['kernel_ms']
        n, views = self.n_local, self.views
        remote = self.total - n
        alg = {'k_physics_tpe+k_finish': (STATE_RW + SCALARS) * n,
               # k_raster additionally writes the newest-frame send buffer when the obs gather is on
               'k_raster': views * (OBS_WRITE + STACK_READ + (FRAME if self.env._newest_ready else 0)) * n,
               'k_stack_push': views * (OBS_WRITE + STACK_READ + FRAME) * remote}
        kname = max((k for k in km if k in alg), key=lambda k: km[k])
        note = None
        # The render / stack kernels are the bandwidth-shaped ones; k_physics_tpe is fp64 latency-bound (its HBM
        # fraction is ~0.01 by construction).  When the two are level (within 25 %) the line keeps reporting the
        # bandwidth-bound kernel -- the one VERDICT tracks -- and says so; per_kernel_frac always holds both.
        streaming = [k for k in km if k in alg and k != 'k_physics_tpe+k_finish']
        if kname == 'k_physics_tpe+k_finish' and streaming:
            best = max(streaming, key=lambda k: km[k])
            if km[kname] <= 1.25 * km[best]:
                note = (f'{kname} is the longest launch ({km[kname]:.2f} ms vs {km[best]:.2f} ms) but is fp64 '
                        f'latency-bound (HBM fraction {alg[kname] / (km[kname] / 1000.0) / 1e9 / peak:.4f}); the '
                        f'roofline is given for the bandwidth-bound kernel {best}')
                kname = best
        achieved = alg[kname] / (km[kname] / 1000.0) / 1e9
        out = {'bound': 'hbm', 'kernel': kname, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
               'frac': achieved / peak, 'traffic': None, 'peak_source': peak_kind, 'kernel_ms': km,
               'algorithmic_bytes_per_launch': alg[kname],
               'per_kernel_frac': {k: alg[k] / (km[k] / 1000.0) / 1e9 / peak for k in km if k in alg},
               'whole_step_achieved_gbs': sum(alg[k] for k in km if k in alg) / (res['ms_total'] / self.args.steps / 1000.0) / 1e9}
        traffic_file = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.exists(traffic_file) and isinstance(self.env_id, str):
            try:
                with open(traffic_file) as fh:
                    tr = json.load(fh)
                per_env = tr.get(kname.split('+')[0])
                if per_env is not None and tr.get('_env_id') == self.env_id:
                    out['traffic'] = per_env / tr.get('_batch', n) * n
                    out['traffic_source'] = tr.get('_source')
            except Exception:  # noqa: BLE001
                pass
        if note:
            out['kernel_note'] = note
        if self.venv.device_sampling:
            out['note'] = ('k_physics_tpe+k_finish is timed in a physics-only loop, where the reset sampler '
                           '(k_sample_layouts, ~1 ms for the environments that finish an episode) runs serially '
                           'after it; inside a step it runs on a side stream next to the render')
        if 'exchange' in km:
            rx = self.env.nvlink_bytes_per_step()
            carrier = 'k_stack_push' if self.env.transport == 'p2p' else 'exchange'
            out['nvlink'] = {'transport': self.env.transport, 'rx_bytes_per_step_per_gpu': rx,
                             'carried_by': carrier + (' (frames read from the owners\' buffers inside the kernel)'
                                                      if carrier == 'k_stack_push' else ' (ncclAllGather)'),
                             'ms': km[carrier], 'achieved_rx_gbs': rx / (km[carrier] / 1000.0) / 1e9,
                             'cap_gbs_per_direction': 900.0}
        return out

    def config(self):
        per = self.n_local
        par = f'env-sharded dp{self.world}'
        if self.world > 1:
            par += '; all-gather of reward/done/score'
            if self.gather_obs == 'newest':
                par += (f' + obs (transport {self.env.transport}: newest frame of every env over NVLink, '
                        'k_stack_push rebuilds the global stacks on every rank'
                        + ("; exchange + push overlap the next step's physics" if self.args.pipeline else '') + ')')
        return {
            'workload': f'{self.label}: {self.total} envs global, {per} per GPU, random actions, auto-reset',
            'env_id': self.env_id if isinstance(self.env_id, str) else self.env_id,
            'batch_per_gpu': per, 'global_batch': self.total, 'parallelism': par,
            'resets': ('device-side rejection sampling of goal sizes and poses at every reset (k_sample_layouts), '
                       f'{self.venv.n_scenes} host-built structure templates') if self.venv.device_sampling
            else 'deterministic Demo layout',
            'episode_phase': (f'uniform over the {self.prelude_steps}-step episode (untimed prelude of '
                              f'{self.prelude_steps} steps with staggered resets)')
            if self.prelude_steps > 0 else 'all envs at episode start',
            'l2': (f'obs ({self.views * per * OBS_WRITE / 1e9:.2f} GB) + state ({per * 3872 / 1e6:.0f} MB) per step '
                   'exceed the 126 MB L2') if self.views * per * OBS_WRITE > 2.5e8
            else 'per-step working set below 2x the 126 MB L2 and NOT flushed (small-batch run, not a BASELINE config)',
        }

    def close(self):
        self.env.close()


def run_b200(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.gpus > 1 and world == 1:
        raise SystemExit('for --gpus N>1 launch with torch.distributed.run (one rank per GPU)')
    if rank == 0:
        entry.build_cuda()
    torch.cuda.set_device(local_rank)
    if world > 1:
        if not os.environ.get('MG_KEEP_NCCL_DEBUG'):
            os.environ['NCCL_DEBUG'] = 'WARN'   # keep NCCL's version banner off stdout (ONE JSON line)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        dist.barrier()
    K, W = args.steps, args.warmup
    name = args.workload or ('cluster65536' if world == 1 else 'config4')
    gather = 'newest' if world > 1 and not args.no_gather_obs else False

    run = Runner(args, name, gather, args.batch)
    res = run.measure(K, W, clocks=True)
    roof = run.roofline(res)
    scaling = 'weak' if run.kind == 'per_gpu' else 'strong'
    line = {
        'metric': 'env_steps_per_s', 'value': res['value'], 'unit': 'env-steps/s',
        'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': res['ms_total'] / K,
        'repeats_ms_per_step': [m / K for m in res['reps']],
        'higher_is_better': True, 'scaling': scaling, 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic', 'config': run.config(),
        'e2e': res['e2e'], 'gpu_launches': int(res['launches']),
        # environments x episodes that hit a physics capacity limit on this rank since the handle was created
        # (prelude + warm-up + all timed loops); anything but 0 would mean results that differ from the reference's
        'overflow_envs': res['overflow_envs'], 'clocks': res['clocks'], 'roofline': roof,
    }
    env_id_for_cpu = run.env_id if isinstance(run.env_id, str) else run.env_id[-2]
    run.close()
    del run
    torch.cuda.empty_cache()

    # ---- further records of the same run (N>1): BASELINE configs[4] and the weak-scaling headline workload
    if not args.no_records and args.workload is None:
        records = []
        # N>1: BASELINE configs[4] and the weak-scaling headline workload with / without the observation gather;
        # N=1: configs[3] and [4] on one GPU, the base points of their strong-scaling curves
        plan = (('config5', 'newest'), ('cluster65536', 'newest'), ('cluster65536', False)) if world > 1 \
            else (('config4', False), ('config5', False))
        for rname, rgather in plan:
            r = Runner(args, rname, rgather)
            rr = r.measure(K, W, e2e=True, kernels=True)
            records.append({
                'workload': rname, 'value': rr['value'], 'unit': 'env-steps/s', 'ms_per_step': rr['ms_total'] / K,
                'repeats_ms_per_step': [m / K for m in rr['reps']],
                'scaling': 'weak' if r.kind == 'per_gpu' else 'strong', 'config': r.config(),
                'e2e': rr['e2e'], 'roofline': r.roofline(rr), 'overflow_envs': rr['overflow_envs']})
            r.close()
            del r
            torch.cuda.empty_cache()
        line['records'] = records
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v = cpu_baseline(env_id_for_cpu, 12.0, 1)
        line['cpu_baseline'] = {
            'value': v, 'unit': 'env-steps/s', 'cores': 1, 'kind': 'port',
            'sample': f'1 env, 12 s of {env_id_for_cpu} random-action steps: physics + allo and ego 384x384 '
                      'renders + 4x4 mean of the kept view (CPU oracle, 1 thread)'}
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
    ap.add_argument('--workload', default=None, choices=list(WORKLOADS))
    ap.add_argument('--batch', type=int, default=None)
    ap.add_argument('--no-gather-obs', action='store_true',
                    help='N>1: gather only reward/done/score (round-1 behaviour)')
    ap.add_argument('--no-records', action='store_true')
    ap.add_argument('--pipeline', action='store_true',
                    help='N>1: let the exchange + stack rebuild of step t overlap the physics of step t+1')
    ap.add_argument('--transport', default='auto', choices=['auto', 'p2p', 'nccl'],
                    help='N>1 observation gather: NVLink peer memory (default) or ncclAllGather')
    ap.add_argument('--prelude', type=int, default=-1,
                    help='untimed steps that stagger the episode phases (-1: one episode, 0: none)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.impl == 'reference':
        world = int(os.environ.get('WORLD_SIZE', '1'))
        env_id, batch, kind = WORKLOADS[args.workload or ('cluster65536' if max(world, args.gpus) == 1 else 'config4')]
        run_reference(args, env_id, args.batch or batch, kind)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
