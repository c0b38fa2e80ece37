"""Multi-process host logic of the env-sharded path on CPU: world_size 2, gloo.
The GPU env is replaced by a deterministic stand-in so only the sharding /
all-gather plumbing of magical_b200.dist is exercised."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


class FakeLocalEnv:
    """Deterministic per-env arithmetic keyed by GLOBAL env index, so that the
    union of the shards must equal the single-process result."""

    def __init__(self, start, n):
        import torch
        self.idx = torch.arange(start, start + n, dtype=torch.float32)
        self.t = 0

    def reset(self):
        import torch
        return (self.idx[:, None] * torch.ones(1, 4)).to(torch.uint8)

    def step(self, actions):
        import torch
        self.t += 1
        obs = ((self.idx + actions.float() + self.t)[:, None] * torch.ones(1, 4)).to(torch.uint8)
        rew = self.idx * 0.5 + actions.float()
        done = ((self.idx.long() + self.t) % 3 == 0).to(torch.uint8)
        score = torch.where(done.bool(), self.idx / 100.0, torch.zeros_like(self.idx))
        return obs, rew, done, {'eval_score': score}

    def close(self):
        pass


def _worker(rank, world, port, total, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch
    import torch.distributed as dist
    from magical_b200 import dist as mdist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    start, stop = mdist.shard_range(total, rank, world)
    env = mdist.ShardedVecEnv(lambda n: FakeLocalEnv(start, n), total, rank, world, gather_obs=True)
    obs0 = env.reset()
    g = torch.Generator().manual_seed(0)
    res = []
    for t in range(4):
        actions = torch.randint(0, 18, (total,), generator=g, dtype=torch.int32)
        obs, rew, done, info = env.step(actions)
        res.append((obs.numpy(), rew.numpy(), done.numpy(), info['eval_score'].numpy()))
    if rank == 0:
        steps = np.stack([np.concatenate([o.reshape(total, -1).astype(np.float64), r[:, None], d[:, None],
                                          s[:, None]], axis=1) for o, r, d, s in res])
        np.savez(out, obs0=obs0.numpy(), steps=steps)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_the_batch():
    from magical_b200 import dist as mdist
    for total in (1, 7, 64, 65536, 1000):
        for world in (1, 2, 3, 8):
            spans = [mdist.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize('total', [10, 7])
def test_two_rank_gloo_equals_single_process(tmp_path, total):
    import torch
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    from magical_b200 import dist as mdist
    out = str(tmp_path / 'r0.npz')
    port = _free_port()
    mp.spawn(_worker, args=(2, port, total, out), nprocs=2, join=True)
    got = np.load(out)
    # single-process reference
    env = mdist.ShardedVecEnv(lambda n: FakeLocalEnv(0, n), total, 0, 1)
    obs0 = env.reset()
    assert np.array_equal(got['obs0'], obs0.numpy())
    g = torch.Generator().manual_seed(0)
    for t in range(4):
        actions = torch.randint(0, 18, (total,), generator=g, dtype=torch.int32)
        obs, rew, done, info = env.step(actions)
        want = np.concatenate([obs.numpy().reshape(total, -1).astype(np.float64), rew.numpy()[:, None],
                               done.numpy()[:, None], info['eval_score'].numpy()[:, None]], axis=1)
        assert np.array_equal(got['steps'][t], want)


class FakeStackEnv:
    """A stand-in with the real observation contract: channel-last stack of 4
    RGB frames, oldest first, refilled with the first frame of the new episode
    when an env auto-resets (the step's obs is then that refilled stack)."""
    auto_reset = True

    def __init__(self, start, n):
        import torch
        self.idx = torch.arange(start, start + n)
        self.t = 0
        self.stack = None

    def _frame(self, salt):
        import torch
        v = (self.idx * 7 + salt).to(torch.uint8)                      # [n]
        px = torch.arange(2 * 2 * 3, dtype=torch.uint8).view(1, 2, 2, 3)
        return v.view(-1, 1, 1, 1) + px                                 # [n, 2, 2, 3]

    def reset(self):
        self.t = 0
        self.stack = self._frame(0).repeat(1, 1, 1, 4)
        return self.stack.clone()

    def step(self, actions):
        import torch
        self.t += 1
        new = self._frame(self.t * 13) + actions.to(torch.uint8).view(-1, 1, 1, 1)
        self.stack = torch.cat([self.stack[..., 3:], new], dim=-1)
        done = ((self.idx + self.t) % 5 == 0)
        first = self._frame(self.t * 31 + 5)
        self.stack[done] = first[done].repeat(1, 1, 1, 4)
        rew = torch.zeros(len(self.idx))
        score = done.float() * 0.5
        return self.stack.clone(), rew, done.to(torch.uint8), {'eval_score': score}

    def close(self):
        pass


def _worker_newest(rank, world, port, total, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from magical_b200 import dist as mdist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    start, _ = mdist.shard_range(total, rank, world)
    envs = {mode: mdist.ShardedVecEnv(lambda n: FakeStackEnv(start, n), total, rank, world, gather_obs=mode)
            for mode in (True, 'newest')}
    obs = {mode: [e.reset().clone()] for mode, e in envs.items()}
    g = torch.Generator().manual_seed(3)
    for t in range(12):
        actions = torch.randint(0, 18, (total,), generator=g, dtype=torch.int32)
        for mode, e in envs.items():
            obs[mode].append(e.step(actions)[0].clone())
    same = all(torch.equal(a, b) for a, b in zip(obs[True], obs['newest']))
    if rank == 1:
        np.savez(out, same=np.array(same), last=obs['newest'][-1].numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('total', [8, 11])
def test_newest_frame_gather_rebuilds_the_same_stacks(tmp_path, total):
    """gather_obs='newest' (one frame per env on the wire) must give every rank
    the same global observation as the full all-gather, through auto-resets and
    with uneven shards; and the same as one process stepping everything."""
    import torch
    import torch.multiprocessing as mp
    out = str(tmp_path / 'r1.npz')
    mp.spawn(_worker_newest, args=(2, _free_port(), total, out), nprocs=2, join=True)
    got = np.load(out)
    assert bool(got['same'])
    env = FakeStackEnv(0, total)
    env.reset()
    g = torch.Generator().manual_seed(3)
    for t in range(12):
        obs = env.step(torch.randint(0, 18, (total,), generator=g, dtype=torch.int32))[0]
    assert np.array_equal(got['last'], obs.numpy())
