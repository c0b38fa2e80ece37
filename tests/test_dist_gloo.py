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
