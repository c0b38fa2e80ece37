"""Env-sharded multi-GPU execution (one process per GPU).

Environments are independent, so the batch is partitioned into contiguous
shards, one per rank, and the physics/render path needs NO collective.  The
only exchange is the optional all-gather of what `step()` returns: always the
tiny per-env vectors (reward f32, done u8, score f32), and on request the
observation shard.  `torch.distributed` (NCCL over NVLink on GPUs, gloo on CPU
in tests) provides the plumbing.
"""
import numpy as np


def shard_range(total, rank, world):
    """Contiguous [start, stop) of rank's shard; sizes differ by at most 1."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def pack_scalars(reward, done, score):
    """reward f32[n], done u8[n], score f32[n] -> one f32[n, 3] send buffer so
    a single all_gather moves all three."""
    import torch
    return torch.stack([reward, done.to(torch.float32), score], dim=1)


def unpack_scalars(packed):
    import torch
    return (packed[:, 0].contiguous(), packed[:, 1].to(torch.uint8),
            packed[:, 2].contiguous())


def all_gather_shards(local, sizes, group=None):
    """All-gather per-rank shards of possibly unequal length along dim 0."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if len(set(sizes)) == 1:
        out = torch.empty((world * sizes[0],) + tuple(local.shape[1:]),
                          dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    # uneven shards: pad to the largest shard, gather, trim
    nmax = max(sizes)
    padded = torch.zeros((nmax,) + tuple(local.shape[1:]), dtype=local.dtype,
                         device=local.device)
    padded[:local.shape[0]] = local
    out = torch.empty((world * nmax,) + tuple(local.shape[1:]),
                      dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    out = out.view((world, nmax) + tuple(local.shape[1:]))
    return torch.cat([out[r, :sizes[r]] for r in range(world)], dim=0)


class ShardedVecEnv:
    """A global batch of `total` envs split over the ranks of a process group.

    `step(global_actions)` takes the GLOBAL action vector (every rank passes
    the same tensor, as a data-parallel learner would after its own
    all-gather), steps the local shard and returns
      obs:    the LOCAL observation shard (or the gathered global batch when
              gather_obs=True),
      reward/done/score: always gathered to the global batch.

    gather_obs='newest' (channel-last stacked layouts LoRes4E / LoRes4A only)
    all-gathers just the newest frame of every env (27 648 B instead of
    110 592 B per env and step) and rebuilds the 4-frame stacks of the global
    batch on each rank: shift by one frame, append the gathered frame, and
    refill the stack of every env that auto-reset in this step with its first
    frame, as `FlattenFrameStack.reset` does (benchmarks/__init__.py:130-136).
    """

    FRAME_C, DEPTH = 3, 4

    def __init__(self, make_local_env, total, rank, world, gather_obs=False,
                 group=None):
        self.total, self.rank, self.world = total, rank, world
        self.start, self.stop = shard_range(total, rank, world)
        self.sizes = [shard_range(total, r, world)[1]
                      - shard_range(total, r, world)[0] for r in range(world)]
        self.local = make_local_env(self.stop - self.start)
        self.gather_obs = gather_obs
        self.group = group

    def reset(self):
        obs = self.local.reset()
        out = self._maybe_gather_obs(obs)
        if self.gather_obs == 'newest' and self.world > 1:
            self._global = out.clone()      # full gather once per reset
            return self._global
        return out

    def _maybe_gather_obs(self, obs):
        if not self.gather_obs or self.world == 1:
            return obs
        return all_gather_shards(obs, self.sizes, self.group)

    def _gather_newest(self, obs, done_global):
        """Global stacks from the newest frames only (see the class docstring)."""
        c, depth = self.FRAME_C, self.DEPTH
        assert obs.shape[-1] == c * depth, \
            "gather_obs='newest' needs a channel-last 4-frame stack"
        newest = all_gather_shards(obs[..., c * (depth - 1):].contiguous(),
                                   self.sizes, self.group)
        g = self._global
        g[..., :c * (depth - 1)] = g[..., c:].clone()
        g[..., c * (depth - 1):] = newest
        if getattr(self.local, 'auto_reset', False):
            fresh = done_global.bool()
            if bool(fresh.any()):
                g[fresh] = newest[fresh].repeat(
                    *([1] * (newest.dim() - 1)), depth)
        return g

    def step(self, global_actions):
        local_actions = global_actions[self.start:self.stop]
        obs, rew, done, info = self.local.step(local_actions)
        if self.world > 1:
            packed = all_gather_shards(
                pack_scalars(rew, done, info['eval_score']), self.sizes,
                self.group)
            rew, done, score = unpack_scalars(packed)
            info = {'eval_score': score}
            if self.gather_obs == 'newest':
                return self._gather_newest(obs, done), rew, done, info
        return self._maybe_gather_obs(obs), rew, done, info

    def close(self):
        self.local.close()
