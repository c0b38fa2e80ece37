"""Env-sharded multi-GPU execution (one process per GPU), SURVEY 8(e).

Environments are independent, so the global batch is partitioned into
contiguous shards, one per rank, and the physics/render path needs NO
collective.  The only exchange is the all-gather of what `step()` returns:
always the per-env scalars (reward f32, done u8, eval_score f32 in ONE packed
buffer, 12 B/env), and on request the observation batch.

gather_obs='newest' is the B200 design for the observation gather.  Every rank
renders its shard straight into ITS SLICE of the global observation tensor
(`mg_bind_obs` / `mg_bind_obs_planes`) and `k_raster` additionally writes each
environment's newest frame into a send buffer (`mg_bind_newest`).  Only that
buffer goes over NVLink (27 648 B instead of 110 592 B per environment and
view); `mg_stack_push` (`k_stack_push`) then folds the received frames into
the stacks of the remote shards with `FlattenFrameStack` semantics (shift,
append, refill on auto-reset; reference benchmarks/__init__.py:118-136).
Gather and push run on a side stream; with pipeline=True `step()` returns
without waiting for them, so they overlap the next step's physics (call
`wait_obs()` before consuming the remote shards).

`torch.distributed` provides the plumbing (NCCL over NVLink on GPUs; gloo on
CPU in the host-logic tests, which inject a reference `stack_push` — the
product path itself has no CPU implementation).
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


def cuda_stack_push(stacks, newest, fresh, env_first, env_count, shard,
                    rank_stride, stream=None):
    """`mg_stack_push` on CUDA tensors: stacks u8 [n, 96, 96, 12] (one view
    plane), newest = the gathered frame buffer positioned at this plane's
    first frame, fresh u8 [n] or None."""
    import torch
    from magical_b200 import _native
    if not stacks.is_cuda:
        raise _native.NativeError(
            'mg_stack_push runs on CUDA tensors only (no CPU implementation '
            'in the product path)')
    s = stream if stream is not None else torch.cuda.current_stream()
    _native.check(_native.load().mg_stack_push(
        stacks.data_ptr(), newest.data_ptr(),
        None if fresh is None else fresh.data_ptr(), int(env_first),
        int(env_count), int(shard), int(rank_stride), int(stacks.shape[1]),
        s.cuda_stream))


class ShardedVecEnv:
    """A global batch of `total` envs split over the ranks of a process group.

    `step(global_actions)` takes the GLOBAL action vector (every rank passes
    the same tensor, as a data-parallel learner would after its own
    all-gather), steps the local shard and returns
      obs:    the LOCAL observation shard (gather_obs=False) or the global
              batch (gather_obs=True: the stacked shards are all-gathered;
              gather_obs='newest': see the module docstring),
      reward/done/score: always gathered to the global batch.

    gather_obs='newest' needs a channel-last 4-frame single- or two-plane
    layout (LoRes4E / LoRes4A / LoResStack) and equal shards.
    """

    FRAME_C, DEPTH = 3, 4
    NEWEST_PREPROCS = ('LoRes4E', 'LoRes4A', 'LoResStack')

    def __init__(self, make_local_env, total, rank, world, gather_obs=False,
                 group=None, pipeline=False, stack_push=None):
        import torch
        self._torch = torch
        self.total, self.rank, self.world = total, rank, world
        self.start, self.stop = shard_range(total, rank, world)
        self.sizes = [shard_range(total, r, world)[1]
                      - shard_range(total, r, world)[0] for r in range(world)]
        self.local = make_local_env(self.stop - self.start)
        self.gather_obs = gather_obs
        self.group = group
        self.pipeline = bool(pipeline)
        self._stack_push = stack_push or cuda_stack_push
        self._t = 0
        self._newest_ready = False
        if gather_obs == 'newest' and world > 1:
            self._setup_newest()

    # ------------------------------------------------------------- set-up
    def _setup_newest(self):
        torch = self._torch
        local = self.local
        preproc = getattr(local, 'preproc', None)
        if preproc not in self.NEWEST_PREPROCS:
            raise ValueError(
                "gather_obs='newest' needs a channel-last 4-frame stack per "
                f"view (one of {self.NEWEST_PREPROCS}); the local env has "
                f"preproc {preproc!r} (LoRes3EA mixes views inside one "
                "pixel, LoResCHW4E is planar: use gather_obs=True)")
        if len(set(self.sizes)) != 1:
            raise ValueError("gather_obs='newest' needs equal shards "
                             f"(total {self.total} over {self.world} ranks)")
        n = self.sizes[0]
        self.views = 2 if preproc == 'LoResStack' else 1
        dev = torch.device(local.device)
        self._cuda = dev.type == 'cuda'
        H, W = tuple(local.obs_shape)[-3:-1]
        gshape = (self.total, H, W, 12)
        if self.views == 2:
            gshape = (2,) + gshape
        # the global observation batch; this rank's shard is rendered in place
        self._global = torch.zeros(gshape, dtype=torch.uint8, device=dev)
        sl = (slice(self.start, self.stop),)
        self._local_view = self._global[(slice(None),) + sl] \
            if self.views == 2 else self._global[sl]
        self._bound = hasattr(local, 'bind_obs')
        if self._bound:
            local.bind_obs(self._local_view)
        fshape = (self.views, n, H, W, 3)
        self._send = [torch.zeros(fshape, dtype=torch.uint8, device=dev)
                      for _ in range(2)]
        self._recv = torch.zeros((self.world,) + fshape, dtype=torch.uint8,
                                 device=dev)
        self._frame_bytes = H * W * 3
        # packed scalars: [reward f32 n | score f32 n | done u8 n | pad 3n]
        # (rows of 12n bytes keep the f32 views of the gathered rows aligned)
        self._sc_send = torch.zeros(12 * n, dtype=torch.uint8, device=dev)
        self._sc_recv = torch.zeros((self.world, 12 * n), dtype=torch.uint8,
                                    device=dev)
        self._sc_views = (self._sc_send[:4 * n].view(torch.float32),
                          self._sc_send[8 * n:9 * n],
                          self._sc_send[4 * n:8 * n].view(torch.float32))
        self._sc_bound = hasattr(local, 'bind_scalars')
        if self._sc_bound:
            local.bind_scalars(*self._sc_views)
        self._fresh = torch.zeros(self.total, dtype=torch.uint8, device=dev)
        if self._cuda:
            self._comm = torch.cuda.Stream(device=dev)
            self._ev_step = torch.cuda.Event()
            self._ev_ready = torch.cuda.Event()
            self._ev_send_free = [torch.cuda.Event(), torch.cuda.Event()]
            self._send_used = [False, False]
        self._newest_ready = True

    # -------------------------------------------------------------- reset
    def reset(self):
        obs = self.local.reset()
        if self._newest_ready:
            if not self._bound:
                self._local_view.copy_(obs)
            self._full_gather()
            return self._global
        return self._maybe_gather_obs(obs)

    def _full_gather(self):
        """All-gather the stacked shards once (reset)."""
        import torch.distributed as dist
        g = self._global
        planes = [g[v] for v in range(2)] if self.views == 2 else [g]
        for p in planes:
            dist.all_gather_into_tensor(
                p, p[self.start:self.stop].contiguous(), group=self.group)

    def _maybe_gather_obs(self, obs):
        if not self.gather_obs or self.world == 1:
            return obs
        if obs.dim() == 5:   # two-plane layouts [2, n, H, W, C]: envs are dim 1
            return self._torch.stack([
                all_gather_shards(obs[v], self.sizes, self.group)
                for v in range(obs.shape[0])])
        return all_gather_shards(obs, self.sizes, self.group)

    # --------------------------------------------------------------- step
    def _exchange(self, send):
        """All-gather the packed scalars and the newest frames, then fold the
        frames into the remote shards' stacks (current stream)."""
        import torch.distributed as dist
        torch = self._torch
        n = self.sizes[0]
        dist.all_gather_into_tensor(self._sc_recv.view(-1), self._sc_send,
                                    group=self.group)
        dist.all_gather_into_tensor(self._recv.view(-1), send.view(-1),
                                    group=self.group)
        rew = self._sc_recv[:, :4 * n].view(torch.float32).reshape(-1)
        score = self._sc_recv[:, 4 * n:8 * n].view(torch.float32).reshape(-1)
        done = self._sc_recv[:, 8 * n:9 * n].reshape(-1)
        fresh = None
        if getattr(self.local, 'auto_reset', False):
            fresh = done
        rank_stride = self.views * n * self._frame_bytes
        for v in range(self.views):
            stacks = self._global[v] if self.views == 2 else self._global
            newest = self._recv.view(-1)[v * n * self._frame_bytes:]
            for first, count in ((0, self.start),
                                 (self.stop, self.total - self.stop)):
                if count > 0:
                    self._stack_push(stacks, newest, fresh, first, count, n,
                                     rank_stride)
        return rew, done, score

    def _step_newest(self, local_actions):
        torch = self._torch
        local = self.local
        k = self._t & 1
        self._t += 1
        send = self._send[k]
        if self._cuda:
            cur = torch.cuda.current_stream()
            if self._send_used[k]:
                cur.wait_event(self._ev_send_free[k])  # gather of step t-2
        if hasattr(local, 'bind_newest'):
            local.bind_newest(send)
        obs, rew, done, info = local.step(local_actions)
        if not self._bound:
            self._local_view.copy_(obs)
        if not hasattr(local, 'bind_newest'):
            if self.views == 2:
                send.copy_(obs[..., 9:])
            else:
                send[0].copy_(obs[..., 9:])
        if not self._sc_bound:
            r, d, s = self._sc_views
            r.copy_(rew)
            d.copy_(done)
            s.copy_(info['eval_score'])
        if not self._cuda:
            rew, done, score = self._exchange(send)
            return self._global, rew, done, {'eval_score': score}
        self._ev_step.record(cur)
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(self._ev_step)
            rew, done, score = self._exchange(send)
            self._ev_send_free[k].record(self._comm)
            self._send_used[k] = True
            self._ev_ready.record(self._comm)
        if not self.pipeline:
            cur.wait_event(self._ev_ready)
        return self._global, rew, done, {'eval_score': score}

    def wait_obs(self):
        """pipeline=True: make the current stream wait until the gather and
        the stack push of the last step() are complete."""
        if self._newest_ready and self._cuda and self._t > 0:
            self._torch.cuda.current_stream().wait_event(self._ev_ready)

    def step(self, global_actions):
        local_actions = global_actions[self.start:self.stop]
        if self._newest_ready:
            return self._step_newest(local_actions)
        obs, rew, done, info = self.local.step(local_actions)
        if self.world > 1:
            packed = all_gather_shards(
                pack_scalars(rew, done, info['eval_score']), self.sizes,
                self.group)
            rew, done, score = unpack_scalars(packed)
            info = {'eval_score': score}
        return self._maybe_gather_obs(obs), rew, done, info

    def nvlink_bytes_per_step(self):
        """Bytes this rank RECEIVES per step (frames + scalars of the other
        ranks) in the 'newest' mode."""
        n = self.sizes[0]
        return (self.world - 1) * n * (self.views * self._frame_bytes + 12)

    def close(self):
        if self._newest_ready and getattr(self, '_cuda', False):
            self._torch.cuda.synchronize()
        self.local.close()
