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


class _DevMem:
    """A raw device pointer presented through __cuda_array_interface__ so torch
    can wrap it without copying."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {
            'shape': (int(nbytes),), 'typestr': '|u1', 'data': (int(ptr), False),
            'version': 2}


class PeerTransport:
    """NVLink peer-memory transport (`mg_comm_*`, csrc/mg_comm.cu): every
    rank's send buffers live in a region its peers map through CUDA IPC, a
    flag barrier replaces the collective's synchronisation, and the stack
    rebuild kernel reads the frames straight from the owners' buffers."""

    N_BUF = 2

    def __init__(self, rank, world, n_local, views, hw, device, group=None):
        import ctypes
        import torch
        import torch.distributed as dist
        from magical_b200 import _native
        self._lib = _native.load()
        self._check = _native.check
        self.rank, self.world, self.n, self.views = rank, world, n_local, views
        self.hw = hw
        self.frame_bytes = hw[0] * hw[1] * 3
        self.scalar_bytes = 12 * n_local + (-12 * n_local) % 16
        fbytes = views * n_local * self.frame_bytes
        assert fbytes % 16 == 0
        self._h = ctypes.c_void_p()
        with torch.cuda.device(device):
            self._check(self._lib.mg_comm_create(rank, world, self.N_BUF, self.scalar_bytes, fbytes,
                                                 ctypes.byref(self._h)))
            mine = (ctypes.c_uint8 * _native.COMM_HANDLE_BYTES)()
            self._check(self._lib.mg_comm_export(self._h, mine))
            send = torch.tensor(list(mine), dtype=torch.uint8, device=device)
            allh = torch.empty(world * _native.COMM_HANDLE_BYTES, dtype=torch.uint8, device=device)
            dist.all_gather_into_tensor(allh, send, group=group)
            host = allh.cpu().numpy().tobytes()
            self._check(self._lib.mg_comm_connect(self._h, host))
            dist.barrier(group=group)     # every rank has mapped every region before anyone signals
            self._frames = [torch.as_tensor(_DevMem(self._lib.mg_comm_frame_ptr(self._h, k), fbytes),
                                            device=device).view(views, n_local, hw[0], hw[1], 3)
                            for k in range(self.N_BUF)]
            self._scalars = [torch.as_tensor(_DevMem(self._lib.mg_comm_scalar_ptr(self._h, k), self.scalar_bytes),
                                             device=device) for k in range(self.N_BUF)]
        self._group = group

    def frames(self, k):
        return self._frames[k]

    def scalars(self, k):
        return self._scalars[k]

    @staticmethod
    def _stream():
        import torch
        return torch.cuda.current_stream().cuda_stream

    def barrier(self):
        self._check(self._lib.mg_comm_barrier(self._h, self._stream()))

    def gather_scalars(self, k, dst):
        assert dst.is_contiguous() and dst.numel() == self.world * self.scalar_bytes
        self._check(self._lib.mg_comm_gather_scalars(self._h, k, dst.data_ptr(), self._stream()))

    def stack_push(self, k, view, stacks, fresh, first, count, modulo=0):
        self._check(self._lib.mg_comm_stack_push(
            self._h, k, view * self.n * self.frame_bytes, stacks.data_ptr(),
            None if fresh is None else fresh.data_ptr(), int(first), int(count), int(modulo), self.n,
            self.hw[0], self._stream()))

    def error(self):
        import ctypes
        out = ctypes.c_int32(0)
        self._check(self._lib.mg_comm_error(self._h, ctypes.byref(out)))
        return int(out.value)

    def close(self):
        import torch.distributed as dist
        if self._h:
            self._frames = self._scalars = None
            dist.barrier(group=self._group)   # nobody reads a region that is about to be freed
            self._lib.mg_comm_destroy(self._h)
            self._h = None


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
                 group=None, pipeline=False, stack_push=None, transport='auto'):
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
        # 'p2p': NVLink peer memory (PeerTransport, gather fused into the stack
        # rebuild kernel); 'nccl': ncclAllGather into a receive buffer +
        # mg_stack_push; 'auto': p2p on CUDA, the collective elsewhere (gloo)
        assert transport in ('auto', 'p2p', 'nccl')
        self.transport = transport
        self._peer = None
        self._t = 0
        self.push_launches = 0   # k_stack_push launches so far (bench bookkeeping)
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
        self._frame_bytes = H * W * 3
        if self.transport == 'auto':
            self.transport = 'p2p' if self._cuda else 'nccl'
        if self.transport == 'p2p':
            self._peer = PeerTransport(self.rank, self.world, n, self.views, (H, W), dev, self.group)
            self._send = [self._peer.frames(k) for k in range(2)]
            self._sc_send = [self._peer.scalars(k) for k in range(2)]
            row = self._peer.scalar_bytes
        else:
            self._send = [torch.zeros(fshape, dtype=torch.uint8, device=dev) for _ in range(2)]
            self._recv = torch.zeros((self.world,) + fshape, dtype=torch.uint8, device=dev)
            row = 12 * n
            self._sc_send = [torch.zeros(row, dtype=torch.uint8, device=dev) for _ in range(2)]
        # packed scalars: [reward f32 n | score f32 n | done u8 n | pad]
        # (rows of >= 12n bytes keep the f32 views of the gathered rows aligned)
        self._sc_recv = torch.zeros((self.world, row), dtype=torch.uint8, device=dev)
        self._sc_views = [(b[:4 * n].view(torch.float32), b[8 * n:9 * n], b[4 * n:8 * n].view(torch.float32))
                          for b in self._sc_send]
        self._sc_bound = hasattr(local, 'bind_scalars')
        self._fresh = torch.zeros(self.total, dtype=torch.uint8, device=dev)
        if self._cuda:
            # high priority: the exchange is enqueued AFTER the next step's
            # k_physics_tpe, whose queued blocks would otherwise all be placed
            # first.  That kernel is bound by shared memory (4 one-warp blocks
            # per SM); the stack-push blocks use no shared memory and fill the
            # registers it leaves free, so both run at full residency.
            import os
            prio = int(os.environ.get('MG_COMM_PRIORITY', '-1'))
            self._comm = torch.cuda.Stream(device=dev, priority=prio)
            self._ev_step = torch.cuda.Event()
            self._ev_ready = torch.cuda.Event()
            self._ev_send_free = [torch.cuda.Event(), torch.cuda.Event()]
            self._ev_barrier = [torch.cuda.Event(), torch.cuda.Event()]
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
    def gather_only(self, k=0):
        """The exchange of a step without the stack rebuild, on the current
        stream: p2p = flag barrier + the peers' packed scalars; nccl = the two
        all-gathers (packed scalars, newest frames)."""
        import torch.distributed as dist
        if self._peer is not None:
            self._peer.barrier()
            if self._cuda:
                self._ev_barrier[k].record(self._torch.cuda.current_stream())
            self._peer.gather_scalars(k, self._sc_recv)
            return
        dist.all_gather_into_tensor(self._sc_recv.view(-1), self._sc_send[k], group=self.group)
        dist.all_gather_into_tensor(self._recv.view(-1), self._send[k].view(-1), group=self.group)

    def push_only(self, fresh=None, k=0):
        """The stack rebuild of the remote shards on the current stream: p2p =
        k_stack_push_p2p reading the owners' buffers over NVLink; nccl =
        k_stack_push from the receive buffer."""
        n = self.sizes[0]
        rank_stride = self.views * n * self._frame_bytes
        for v in range(self.views):
            stacks = self._global[v] if self.views == 2 else self._global
            if self._peer is not None:
                # one launch over all remote shards, starting at the NEXT rank and
                # wrapping around: every rank reads a different owner at any time
                self._peer.stack_push(k, v, stacks, fresh, self.stop % self.total,
                                      self.total - n, modulo=self.total)
                self.push_launches += 1
                continue
            for first, count in ((0, self.start), (self.stop, self.total - self.stop)):
                if count <= 0:
                    continue
                newest = self._recv.view(-1)[v * n * self._frame_bytes:]
                self._stack_push(stacks, newest, fresh, first, count, n, rank_stride)
                self.push_launches += 1

    def _exchange(self, k):
        """Exchange the packed scalars and the newest frames of send buffer k
        and fold the frames into the remote shards' stacks (current stream)."""
        torch = self._torch
        n = self.sizes[0]
        self.gather_only(k)
        rew = self._sc_recv[:, :4 * n].view(torch.float32).reshape(-1)
        score = self._sc_recv[:, 4 * n:8 * n].view(torch.float32).reshape(-1)
        done = self._sc_recv[:, 8 * n:9 * n].reshape(-1)
        fresh = None
        if getattr(self.local, 'auto_reset', False):
            fresh = done
        self.push_only(fresh, k)
        return rew, done, score

    def _step_newest(self, local_actions):
        torch = self._torch
        local = self.local
        k = self._t & 1
        t = self._t
        self._t += 1
        send = self._send[k]
        if self._cuda:
            cur = torch.cuda.current_stream()
            if t >= 2:
                # send buffer k was last used by step t-2.  nccl: its gather has
                # finished locally; p2p: the PEERS have finished reading it once
                # this rank has passed the barrier of step t-1 (they enter that
                # barrier only after their stack push of step t-2)
                cur.wait_event(self._ev_barrier[1 - k] if self._peer is not None
                               else self._ev_send_free[k])
        if hasattr(local, 'bind_newest'):
            local.bind_newest(send)
        if self._sc_bound:
            local.bind_scalars(*self._sc_views[k])
        obs, rew, done, info = local.step(local_actions)
        if not self._bound:
            self._local_view.copy_(obs)
        if not hasattr(local, 'bind_newest'):
            if self.views == 2:
                send.copy_(obs[..., 9:])
            else:
                send[0].copy_(obs[..., 9:])
        if not self._sc_bound:
            r, d, s = self._sc_views[k]
            r.copy_(rew)
            d.copy_(done)
            s.copy_(info['eval_score'])
        if not self._cuda:
            rew, done, score = self._exchange(k)
            return self._global, rew, done, {'eval_score': score}
        self._ev_step.record(cur)
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(self._ev_step)
            rew, done, score = self._exchange(k)
            self._ev_send_free[k].record(self._comm)
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
        if self._peer is not None:
            assert self._peer.error() == 0, 'peer barrier timed out (a rank died?)'
            self._peer.close()
            self._peer = None
        self.local.close()
