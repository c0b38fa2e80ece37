"""Batched, GPU-resident MAGICAL environments.

`MagicalVecEnv` is the batched counterpart of the reference's per-process
`gym.Env` stack (`BaseEnv.step/reset/render`, magical/base_env.py:177-338,
wrapped by the LoRes* preprocessors, magical/benchmarks/__init__.py:208-274):
one object steps `batch` environments of one registered env id on one GPU
through the C ABI.  PyTorch is used only to own the returned device buffers
(observation, reward, done, score).
"""
import numpy as np

from magical_b200 import _native
from magical_b200 import scene as sc

PREPROC_TO_MODE = {
    None: sc.OBS_RAW,
    'LoRes4E': sc.OBS_LORES4E,
    'LoRes4A': sc.OBS_LORES4A,
    'LoRes3EA': sc.OBS_LORES3EA,
    'LoResStack': sc.OBS_LORESSTACK,
    'LoResCHW4E': sc.OBS_LORESCHW4E,
}


def obs_shape(mode, batch, res=384):
    if mode in (sc.OBS_LORES4E, sc.OBS_LORES4A, sc.OBS_LORES3EA):
        return (batch, 96, 96, 12)
    if mode == sc.OBS_LORESSTACK:
        return (2, batch, 96, 96, 12)
    if mode == sc.OBS_LORESCHW4E:
        return (batch, 12, 96, 96)
    return (2, batch, res, res, 3)


class MagicalVecEnv:
    """`batch` environments built by one task object (`task.build_scene()`).

    n_scenes > 1 pre-samples that many scenes from the task's RandomState
    (needed for the randomised Test* variants); each reset then draws a scene
    index per environment from `self.rng`.
    """

    def __init__(self, task, batch, preproc=None, device=0, auto_reset=True,
                 n_scenes=1, seed=None, scenes=None, stream=None,
                 alloc_obs=True, keep_scene=False, default_scene_ids=None,
                 device_sampling=False):
        import torch
        if not torch.cuda.is_available():
            raise _native.NativeError(
                "MagicalVecEnv needs a CUDA device: the hot path has no CPU "
                "implementation in this package")
        self._torch = torch
        self._lib = _native.load()
        self.task = task
        self.batch = int(batch)
        self.preproc = preproc
        self.mode = PREPROC_TO_MODE[preproc]
        self.device = torch.device('cuda', device)
        self.auto_reset = bool(auto_reset)
        self.max_episode_steps = task.max_episode_steps
        self.rng = np.random.RandomState(seed)
        if seed is not None:
            task.seed(seed)
        # device_sampling: the scenes are TEMPLATES (structure + dynamics); every
        # reset re-samples goal sizes and poses on the device (csrc/mg_sample.cu)
        self.device_sampling = bool(device_sampling)
        self.programs = None
        if self.device_sampling:
            assert scenes is None and not keep_scene
            built = [task.build_template() for _ in range(n_scenes)]
            scenes = [b[0] for b in built]
            self.programs = np.ascontiguousarray(
                np.stack([b[1] for b in built])).astype(sc.placement_dt,
                                                        copy=False)
        if scenes is None:
            scenes = [task.build_scene() for _ in range(n_scenes)]
        self.scenes = np.ascontiguousarray(np.stack(scenes)).astype(
            sc.scene_dt, copy=False)
        self.n_scenes = len(self.scenes)
        self.skipped_scenes = 0
        # pool entries resets draw from, and the env-step at which that range
        # last changed (refresh_pool's safety interlock)
        self._draw = (0, 0 if keep_scene else self.n_scenes)
        self._steps = 0
        self._draw_changed_at = 0
        self._all_in_range = False  # every env known to play a draw-range entry
        # keep_scene batches (e.g. a mixed-task batch): env -> scene binding
        # applied by a full reset() that names no scene ids
        self.default_scene_ids = None if default_scene_ids is None else \
            np.ascontiguousarray(default_scene_ids, dtype=np.int32)
        res = task.res_hw[0]
        cfg = _native.make_config(device=device, batch=self.batch,
                                  n_scenes=self.n_scenes, obs_mode=self.mode,
                                  res=res, auto_reset=int(self.auto_reset),
                                  reset_seed=int(self.rng.randint(1 << 31)),
                                  keep_scene=int(bool(keep_scene)),
                                  device_sampling=int(self.device_sampling))
        import ctypes
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            self._stream = stream if stream is not None \
                else torch.cuda.current_stream()
            _native.check(self._lib.mg_create(
                cfg.ctypes.data, self.scenes.ctypes.data,
                ctypes.c_void_p(self._stream.cuda_stream),
                ctypes.byref(self._h)))
            # alloc_obs=False: physics-only use (step_physics / eval_score), or
            # the caller binds its own buffer with bind_obs(); until then
            # step() and render() fail with "no observation buffer bound"
            self.obs_shape = obs_shape(self.mode, self.batch, res)
            self.obs = torch.zeros(obs_shape(self.mode, self.batch, res),
                                   dtype=torch.uint8, device=self.device) \
                if alloc_obs else None
            self.reward = torch.zeros(self.batch, dtype=torch.float32,
                                      device=self.device)
            self.done = torch.zeros(self.batch, dtype=torch.uint8,
                                    device=self.device)
            self.score = torch.zeros(self.batch, dtype=torch.float32,
                                     device=self.device)
        if self.obs is not None:
            _native.check(self._lib.mg_bind_obs(self._h, self.obs.data_ptr(),
                                                self.obs.numel()))
        if self.device_sampling:
            _native.check(self._lib.mg_set_placement(
                self._h, 0, self.n_scenes, self.programs.ctypes.data))

    # -- gym-like batched API ---------------------------------------------
    def reset(self, env_ids=None, scene_ids=None):
        """Reset all (or the given) environments; returns the observation
        tensor (first frame replicated over the stack)."""
        ids_p = sid_p = None
        n = self.batch
        if env_ids is not None:
            env_ids = np.ascontiguousarray(env_ids, dtype=np.int32)
            n = len(env_ids)
            ids_p = env_ids.ctypes.data
        if scene_ids is None and self.default_scene_ids is not None:
            scene_ids = self.default_scene_ids if env_ids is None \
                else self.default_scene_ids[env_ids]
        if scene_ids is None and self.device_sampling:
            first, count = self._draw    # template ids, always explicit
            scene_ids = first + self.rng.randint(0, max(count, 1), size=n)
        if scene_ids is None and self.n_scenes > 1 and self._draw[1] > 0:
            # host-side draws respect the draw range, like the device-side
            # redraw of an auto-reset does
            first, count = self._draw
            scene_ids = first + self.rng.randint(0, count, size=n)
        if scene_ids is not None:
            scene_ids = np.ascontiguousarray(scene_ids, dtype=np.int32)
            assert len(scene_ids) == n
            sid_p = scene_ids.ctypes.data
        _native.check(self._lib.mg_reset(self._h, ids_p, n, sid_p))
        if env_ids is None and scene_ids is not None and self._draw[1] > 0:
            first, count = self._draw
            self._all_in_range = bool(((scene_ids >= first)
                                       & (scene_ids < first + count)).all())
        return self.obs

    # -- scene pool maintenance (randomised variants) ----------------------
    def update_scenes(self, first, scenes):
        """Overwrite pool entries [first, first + len(scenes)); no environment
        may still be playing them (keep them outside the draw range for one
        episode length first)."""
        scenes = np.ascontiguousarray(np.stack(scenes)).astype(sc.scene_dt,
                                                                copy=False)
        try:
            _native.check(self._lib.mg_update_scenes(
                self._h, int(first), len(scenes), scenes.ctypes.data))
            self.scenes[first:first + len(scenes)] = scenes
            return
        except _native.NativeError as ex:
            if 'exceeds the physics layout' not in str(ex):
                raise
        # Some scene has more bodies / shape groups than any scene of the pool
        # the handle was created with (the physics kernel's shared-memory
        # layout is fixed at mg_create): stream the others, keep the old scene
        # in those slots, and count them.  Create the handle with a pool large
        # enough to contain the task's biggest layouts to avoid this.
        for i in range(len(scenes)):
            rc = self._lib.mg_update_scenes(self._h, int(first) + i, 1,
                                            scenes[i:i + 1].ctypes.data)
            if rc == 0:
                self.scenes[first + i] = scenes[i]
            else:
                self.skipped_scenes += 1

    def set_draw_range(self, first, count):
        """Pool entries an auto-reset draws from (count 0: keep the scene)."""
        _native.check(self._lib.mg_set_draw_range(self._h, int(first),
                                                  int(count)))
        if (int(first), int(count)) != self._draw:
            self._draw_changed_at = self._steps
            self._all_in_range = False
        self._draw = (int(first), int(count))

    def refresh_pool(self, sampler=None, block=True):
        """Double-buffered pool streaming: put fresh scenes into the half of
        the pool that is currently NOT drawn from, then make it the draw range.
        The first call only narrows the draw range to the first half (no
        entry can be overwritten while environments may be playing it); later
        calls return None without doing anything until one episode length of
        steps has passed since the last switch (environments still playing
        the other half finish within one episode).  The scenes come from `sampler`
        (a `pool_sampler.ScenePoolSampler` running ahead in worker processes;
        with block=False nothing happens and None is returned if it has fewer
        than half a pool ready) or, without one, are sampled here from the
        task's RandomState.  Returns the new draw range."""
        half = self.n_scenes // 2
        assert half >= 1, 'refresh_pool needs a pool of at least 2 scenes'
        first, count = self._draw
        if count != half or first not in (0, half):
            # First call (or a custom range): environments may be playing any
            # entry, so nothing can be overwritten yet.  Only narrow the draw
            # range to the first half; the other half drains within one episode
            # and is replaced by the next call.
            self.set_draw_range(0, half)
            return 0, half
        if not self._all_in_range and \
                self._steps - self._draw_changed_at < self.max_episode_steps:
            # environments bound to the idle half before the last switch may
            # still be mid-episode on it (mg_update_scenes' contract)
            return None
        new_first = half if first == 0 else 0
        if sampler is not None:
            fresh = sampler.take(half, block=block)
            if fresh is None:
                return None
        else:
            fresh = [self.task.build_scene() for _ in range(half)]
        self.update_scenes(new_first, fresh)
        self.set_draw_range(new_first, half)
        return new_first, half

    def step(self, actions):
        """actions: int32 CUDA tensor [batch] (other int tensors / arrays are
        converted).  Returns (obs, reward, done, info) with device tensors;
        info['eval_score'] is non-zero only where done (base_env.py:275-288).
        The observation tensor is reused between calls."""
        actions = self._as_actions(actions)
        _native.check(self._lib.mg_step(
            self._h, actions.data_ptr(), self.reward.data_ptr(),
            self.done.data_ptr(), self.score.data_ptr()))
        self._steps += 1
        return self.obs, self.reward, self.done, {'eval_score': self.score}

    def step_physics(self, actions):
        """Physics + bookkeeping only (no render)."""
        actions = self._as_actions(actions)
        _native.check(self._lib.mg_step_physics(
            self._h, actions.data_ptr(), self.reward.data_ptr(),
            self.done.data_ptr(), self.score.data_ptr()))
        self._steps += 1
        return self.reward, self.done, {'eval_score': self.score}

    def step_render(self):
        """The render half of `step()`: pairs 1:1 with `step_physics()`
        (step == step_physics + step_render); pushes a frame onto the stacks."""
        _native.check(self._lib.mg_step_render(self._h))
        return self.obs

    def render(self):
        """Redraw the current state without advancing the frame stacks (the
        newest frame is replaced in place): idempotent, like the reference's
        `env.render()` (base_env.py:309-338)."""
        _native.check(self._lib.mg_render(self._h))
        return self.obs

    # -- output buffers owned by the caller (multi-GPU sharding) -----------
    def bind_obs(self, obs):
        """Render into `obs` from now on: a contiguous CUDA uint8 tensor of
        this handle's layout, e.g. this rank's slice of a global observation
        tensor.  For the two-plane layouts (LoResStack, raw) `obs` may also be
        a [2, B, ...] view whose planes are contiguous but apart."""
        assert obs.dtype == self._torch.uint8 and obs.device == self.device
        shape = obs_shape(self.mode, self.batch, self.task.res_hw[0])
        assert tuple(obs.shape) == shape, (tuple(obs.shape), shape)
        if obs.is_contiguous():
            _native.check(self._lib.mg_bind_obs(self._h, obs.data_ptr(),
                                                obs.numel()))
        else:
            assert len(shape) == 5 and obs[0].is_contiguous() \
                and obs[1].is_contiguous(), 'planes must be contiguous'
            _native.check(self._lib.mg_bind_obs_planes(
                self._h, obs[0].data_ptr(), obs[0].numel(),
                obs[1].data_ptr() - obs[0].data_ptr()))
        self.obs = obs
        return obs

    def newest_shape(self):
        """Shape of the optional newest-frame output, or None when the layout
        has none (only LoRes4E / LoRes4A / LoResStack)."""
        n = int(self._lib.mg_newest_nbytes(self._h))
        if n <= 0:
            return None
        return (n // (self.batch * 96 * 96 * 3), self.batch, 96, 96, 3)

    def bind_newest(self, newest):
        """Also write every environment's newest frame alone into `newest`
        (u8 [views, B, 96, 96, 3], CUDA, contiguous; None unbinds): the send
        buffer of the multi-GPU observation gather."""
        if newest is None:
            _native.check(self._lib.mg_bind_newest(self._h, None, 0))
        else:
            assert newest.is_contiguous() and newest.device == self.device
            _native.check(self._lib.mg_bind_newest(self._h, newest.data_ptr(),
                                                   newest.numel()))
        self._newest = newest

    def bind_scalars(self, reward, done, score):
        """Let step() write reward f32[B] / done u8[B] / eval_score f32[B] into
        caller-owned CUDA tensors (e.g. views of one packed send buffer)."""
        torch = self._torch
        for t, dt in ((reward, torch.float32), (done, torch.uint8),
                      (score, torch.float32)):
            assert t.dtype == dt and t.shape == (self.batch,) \
                and t.is_contiguous() and t.device == self.device
        self.reward, self.done, self.score = reward, done, score

    def eval_score(self):
        """score_on_end_of_traj of every env's current state."""
        out = self._torch.zeros_like(self.score)
        _native.check(self._lib.mg_score(self._h, out.data_ptr()))
        return out

    def _as_actions(self, actions):
        torch = self._torch
        if not torch.is_tensor(actions):
            actions = torch.as_tensor(np.asarray(actions, dtype=np.int32))
        if actions.dtype != torch.int32 or actions.device != self.device \
                or not actions.is_contiguous():
            actions = actions.to(device=self.device, dtype=torch.int32,
                                 non_blocking=True).contiguous()
        assert actions.shape == (self.batch,), actions.shape
        self._last_actions = actions  # keep alive until the kernels ran
        return actions

    # -- introspection (tests) --------------------------------------------
    def get_state(self, env):
        st = np.zeros((), dtype=sc.state_dt)
        _native.check(self._lib.mg_get_state(self._h, int(env),
                                             st.ctypes.data))
        return st

    def get_poses(self):
        """(x, y, angle) of every body of every environment, host float64
        [batch, MAX_BODIES, 3], in one copy."""
        out = np.zeros((self.batch, sc.MAX_BODIES, 4), dtype=np.float64)
        _native.check(self._lib.mg_get_poses(self._h, out.ctypes.data))
        return out[:, :, :3]

    def get_env_scene(self, env):
        """The compiled scene record environment `env` is playing right now
        (with device_sampling: its own sampled layout)."""
        rec = np.zeros((), dtype=sc.scene_dt)
        _native.check(self._lib.mg_get_env_scene(self._h, int(env),
                                                 rec.ctypes.data))
        return rec

    def sampler_failures(self):
        """Resets for which the device sampler found no placement within the
        reference's try / retry limits and played the template's own layout."""
        import ctypes
        out = ctypes.c_int64(0)
        _native.check(self._lib.mg_sampler_failures(self._h, ctypes.byref(out)))
        return int(out.value)

    def set_state(self, env, state):
        """Restore one environment from a `get_state` snapshot (poses,
        velocities, bias velocities, joint accumulators, contact cache,
        episode step): checkpoint / resume of the simulator state."""
        st = np.ascontiguousarray(state, dtype=sc.state_dt)
        _native.check(self._lib.mg_set_state(self._h, int(env),
                                             st.ctypes.data))

    def set_pose(self, env, body, x, y, angle):
        _native.check(self._lib.mg_set_pose(self._h, int(env), int(body),
                                            float(x), float(y), float(angle)))

    def launch_count(self):
        return int(self._lib.mg_launch_count(self._h))

    def overflow_count(self):
        """Environments x episodes that hit a physics capacity limit since
        creation (0 on every measured workload; synchronises the stream)."""
        import ctypes
        out = ctypes.c_int64(0)
        _native.check(self._lib.mg_overflow_count(self._h, ctypes.byref(out)))
        return int(out.value)

    def synchronize(self):
        _native.check(self._lib.mg_synchronize(self._h))

    def close(self):
        if getattr(self, '_h', None):
            self._lib.mg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
