"""`make()` / `make_vec()` and the batch-of-1 gym-style adaptor.

`MagicalEnv` keeps the call surface of the reference's environments as seen
through `gym.make` (reset() -> obs, step(a) -> (obs, rew, done, info),
seed(), close(), render('rgb_array'), action_space, observation_space,
max_episode_steps, fps; magical/base_env.py:97-140,177-343 and
tests/test_rollout_preproc.py:17-36) on top of a `MagicalVecEnv` of batch 1.
"""
import collections

import numpy as np

from magical_b200 import benchmarks, gymshim
from magical_b200 import entities as en
from magical_b200 import scene as sc
from magical_b200.vec_env import MagicalVecEnv, PREPROC_TO_MODE


def _spec(env_id):
    benchmarks.register_envs()
    try:
        return benchmarks.ENV_SPECS[env_id]
    except KeyError:
        raise KeyError(f"no registered MAGICAL env with id '{env_id}'")


def make_task(env_id):
    """Instantiate the task (scene factory) behind a registered env id."""
    spec = _spec(env_id)
    return spec.entry_point(**spec.kwargs), spec


def make_vec(env_id, batch, device=0, auto_reset=True, n_scenes=None,
             seed=None, stream=None, alloc_obs=True, keep_scene=False,
             device_sampling=False):
    """Batched GPU env for a registered id.  Demo variants share one scene;
    randomised Test variants pre-sample `n_scenes` scenes (default 64).
    device_sampling=True (randomised variants): the n_scenes scenes are
    structure templates and every reset draws fresh goal sizes and poses on
    the GPU.
    alloc_obs=False leaves the observation buffer to the caller (`bind_obs`,
    e.g. a slice of a multi-GPU global batch); keep_scene=True restarts every
    env on the scene it is bound to instead of redrawing from the pool."""
    task, spec = make_task(env_id)
    if n_scenes is None:
        n_scenes = 64 if benchmarks.EnvName(env_id).is_test else 1
    return MagicalVecEnv(task, batch, preproc=spec.preproc, device=device,
                         auto_reset=auto_reset, n_scenes=n_scenes, seed=seed,
                         stream=stream, alloc_obs=alloc_obs,
                         keep_scene=keep_scene,
                         device_sampling=device_sampling)


def make_vec_mixed(env_ids, batch, device=0, auto_reset=True, stream=None,
                   alloc_obs=True, first_env=0, total=None):
    """One batch holding several registered ids with the same preprocessor
    (BASELINE config 5: all 8 Demo tasks in one batch).  The batch is cut into
    len(env_ids) contiguous groups, group k plays env_ids[k] (contiguous so
    that the warps of the physics kernel see one scene structure each); every
    env restarts on its own task at auto-reset.  `first_env` / `total` place
    this batch inside a larger, sharded global batch whose groups are defined
    over `total` envs."""
    made = [make_task(e) for e in env_ids]
    preprocs = {spec.preproc for _, spec in made}
    assert len(preprocs) == 1, f'one preprocessor per batch, got {preprocs}'
    scenes = [task.build_scene() for task, _ in made]
    total = batch if total is None else total
    gids = (np.arange(first_env, first_env + batch) * len(env_ids)) // total
    venv = MagicalVecEnv(made[0][0], batch, preproc=preprocs.pop(),
                         device=device, auto_reset=auto_reset, scenes=scenes,
                         stream=stream, alloc_obs=alloc_obs, keep_scene=True,
                         default_scene_ids=gids)
    venv.max_episode_steps = max(spec.max_episode_steps for _, spec in made)
    venv.env_ids = list(env_ids)
    return venv


def make(env_id, device=0):
    """Single environment with the reference's gym call surface."""
    return MagicalEnv(env_id, device=device)


class MagicalEnv(gymshim.Env):
    def __init__(self, env_id, device=0):
        self.task, self.spec = make_task(env_id)
        self.env_id = env_id
        self._device = device
        self._venv = None
        self._raw = self._raw_scene = None
        self._randomised = benchmarks.EnvName(env_id).is_test
        self.max_episode_steps = self.spec.max_episode_steps
        self.fps = self.task.fps
        self.action_space = gymshim.Discrete(len(en.ACTION_NUMS_FLAGS_NAMES))
        mode = PREPROC_TO_MODE[self.spec.preproc]
        self._mode = mode
        h, w = self.task.res_hw

        def box(shape):
            return gymshim.Box(low=0, high=255, shape=shape, dtype=np.uint8)

        if mode == sc.OBS_RAW:
            self.observation_space = gymshim.Dict(collections.OrderedDict(
                [('allo', box((h, w, 3))), ('ego', box((h, w, 3)))]))
        elif mode == sc.OBS_LORESSTACK:
            self.observation_space = gymshim.Dict(collections.OrderedDict(
                [('allo', box((96, 96, 12))), ('ego', box((96, 96, 12)))]))
        elif mode == sc.OBS_LORESCHW4E:
            self.observation_space = box((12, 96, 96))
        else:
            self.observation_space = box((96, 96, 12))

    # reference helpers (base_env.py:124-140)
    def action_to_flags(self, int_action):
        return self.task.action_to_flags(int_action)

    def flags_to_action(self, flags):
        return self.task.flags_to_action(flags)

    def seed(self, seed=None):
        return self.task.seed(seed)

    def _host_obs(self):
        obs = self._venv.obs.cpu().numpy()
        if self._mode in (sc.OBS_RAW, sc.OBS_LORESSTACK):
            return collections.OrderedDict([('allo', obs[0, 0]),
                                            ('ego', obs[1, 0])])
        return obs[0]

    def reset(self):
        if self._venv is None or self._randomised:
            if self._venv is not None:
                self._venv.close()
            self._venv = MagicalVecEnv(self.task, 1, preproc=self.spec.preproc,
                                       device=self._device, auto_reset=False)
        self._venv.reset()
        return self._host_obs()

    def step(self, action):
        assert self._venv is not None, 'call reset() before step()'
        if not self.action_space.contains(action):
            raise ValueError(f'invalid action {action!r}')
        _, rew, done, info = self._venv.step(
            np.asarray([int(action)], dtype=np.int32))
        return (self._host_obs(), float(rew[0].item()),
                bool(done[0].item()),
                {'eval_score': float(info['eval_score'][0].item())})

    def render(self, mode='rgb_array'):
        """Full-resolution views of the current state, as `BaseEnv.render`
        returns them under every preprocessor (base_env.py:309-338):
        OrderedDict(allo, ego) of (H, W, 3) u8 frames."""
        if mode != 'rgb_array':
            raise NotImplementedError(
                "only mode='rgb_array' exists on the GPU path (no window)")
        if self._venv is None:
            return None
        if self._mode == sc.OBS_RAW:
            self._venv.render()
            return self._host_obs()
        # preprocessed ids: rasterise the same poses through a raw-mode handle
        st = self._venv.get_state(0)
        scene = self._venv.scenes[int(st['scene'])]
        if self._raw is None or self._raw_scene is not self._venv:
            if self._raw is not None:
                self._raw.close()
            self._raw = MagicalVecEnv(self.task, 1, preproc=None,
                                      device=self._device, auto_reset=False,
                                      scenes=[scene])
            self._raw.reset()
            self._raw_scene = self._venv
        for b in range(int(st['n_bodies'])):
            self._raw.set_pose(0, b, float(st['pos'][b][0]),
                               float(st['pos'][b][1]), float(st['angle'][b]))
        obs = self._raw.render().cpu().numpy()
        return collections.OrderedDict([('allo', obs[0, 0]),
                                        ('ego', obs[1, 0])])

    def close(self):
        for v in (self._venv, self._raw):
            if v is not None:
                v.close()
        self._venv = self._raw = self._raw_scene = None
