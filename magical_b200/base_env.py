"""Task base class: world constants, reset-time scene construction.

`BaseEnv` here carries what the reference's `BaseEnv`
(magical/base_env.py:60-234) does *up to the native boundary*: the world
constants, the seeding contract, the physics-variable sampling and the
`reset()` sequence (arena first, then the task's `on_reset()` entities).
Instead of building a live pymunk space it emits a compiled scene record; the
per-step work (`step()`/`render()`, base_env.py:236-338) is done on the GPU
by `magical_b200.vec_env.MagicalVecEnv` for whole batches of these scenes.
"""
import abc
import math

import numpy as np

from magical_b200 import entities as en
from magical_b200 import scene as sc
from magical_b200.phys_vars import PhysicsVariables

__all__ = ['BaseEnv', 'PhysicsVariables']


class BaseEnv(abc.ABC):
    # constants for all envs (reference base_env.py:61-76)
    ROBOT_RAD = 0.2
    ROBOT_MASS = 1.0
    SHAPE_RAD = ROBOT_RAD * 0.6
    ARENA_BOUNDS_LRBT = [-1, 1, -1, 1]
    ARENA_SIZE_MAX = max(ARENA_BOUNDS_LRBT)
    RAND_GOAL_MIN_SIZE = 0.5
    RAND_GOAL_MAX_SIZE = 0.8
    RAND_GOAL_SIZE_RANGE = RAND_GOAL_MAX_SIZE - RAND_GOAL_MIN_SIZE
    JITTER_PCT = 0.05
    JITTER_POS_BOUND = ARENA_SIZE_MAX * JITTER_PCT / 2.0
    JITTER_ROT_BOUND = JITTER_PCT * math.pi
    JITTER_TARGET_BOUND = JITTER_PCT * RAND_GOAL_SIZE_RANGE / 2

    TASK_ID = None  # MG_TASK_* of the subclass

    def __init__(self, *, res_hw=(256, 256), fps=20, phys_steps=10,
                 phys_iter=10, max_episode_steps=None, rand_dynamics=False,
                 ego_view=True, allo_view=True):
        # the native step hard-codes what the registry always passes
        # (benchmarks/__init__.py:401-404; base_env.py:237 ignores
        # phys_steps anyway): 8 fps, 10 sub-steps, 10 solver iterations
        if fps != 8 or phys_iter != 10:
            raise ValueError(
                "the B200 path implements the registered configuration only "
                f"(fps=8, phys_iter=10); got fps={fps}, phys_iter={phys_iter}")
        self.phys_iter = phys_iter
        self.phys_steps = phys_steps
        self.fps = fps
        self.res_hw = tuple(res_hw)
        self.max_episode_steps = max_episode_steps
        self.ego_view = ego_view
        self.allo_view = allo_view
        assert self.ego_view or self.allo_view, \
            "must use egocentric view or allocentric view (or both)"
        self.rand_dynamics = rand_dynamics
        self._builder = None
        self._robot = None
        self._entities = None
        self.seed()

    # -- action helpers (base_env.py:124-131) -----------------------------
    def action_to_flags(self, int_action):
        return en.ACTION_ID_TO_FLAGS[int(int_action)]

    def flags_to_action(self, flags):
        return en.FLAGS_TO_ACTION_ID[tuple(flags)]

    def seed(self, seed=None):
        """Same contract as reference base_env.py:133-140."""
        if seed is None:
            seed = np.random.randint(0, (1 << 31) - 1)
        self.rng = np.random.RandomState(seed=seed)
        return [seed]

    # -- scene construction -----------------------------------------------
    def _make_robot(self, init_pos, init_angle):
        return en.Robot(radius=self.ROBOT_RAD, init_pos=init_pos,
                        init_angle=init_angle, mass=self.ROBOT_MASS)

    def _make_shape(self, **kwargs):
        return en.Shape(shape_size=self.SHAPE_RAD, **kwargs)

    @abc.abstractmethod
    def on_reset(self):
        """Create the task's entities with add_entities(); must add exactly
        one robot."""

    def add_entities(self, entities):
        for entity in entities:
            if isinstance(entity, en.Robot):
                self._robot = entity
            if isinstance(entity, en.GoalRegion):
                self._goal_entities.append(entity)
            self._entities.append(entity)
            self._builder.entities.append(entity)
            entity.setup(self._builder)

    def build_template(self):
        """A compiled scene plus its PLACEMENT PROGRAM: what `on_reset()` asked
        the rejection sampler to randomise (goal sizes, entity poses with
        their limits), recorded while the scene is built, so that the device
        can re-sample sizes and poses for this structure at every reset
        (`mg_set_placement`, csrc/mg_sample.cu).  Raises NotImplementedError
        for tasks whose reset uses sampler features the device kernel does not
        implement (FindDupe / FixColour: `ignore_ents`, follow-up placements
        relative to another entity)."""
        self._placement_log = {'hw': [], 'ents': None, 'unsupported': None}
        try:
            record = self.build_scene()
            log = self._placement_log
        finally:
            self._placement_log = None
        if log['unsupported']:
            raise NotImplementedError(
                f'{type(self).__name__}: device-side layout sampling does not '
                f"cover this task's reset ({log['unsupported']})")
        prog = np.zeros((), dtype=sc.placement_dt)
        prog['arena'] = self.ARENA_BOUNDS_LRBT
        prog['goal_prims'] = -1
        n_goals = int(record['n_goals'])
        for ent in self._goal_entities:
            prog['goal_prims'][ent.goal_index] = ent.prim_ids
        assert len(log['hw']) <= n_goals, 'more size draws than goal regions'
        prog['n_hw'] = len(log['hw'])
        for k, (mn, mx, cur, linf) in enumerate(log['hw']):
            # the k-th (h, w) draw sizes the k-th goal region created
            h = prog['hw'][k]
            h['goal'] = k
            h['min_side'], h['max_side'] = mn, mx
            h['cur_h'], h['cur_w'] = cur if cur is not None else (0.0, 0.0)
            h['linf'] = -1.0 if linf is None else linf
        ents = log['ents'] or []
        assert len(ents) <= sc.MAX_PLACE_ENTS
        prog['n_ents'] = len(ents)
        for k, e in enumerate(ents):
            pe = prog['ents'][k]
            pe['rand_pos'], pe['rand_rot'] = int(e['rand_pos']), int(e['rand_rot'])
            pe['pos_limit'] = -1.0 if e['pos_limit'] is None else e['pos_limit']
            pe['rot_limit'] = -1.0 if e['rot_limit'] is None else e['rot_limit']
            pe['orig'] = e['orig']
            if e['goal'] is not None:
                pe['kind'], pe['goal'] = 1, e['goal']
            else:
                assert len(e['bodies']) <= sc.MAX_PLACE_BODIES and len(e['groups']) <= 4
                pe['kind'] = 0
                pe['n_bodies'], pe['n_groups'] = len(e['bodies']), len(e['groups'])
                pe['bodies'][:len(e['bodies'])] = e['bodies']
                pe['groups'][:len(e['groups'])] = e['groups']
        return record, prog

    def build_scene(self):
        """The scene-construction half of reference `reset()`
        (base_env.py:177-223): returns one compiled scene record."""
        self._entities = []
        self._goal_entities = []
        self._robot = None
        if self.rand_dynamics:
            phys_vars = PhysicsVariables.sample(self.rng)
        else:
            phys_vars = PhysicsVariables.defaults()
        self._builder = sc.SceneBuilder(
            task=self.TASK_ID,
            max_steps=(self.max_episode_steps
                       if self.max_episode_steps is not None else 0),
            phys_vars=phys_vars,
            debug_reward=getattr(self, 'debug_reward', False))
        arena_l, arena_r, arena_b, arena_t = self.ARENA_BOUNDS_LRBT
        self._arena = en.ArenaBoundaries(left=arena_l, right=arena_r,
                                         bottom=arena_b, top=arena_t)
        self.add_entities([self._arena])
        reset_rv = self.on_reset()
        assert reset_rv is None
        assert isinstance(self._robot, en.Robot)
        self.finalise_scene(self._builder)
        record = self._builder.compile()
        self._builder = None
        return record

    def finalise_scene(self, builder):
        """Hook: write task-specific score metadata (roles, labels, expected
        blocks) into the builder after all entities exist."""

    # -- reset-time randomisation helpers (reference geom.py:116-384) ------
    def randomise_hw(self, min_side, max_side, current_hw=None,
                     linf_bound=None):
        """geom.randomise_hw (geom.py:344-359): one rng.uniform call that
        draws (h, w) together."""
        minima = np.asarray((min_side, min_side), dtype='float64')
        maxima = np.asarray((max_side, max_side), dtype='float64')
        if linf_bound is not None:
            current_hw = np.asarray(current_hw, dtype='float64')
            minima = np.maximum(minima, current_hw - linf_bound)
            maxima = np.minimum(maxima, current_hw + linf_bound)
        log = getattr(self, '_placement_log', None)
        if log is not None:
            log['hw'].append((float(min_side), float(max_side),
                              None if current_hw is None
                              else (float(current_hw[0]), float(current_hw[1])),
                              None if linf_bound is None else float(linf_bound)))
        h, w = self.rng.uniform(minima, maxima)
        return float(h), float(w)

    def randomise_all_poses(self, entities, **kwargs):
        from magical_b200 import placement
        log = getattr(self, '_placement_log', None)
        if log is not None:
            self._log_placement(log, list(entities), **kwargs)
        placement.randomise_all_poses(self._builder, entities,
                                      self.ARENA_BOUNDS_LRBT, self.rng,
                                      **kwargs)

    def _log_placement(self, log, entities, rand_pos=True, rand_rot=True,
                       rel_pos_linf_limits=None, rel_rot_limits=None,
                       ignore_ents=None, max_retries=10):
        from magical_b200 import placement
        if log['ents'] is not None:
            log['unsupported'] = 'more than one randomise_all_poses call'
            return
        if ignore_ents:
            log['unsupported'] = 'ignore_ents'
            return
        n = len(entities)
        lists = [placement._listify(v, n) for v in
                 (rand_pos, rand_rot, rel_pos_linf_limits, rel_rot_limits)]
        out = []
        for ent, rp, rr, pl, rl in zip(entities, *lists):
            if isinstance(ent, en.GoalRegion):
                out.append(dict(goal=ent.goal_index, bodies=[], groups=[],
                                orig=(ent.x, ent.y, 0.0), rand_pos=rp,
                                rand_rot=rr, pos_limit=pl, rot_limit=rl))
            else:
                (pos, ang) = placement.entity_pose(self._builder, ent)
                out.append(dict(goal=None, bodies=list(ent.body_ids),
                                groups=list(ent.cgroup_ids),
                                orig=(pos[0], pos[1], ang), rand_pos=rp,
                                rand_rot=rr, pos_limit=pl, rot_limit=rl))
        log['ents'] = out

    def randomise_pose(self, entity, **kwargs):
        from magical_b200 import placement
        log = getattr(self, '_placement_log', None)
        if log is not None:
            log['unsupported'] = 'a separate randomise_pose call'
        placement.randomise_pose(self._builder, entity,
                                 self.ARENA_BOUNDS_LRBT, self.rng, **kwargs)

    def shift_entity(self, entity, position=None, angle=None):
        from magical_b200 import placement
        placement.shift_entity(self._builder, entity, position=position,
                               angle=angle)

    def entity_pos(self, entity):
        from magical_b200 import placement
        return placement.entity_pose(self._builder, entity)[0]
