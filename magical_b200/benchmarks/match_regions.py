"""MatchRegions: push all blocks of the goal's colour into the goal region and
keep the others out.  Restates reference
`magical/benchmarks/match_regions.py`."""
import math

import numpy as np

from magical_b200 import entities as en
from magical_b200 import scene as sc
from magical_b200.base_env import BaseEnv

T = en.ShapeType
DEFAULT_TARGET_TYPES = [T.STAR, T.SQUARE]
# one list per distractor colour, in SHAPE_COLOURS order minus the target
DEFAULT_DISTRACTOR_TYPES = [[], [T.PENTAGON], [T.CIRCLE, T.PENTAGON]]
DEFAULT_TARGET_POSES = [(0.8, -0.7, 2.37), (-0.68, 0.72, 1.28)]
DEFAULT_DISTRACTOR_POSES = [[], [(-0.05, -0.2, -1.09)],
                            [(-0.75, -0.55, 2.78), (0.3, -0.82, -1.15)]]


class MatchRegionsEnv(BaseEnv):
    TASK_ID = sc.TASK_MATCH_REGIONS

    def __init__(self, rand_target_colour=False, rand_shape_type=False,
                 rand_shape_count=False, rand_layout_minor=False,
                 rand_layout_full=False, **kwargs):
        super().__init__(**kwargs)
        self.rand_target_colour = rand_target_colour
        self.rand_shape_type = rand_shape_type
        self.rand_shape_count = rand_shape_count
        self.rand_layout_minor = rand_layout_minor
        self.rand_layout_full = rand_layout_full
        if self.rand_shape_count:
            assert self.rand_layout_full and self.rand_shape_type \
                and self.rand_target_colour, \
                "randomised shape count needs full layout, type and colour " \
                "randomisation"

    def on_reset(self):
        robot = self._make_robot(np.asarray((-0.5, 0.1)), -math.pi * 1.2)
        if self.rand_target_colour:
            target_colour = self.rng.choice(en.SHAPE_COLOURS)
        else:
            target_colour = en.ShapeColour.GREEN
        distractor_colours = [c for c in en.SHAPE_COLOURS
                              if c != target_colour]
        target_h, target_w, target_x, target_y = 0.7, 0.6, 0.1, 0.7
        if self.rand_layout_minor or self.rand_layout_full:
            hw_bound = self.JITTER_TARGET_BOUND if self.rand_layout_minor \
                else None
            target_h, target_w = self.randomise_hw(
                self.RAND_GOAL_MIN_SIZE, self.RAND_GOAL_MAX_SIZE,
                current_hw=(target_h, target_w), linf_bound=hw_bound)
        sensor = en.GoalRegion(target_x, target_y, target_h, target_w,
                               target_colour)
        self.add_entities([sensor])
        self._sensor_ref = sensor

        if self.rand_shape_count:
            target_count = self.rng.randint(1, 2 + 1)
            distractor_counts = [self.rng.randint(0, 2 + 1)
                                 for _ in distractor_colours]
        else:
            target_count = len(DEFAULT_TARGET_TYPES)
            distractor_counts = [len(l) for l in DEFAULT_DISTRACTOR_TYPES]
        if self.rand_shape_type:
            types_np = np.asarray(en.SHAPE_TYPES, dtype='object')
            target_types = [self.rng.choice(types_np)
                            for _ in range(target_count)]
            distractor_types = [[self.rng.choice(types_np) for _ in range(n)]
                                for n in distractor_counts]
        else:
            target_types = DEFAULT_TARGET_TYPES
            distractor_types = DEFAULT_DISTRACTOR_TYPES
        if self.rand_layout_full:
            target_poses = [(0, 0, 0)] * target_count
            distractor_poses = [[(0, 0, 0)] * n for n in distractor_counts]
        else:
            target_poses = DEFAULT_TARGET_POSES
            distractor_poses = DEFAULT_DISTRACTOR_POSES

        self._target_shapes = [
            self._make_shape(shape_type=st, colour_name=target_colour,
                             init_pos=(x, y), init_angle=a)
            for st, (x, y, a) in zip(target_types, target_poses)]
        self._distractor_shapes = []
        for colour, types, poses in zip(distractor_colours, distractor_types,
                                        distractor_poses):
            for st, (x, y, a) in zip(types, poses):
                self._distractor_shapes.append(
                    self._make_shape(shape_type=st, colour_name=colour,
                                     init_pos=(x, y), init_angle=a))
        shape_ents = self._target_shapes + self._distractor_shapes
        self.add_entities(shape_ents)
        self.add_entities([robot])

        if self.rand_layout_minor or self.rand_layout_full:
            all_ents = (sensor, robot, *shape_ents)
            if self.rand_layout_minor:
                pos_limits = self.JITTER_POS_BOUND
                rot_limits = self.JITTER_ROT_BOUND
            else:
                pos_limits = rot_limits = None
            rand_rot = [False] + [True] * (len(all_ents) - 1)
            self.randomise_all_poses(all_ents, rand_pos=True,
                                     rand_rot=rand_rot,
                                     rel_pos_linf_limits=pos_limits,
                                     rel_rot_limits=rot_limits)

    def finalise_scene(self, builder):
        # device score = frac(targets in goal) * (1 - contamination)
        # (match_regions.py:193-213)
        for ent in self._target_shapes:
            builder.blocks[ent.block_index]['role'] = 1
        for ent in self._distractor_shapes:
            builder.blocks[ent.block_index]['role'] = 2
