"""MoveToCorner: push the single block into the top-left corner.
Scene + score restate reference `magical/benchmarks/move_to_corner.py`."""
import math
import warnings

import numpy as np

from magical_b200 import entities as en
from magical_b200 import scene as sc
from magical_b200.base_env import BaseEnv


class MoveToCornerEnv(BaseEnv):
    TASK_ID = sc.TASK_MOVE_TO_CORNER

    def __init__(self, rand_shape_colour=False, rand_shape_type=False,
                 rand_poses=False, debug_reward=False, **kwargs):
        super().__init__(**kwargs)
        self.rand_shape_colour = rand_shape_colour
        self.rand_shape_type = rand_shape_type
        self.rand_poses = rand_poses
        self.debug_reward = debug_reward
        if self.debug_reward:
            warnings.warn(
                "DEBUG REWARD ENABLED IN MOVE-TO-CORNER ENV! This reward is "
                "ONLY intended for training RL algorithms during debugging")

    def on_reset(self):
        # robot first, so the block is drawn on top of it
        # (move_to_corner.py:33-37)
        robot = self._make_robot(np.asarray((0.4, -0.0)), 0.55 * math.pi)
        self.add_entities([robot])

        shape_colour = 'red'
        shape_type = en.ShapeType.SQUARE
        if self.rand_shape_colour:
            shape_colour = self.rng.choice(
                np.asarray(en.SHAPE_COLOURS, dtype='object'))
        if self.rand_shape_type:
            shape_type = self.rng.choice(
                np.asarray(en.SHAPE_TYPES, dtype='object'))
        shape = self._make_shape(shape_type=shape_type,
                                 colour_name=shape_colour,
                                 init_pos=np.asarray((0.1, -0.65)),
                                 init_angle=0.13 * math.pi)
        self.add_entities([shape])
        self._shape_ref = shape

        if self.rand_poses:
            self.randomise_all_poses(
                (self._robot, self._shape_ref), rand_pos=True, rand_rot=True,
                rel_pos_linf_limits=self.JITTER_POS_BOUND,
                rel_rot_limits=self.JITTER_ROT_BOUND)

    def finalise_scene(self, builder):
        # score: distance of block 0 from the (-1, 1) corner, evaluated on
        # the device (move_to_corner.py:66-75)
        builder.blocks[self._shape_ref.block_index]['role'] = 1
