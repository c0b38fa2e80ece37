"""MoveToRegion: drive the robot into the coloured goal region.
Restates reference `magical/benchmarks/move_to_region.py`."""
import numpy as np

from magical_b200 import entities as en
from magical_b200 import scene as sc
from magical_b200.base_env import BaseEnv

SMALL_POS_BOUND = 0.05
DEFAULT_ROBOT_POSE = ((0.058, 0.53), -2.13)
DEFAULT_GOAL_COLOUR = en.ShapeColour.BLUE
DEFAULT_GOAL_XYHW = (-0.62, -0.17, 0.76, 0.75)


class MoveToRegionEnv(BaseEnv):
    TASK_ID = sc.TASK_MOVE_TO_REGION

    def __init__(self, rand_poses_minor=False, rand_poses_full=False,
                 rand_goal_colour=False, **kwargs):
        super().__init__(**kwargs)
        assert not (rand_poses_minor and rand_poses_full), \
            "cannot specify both 'rand_poses_minor' and 'rand_poses_full'"
        self.rand_poses_minor = rand_poses_minor
        self.rand_poses_full = rand_poses_full
        self.rand_goal_colour = rand_goal_colour

    def on_reset(self):
        goal_xyhw = DEFAULT_GOAL_XYHW
        if self.rand_poses_minor or self.rand_poses_full:
            hw_bound = self.JITTER_TARGET_BOUND if self.rand_poses_minor \
                else None
            sampled_hw = self.randomise_hw(self.RAND_GOAL_MIN_SIZE,
                                           self.RAND_GOAL_MAX_SIZE,
                                           current_hw=goal_xyhw[2:],
                                           linf_bound=hw_bound)
            goal_xyhw = (*goal_xyhw[:2], *sampled_hw)
        if self.rand_goal_colour:
            goal_colour = self.rng.choice(
                np.asarray(en.SHAPE_COLOURS, dtype='object'))
        else:
            goal_colour = DEFAULT_GOAL_COLOUR
        goal = en.GoalRegion(*goal_xyhw, goal_colour)
        self.add_entities([goal])
        self._goal_ref = goal

        robot = self._make_robot(*DEFAULT_ROBOT_POSE)
        self.add_entities([robot])

        if self.rand_poses_minor or self.rand_poses_full:
            if self.rand_poses_minor:
                pos_limits = self.JITTER_POS_BOUND
                rot_limits = [None, self.JITTER_ROT_BOUND]
            else:
                pos_limits = rot_limits = None
            self.randomise_all_poses((self._goal_ref, self._robot),
                                     rand_pos=True, rand_rot=(False, True),
                                     rel_pos_linf_limits=pos_limits,
                                     rel_rot_limits=rot_limits)
