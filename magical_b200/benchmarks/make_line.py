"""MakeLine: arrange all blocks in a straight line.  Restates reference
`magical/benchmarks/make_line.py` (the RANSAC-style `longest_line` score runs
on the device; `longest_line` below is the host restatement used by tests)."""
import itertools as it

import numpy as np

from magical_b200 import entities as en
from magical_b200 import scene as sc
from magical_b200.base_env import BaseEnv

INLIER_RAD_MULT = 1.5
MAX_SEP_RADS = 3.5
MIN_BLOCKS = 3
MAX_BLOCKS = 4
DEFAULT_ROBOT_POSE = ((0.702, -0.255), 0.347)
DEFAULT_BLOCK_COLOURS = [en.ShapeColour.BLUE, en.ShapeColour.YELLOW,
                         en.ShapeColour.RED, en.ShapeColour.GREEN]
DEFAULT_BLOCK_SHAPES = [en.ShapeType.STAR, en.ShapeType.CIRCLE,
                        en.ShapeType.STAR, en.ShapeType.PENTAGON]
DEFAULT_BLOCK_POSES = [((0.790, -0.820), -0.721), ((-0.177, 0.383), -1.733),
                       ((-0.051, -0.128), 2.696), ((-0.292, -0.745), -0.159)]


def longest_line(points, inlier_dist, max_separation):
    """Size of the largest set of points lying within `inlier_dist` of some
    line through two of them, with consecutive projections at most
    `max_separation` apart (make_line.py:31-71)."""
    points = np.asarray(points, dtype='float64')
    npts = len(points)
    best = min(1, npts)
    for i, j in it.combinations(range(npts), 2):
        offs = points - points[i]
        unit = offs[j] / np.linalg.norm(offs[j])
        proj = offs @ unit
        dists = np.linalg.norm(offs - proj[:, None] * unit, axis=1)
        inliers = np.nonzero(dists <= inlier_dist)[0]
        if len(inliers) <= best:
            continue
        seps = np.abs(np.diff(np.sort(proj[inliers])))
        run = longest = 0
        for ok in seps <= max_separation:
            run = run + 1 if ok else 0
            longest = max(longest, run)
        best = max(best, longest + 1)
    return best


class MakeLineEnv(BaseEnv):
    TASK_ID = sc.TASK_MAKE_LINE

    def __init__(self, rand_colours=False, rand_shapes=False, rand_count=False,
                 rand_layout_minor=False, rand_layout_full=False, **kwargs):
        super().__init__(**kwargs)
        self.rand_colours = rand_colours
        self.rand_shapes = rand_shapes
        self.rand_count = rand_count
        self.rand_layout_minor = rand_layout_minor
        self.rand_layout_full = rand_layout_full
        if self.rand_count:
            assert self.rand_layout_full and self.rand_shapes \
                and self.rand_colours, "if shape count is randomised then " \
                "layout, shapes, and colours must be fully randomised too"
        self.inlier_dist = self.SHAPE_RAD * INLIER_RAD_MULT
        self.max_sep = self.SHAPE_RAD * MAX_SEP_RADS

    def on_reset(self):
        robot = self._make_robot(*DEFAULT_ROBOT_POSE)
        block_shapes = DEFAULT_BLOCK_SHAPES
        block_colours = DEFAULT_BLOCK_COLOURS
        block_poses = DEFAULT_BLOCK_POSES
        if self.rand_count:
            n_blocks = self.rng.randint(MIN_BLOCKS, MAX_BLOCKS + 1)
            block_poses = block_poses[:1] * n_blocks
        else:
            n_blocks = len(block_shapes)
        if self.rand_colours:
            block_colours = self.rng.choice(en.SHAPE_COLOURS,
                                            size=n_blocks).tolist()
        if self.rand_shapes:
            block_shapes = self.rng.choice(en.SHAPE_TYPES,
                                           size=n_blocks).tolist()
        self._blocks = [
            self._make_shape(shape_type=bshape, colour_name=bcol,
                             init_pos=bpos, init_angle=bangle)
            for bshape, bcol, (bpos, bangle) in zip(block_shapes,
                                                    block_colours,
                                                    block_poses)]
        self.add_entities(self._blocks)
        self.add_entities([robot])
        if self.rand_layout_minor or self.rand_layout_full:
            if self.rand_layout_minor:
                pos_limits = self.JITTER_POS_BOUND
                rot_limit = self.JITTER_ROT_BOUND
            else:
                pos_limits = rot_limit = None
            self.randomise_all_poses((robot, *self._blocks), rand_pos=True,
                                     rand_rot=True,
                                     rel_pos_linf_limits=pos_limits,
                                     rel_rot_limits=rot_limit)
