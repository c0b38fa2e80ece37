"""ClusterColour / ClusterShape: sort blocks into one cluster per colour or
per shape type.  Restates reference `magical/benchmarks/cluster.py`."""
import enum

import numpy as np

from magical_b200 import entities as en
from magical_b200 import scene as sc
from magical_b200.base_env import BaseEnv

C, T = en.ShapeColour, en.ShapeType


class BaseClusterEnv(BaseEnv):
    class ClusterBy(str, enum.Enum):
        COLOUR = 'colour'
        TYPE = 'type'

    def __init__(self, rand_shape_colour=False, rand_shape_type=False,
                 rand_layout_minor=False, rand_layout_full=False,
                 rand_shape_count=False, cluster_by=ClusterBy.COLOUR,
                 **kwargs):
        super().__init__(**kwargs)
        self.rand_shape_colour = rand_shape_colour
        self.rand_shape_type = rand_shape_type
        self.rand_shape_count = rand_shape_count
        assert not (rand_layout_minor and rand_layout_full)
        self.rand_layout_minor = rand_layout_minor
        self.rand_layout_full = rand_layout_full
        self.cluster_by = cluster_by
        if self.rand_shape_count:
            assert self.rand_layout_full and self.rand_shape_type \
                and self.rand_shape_colour, \
                "if shape count is randomised then layout, type and colour " \
                "must be randomised too"

    def on_reset(self):
        robot = self._make_robot(*self.DEFAULT_ROBOT_POSE)
        colours = self.DEFAULT_BLOCK_COLOURS
        shape_types = self.DEFAULT_BLOCK_SHAPES
        poses = self.DEFAULT_BLOCK_POSES
        n_shapes = len(colours)
        if self.rand_shape_count:
            n_shapes = self.rng.randint(7, 10 + 1)
            poses = [((0, 0), 0)] * n_shapes
        if self.rand_shape_colour:
            # at least one of each colour (cluster.py:90-98)
            colours = list(en.SHAPE_COLOURS)
            colours.extend([self.rng.choice(en.SHAPE_COLOURS)
                            for _ in range(n_shapes - len(colours))])
            self.rng.shuffle(colours)
        if self.rand_shape_type:
            shape_types = list(en.SHAPE_TYPES)
            shape_types.extend([self.rng.choice(en.SHAPE_TYPES)
                                for _ in range(n_shapes - len(shape_types))])
            self.rng.shuffle(shape_types)
        assert len(poses) == len(colours) == len(shape_types) == n_shapes

        shape_ents = [
            self._make_shape(shape_type=st, colour_name=col, init_pos=(x, y),
                             init_angle=angle)
            for ((x, y), angle), col, st in zip(poses, colours, shape_types)]
        self.add_entities(shape_ents)
        self._shape_ents = shape_ents
        if self.cluster_by == self.ClusterBy.COLOUR:
            self._c_values = [en.ShapeColour(c).value for c in colours]
        else:
            self._c_values = [en.ShapeType(t).value for t in shape_types]
        # robot last so it is drawn on top (cluster.py:141-144)
        self.add_entities([robot])

        if self.rand_layout_full or self.rand_layout_minor:
            if self.rand_layout_full:
                pos_limit = rot_limit = None
            else:
                pos_limit = self.JITTER_POS_BOUND
                rot_limit = self.JITTER_ROT_BOUND
            self.randomise_all_poses([robot, *shape_ents], rand_pos=True,
                                     rand_rot=True,
                                     rel_pos_linf_limits=pos_limit,
                                     rel_rot_limits=rot_limit)

    def finalise_scene(self, builder):
        # np.unique order of the characteristic values = centroid index
        # (cluster.py:127-139); the score itself runs on the device
        uniq = sorted(set(self._c_values))
        builder.n_labels = len(uniq)
        for ent, value in zip(self._shape_ents, self._c_values):
            builder.blocks[ent.block_index]['label'] = uniq.index(value)


class ClusterColourEnv(BaseClusterEnv):
    TASK_ID = sc.TASK_CLUSTER_COLOUR
    DEFAULT_ROBOT_POSE = ((0.71692, -0.34374), 0.83693)
    DEFAULT_BLOCK_COLOURS = [C.BLUE, C.BLUE, C.BLUE, C.GREEN, C.GREEN, C.RED,
                             C.YELLOW, C.YELLOW]
    DEFAULT_BLOCK_SHAPES = [T.CIRCLE, T.STAR, T.SQUARE, T.PENTAGON, T.PENTAGON,
                            T.SQUARE, T.STAR, T.PENTAGON]
    DEFAULT_BLOCK_POSES = [
        ((-0.5147, 0.14149), -0.38871),
        ((-0.1347, -0.71414), 1.0533),
        ((-0.74247, -0.097592), 1.1571),
        ((-0.077363, -0.42964), -0.64379),
        ((0.51978, 0.1853), -1.1762),
        ((-0.5278, -0.21642), 2.9356),
        ((-0.54039, 0.48292), 0.072818),
        ((-0.16761, 0.64303), -2.3255),
    ]

    def __init__(self, *args, **kwargs):
        super().__init__(*args, cluster_by=BaseClusterEnv.ClusterBy.COLOUR,
                         **kwargs)


class ClusterShapeEnv(BaseClusterEnv):
    TASK_ID = sc.TASK_CLUSTER_SHAPE
    DEFAULT_ROBOT_POSE = ((0.286, -0.202), -1.878)
    DEFAULT_BLOCK_COLOURS = [C.YELLOW, C.BLUE, C.RED, C.RED, C.GREEN,
                             C.YELLOW, C.BLUE, C.GREEN]
    DEFAULT_BLOCK_SHAPES = [T.SQUARE, T.PENTAGON, T.PENTAGON, T.PENTAGON,
                            T.CIRCLE, T.STAR, T.STAR, T.CIRCLE]
    DEFAULT_BLOCK_POSES = [
        ((-0.414, 0.297), -1.731),
        ((0.068, 0.705), 2.184),
        ((0.821, 0.220), 0.650),
        ((-0.461, -0.749), -2.673),
        ((0.867, -0.149), -2.215),
        ((-0.785, -0.140), -0.405),
        ((-0.305, -0.226), 1.341),
        ((0.758, -0.708), -2.140),
    ]

    def __init__(self, *args, **kwargs):
        super().__init__(*args, cluster_by=BaseClusterEnv.ClusterBy.TYPE,
                         **kwargs)
