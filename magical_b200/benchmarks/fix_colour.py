"""FixColour: every goal region holds one block; remove the one block whose
colour does not match its region.  Restates reference
`magical/benchmarks/fix_colour.py`."""
from magical_b200 import entities as en
from magical_b200 import scene as sc
from magical_b200.base_env import BaseEnv

C, T = en.ShapeColour, en.ShapeType
MIN_REGIONS = 2
MAX_REGIONS = 3
MIN_GOAL_SIZE = 0.4
MAX_GOAL_SIZE = 0.5
DEFAULT_ROBOT_POSE = ((0.368, 0.586), 0.718)
DEFAULT_BLOCK_COLOURS = [C.GREEN, C.GREEN, C.BLUE]
DEFAULT_BLOCK_SHAPES = [T.PENTAGON, T.SQUARE, T.PENTAGON]
DEFAULT_BLOCK_POSES = [((0.289, 0.030), 0.307), ((0.133, -0.561), 1.699),
                       ((-0.336, 0.000), -1.529)]
DEFAULT_REGION_XYHWS = [(-0.032, 0.348, 0.427, 0.468),
                        (0.019, -0.391, 0.460, 0.458),
                        (-0.681, 0.196, 0.498, 0.418)]
DEFAULT_REGION_COLOURS = [C.GREEN, C.GREEN, C.RED]


class FixColourEnv(BaseEnv):
    TASK_ID = sc.TASK_FIX_COLOUR

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

    def on_reset(self):
        robot = self._make_robot(*DEFAULT_ROBOT_POSE)
        block_colours = DEFAULT_BLOCK_COLOURS
        region_colours = DEFAULT_REGION_COLOURS
        block_shapes = DEFAULT_BLOCK_SHAPES
        block_poses = DEFAULT_BLOCK_POSES
        region_xyhws = DEFAULT_REGION_XYHWS
        n_regions = len(block_colours)
        if self.rand_count:
            n_regions = self.rng.randint(MIN_REGIONS, MAX_REGIONS + 1)
            block_poses = block_poses[:1] * n_regions
            region_xyhws = region_xyhws[:1] * n_regions
        if self.rand_colours:
            region_colours = self.rng.choice(en.SHAPE_COLOURS,
                                             size=n_regions).tolist()
            block_colours = list(region_colours)
            # one block gets a colour different from its region's
            odd_idx = self.rng.randint(len(block_colours))
            new_col_idx = self.rng.randint(len(en.SHAPE_COLOURS) - 1)
            if en.SHAPE_COLOURS[new_col_idx] == block_colours[odd_idx]:
                new_col_idx += 1
            block_colours[odd_idx] = en.SHAPE_COLOURS[new_col_idx]
        if self.rand_shapes:
            block_shapes = self.rng.choice(en.SHAPE_TYPES,
                                           size=n_regions).tolist()
        if self.rand_layout_minor or self.rand_layout_full:
            hw_bound = self.JITTER_TARGET_BOUND if self.rand_layout_minor \
                else None
            region_xyhws = [
                (x, y, *self.randomise_hw(MIN_GOAL_SIZE, MAX_GOAL_SIZE,
                                          current_hw=hw, linf_bound=hw_bound))
                for x, y, *hw in region_xyhws]
        sensors = [en.GoalRegion(*xyhw, colour)
                   for colour, xyhw in zip(region_colours, region_xyhws)]
        self.add_entities(sensors)
        self._sensors = sensors

        blocks = []
        self._target_blocks = []
        for bshape, bcol, tcol, (bpos, bangle) in zip(
                block_shapes, block_colours, region_colours, block_poses):
            block = self._make_shape(shape_type=bshape, colour_name=bcol,
                                     init_pos=bpos, init_angle=bangle)
            blocks.append(block)
            # mismatching region must end up empty, the others keep theirs
            self._target_blocks.append([] if bcol != tcol else [block])
        self.add_entities(blocks)
        self.add_entities([robot])

        if self.rand_layout_minor or self.rand_layout_full:
            if self.rand_layout_minor:
                pos_limits = self.JITTER_POS_BOUND
                rot_limit = self.JITTER_ROT_BOUND
            else:
                pos_limits = rot_limit = None
            rand_rot = [False] * n_regions + [True]
            self.randomise_all_poses((*sensors, robot), rand_pos=True,
                                     rand_rot=rand_rot,
                                     rel_pos_linf_limits=pos_limits,
                                     rel_rot_limits=rot_limit,
                                     ignore_ents=blocks)
            for block, sensor in zip(blocks, sensors):
                self.shift_entity(block, position=self.entity_pos(sensor))
            for block, sensor, (_, _, *sensor_hw) in zip(blocks, sensors,
                                                         region_xyhws):
                block_pos_limit = max(0, min(sensor_hw) / 2 - self.SHAPE_RAD)
                if self.rand_layout_minor:
                    block_pos_limit = min(self.JITTER_POS_BOUND,
                                          block_pos_limit)
                self.randomise_pose(block, ignore_ents=[sensor],
                                    rand_pos=True, rand_rot=True,
                                    rel_pos_linf_limit=block_pos_limit,
                                    rel_rot_limit=rot_limit)

    def finalise_scene(self, builder):
        # device score = 1 iff each goal holds exactly its expected block
        # list (fix_colour.py:193-202)
        for sensor, expected in zip(self._sensors, self._target_blocks):
            builder.goals[sensor.goal_index]['expect_block'] = \
                expected[0].block_index if expected else -1
