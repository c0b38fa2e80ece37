"""FindDupe: bring a duplicate of the query block into the goal region.
Restates reference `magical/benchmarks/find_dupe.py`."""
from magical_b200 import entities as en
from magical_b200 import scene as sc
from magical_b200.base_env import BaseEnv

C, T = en.ShapeColour, en.ShapeType
DEFAULT_QUERY_COLOUR = C.YELLOW
DEFAULT_QUERY_SHAPE = T.PENTAGON
# the last outside block always duplicates the query block
DEFAULT_OUT_BLOCK_SHAPES = [T.PENTAGON, T.CIRCLE, T.CIRCLE, T.SQUARE, T.STAR,
                            DEFAULT_QUERY_SHAPE]
DEFAULT_OUT_BLOCK_COLOURS = [C.GREEN, C.RED, C.RED, C.YELLOW, C.BLUE,
                             DEFAULT_QUERY_COLOUR]
DEFAULT_OUT_BLOCK_POSES = [
    ((-0.066751, 0.7552), -2.9266),
    ((-0.05195, 0.31468), 1.5418),
    ((0.57528, -0.46865), -2.2141),
    ((0.40594, -0.74977), 0.24582),
    ((0.45254, 0.3681), -1.0834),
    ((0.76849, -0.10652), 0.10028),
]
DEFAULT_ROBOT_POSE = ((-0.57, 0.25), 3.83)
DEFAULT_TARGET_REGION_XYHW = (-0.72, -0.22, 0.67, 0.72)
DEFAULT_QUERY_BLOCK_POSE = ((-0.33, -0.49), -0.51)


class FindDupeEnv(BaseEnv):
    TASK_ID = sc.TASK_FIND_DUPE

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
        query_colour = DEFAULT_QUERY_COLOUR
        query_shape = DEFAULT_QUERY_SHAPE
        out_colours = DEFAULT_OUT_BLOCK_COLOURS
        out_shapes = DEFAULT_OUT_BLOCK_SHAPES
        n_out = len(DEFAULT_OUT_BLOCK_COLOURS)
        if self.rand_count:
            n_out = self.rng.randint(1, 5 + 1) + 1
        n_distractors = n_out - 1
        if self.rand_colours:
            query_colour = self.rng.choice(en.SHAPE_COLOURS)
            out_colours = self.rng.choice(en.SHAPE_COLOURS,
                                          size=n_distractors).tolist()
            out_colours.append(query_colour)
        if self.rand_shapes:
            query_shape = self.rng.choice(en.SHAPE_TYPES)
            out_shapes = self.rng.choice(en.SHAPE_TYPES,
                                         size=n_distractors).tolist()
            out_shapes.append(query_shape)

        region_xyhw = DEFAULT_TARGET_REGION_XYHW
        if self.rand_layout_minor or self.rand_layout_full:
            hw_bound = self.JITTER_TARGET_BOUND if self.rand_layout_minor \
                else None
            target_hw = self.randomise_hw(self.RAND_GOAL_MIN_SIZE,
                                          self.RAND_GOAL_MAX_SIZE,
                                          current_hw=region_xyhw[2:],
                                          linf_bound=hw_bound)
            region_xyhw = (*region_xyhw[:2], *target_hw)
        sensor = en.GoalRegion(*region_xyhw, query_colour)
        self.add_entities([sensor])
        self._sensor_ref = sensor

        out_poses = DEFAULT_OUT_BLOCK_POSES
        if self.rand_count:
            out_poses = [((0, 0), 0)] * n_out
        outside_blocks = []
        self._target_set = []
        for bshape, bcol, (bpos, bangle) in zip(out_shapes, out_colours,
                                                out_poses):
            block = self._make_shape(shape_type=bshape, colour_name=bcol,
                                     init_pos=bpos, init_angle=bangle)
            outside_blocks.append(block)
            if bcol == query_colour and bshape == query_shape:
                self._target_set.append(block)
        self.add_entities(outside_blocks)
        query_block = self._make_shape(
            shape_type=query_shape, colour_name=query_colour,
            init_pos=DEFAULT_QUERY_BLOCK_POSE[0],
            init_angle=DEFAULT_QUERY_BLOCK_POSE[1])
        self._target_set.append(query_block)
        self.add_entities([query_block])
        self._distractor_set = [b for b in outside_blocks
                                if b not in self._target_set]
        self.add_entities([robot])

        if self.rand_layout_minor or self.rand_layout_full:
            all_ents = (sensor, robot, *outside_blocks)
            if self.rand_layout_minor:
                pos_limits = self.JITTER_POS_BOUND
                rot_limit = self.JITTER_ROT_BOUND
            else:
                pos_limits = rot_limit = None
            rand_rot = [False] + [True] * (len(all_ents) - 1)
            self.randomise_all_poses(all_ents, rand_pos=True,
                                     rand_rot=rand_rot,
                                     rel_pos_linf_limits=pos_limits,
                                     rel_rot_limits=rot_limit,
                                     ignore_ents=[query_block])
            # the query block goes (mostly) inside the freshly placed goal
            query_pos_limit = max(
                0, min(region_xyhw[2:]) / 2 - self.SHAPE_RAD / 2)
            if self.rand_layout_minor:
                query_pos_limit = min(self.JITTER_POS_BOUND, query_pos_limit)
            self.shift_entity(query_block, position=self.entity_pos(sensor))
            self.randomise_pose(query_block, ignore_ents=[sensor],
                                rand_pos=True, rand_rot=True,
                                rel_pos_linf_limit=query_pos_limit,
                                rel_rot_limit=rot_limit)

    def finalise_scene(self, builder):
        # device score = [>=2 target-set blocks in goal] * (1 - contamination)
        # (find_dupe.py:203-216)
        for ent in self._target_set:
            builder.blocks[ent.block_index]['role'] = 1
        for ent in self._distractor_set:
            builder.blocks[ent.block_index]['role'] = 2
