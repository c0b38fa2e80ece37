"""World entities (robot, pushable shapes, goal regions, arena walls) as
scene-table builders.

Mirrors the class names, constructor arguments and constants of the reference
`magical/entities.py`, but `setup()` targets a `scene.SceneBuilder` instead of
a live `pymunk.Space` + pyglet `Viewer`: every body, collision shape, joint
and draw geom the reference creates imperatively (entities.py:238-437,
502-537, 614-757, 790-819) is appended, in the same order, to flat tables
that the CUDA library steps and rasterises.
"""
import enum
import math

import numpy as np

from magical_b200 import geom as gtools
from magical_b200 import scene as sc
from magical_b200.style import (
    COLOURS_RGB, GOAL_LINE_THICKNESS, ROBOT_LINE_THICKNESS,
    SHAPE_LINE_THICKNESS, darken_rgb, lighten_rgb, to_u8)

INF = float('inf')


# ---------------------------------------------------------------------------
# Actions (reference entities.py:148-190)
# ---------------------------------------------------------------------------

class RobotAction(enum.IntFlag):
    NONE = 0
    UP = 1
    DOWN = 2
    LEFT = 4
    RIGHT = 8
    OPEN = 16
    CLOSE = 32


def _action_table():
    # id = 9*grip + 3*lr + ud, grip in (open, close), lr in (none, left,
    # right), ud in (none, up, down)
    table = []
    for grip, grip_name in ((RobotAction.OPEN, 'Open'),
                            (RobotAction.CLOSE, 'Close')):
        for lr, lr_name in ((RobotAction.NONE, ''), (RobotAction.LEFT, 'Left'),
                            (RobotAction.RIGHT, 'Right')):
            for ud, ud_name in ((RobotAction.NONE, ''), (RobotAction.UP, 'Up'),
                                (RobotAction.DOWN, 'Down')):
                table.append((len(table), (ud, lr, grip),
                              ud_name + lr_name + grip_name))
    return tuple(table)


ACTION_NUMS_FLAGS_NAMES = _action_table()
ACTION_ID_TO_FLAGS = {a: flags for a, flags, _ in ACTION_NUMS_FLAGS_NAMES}
FLAGS_TO_ACTION_ID = {flags: a for a, flags, _ in ACTION_NUMS_FLAGS_NAMES}


# ---------------------------------------------------------------------------
# Shape vocabulary (reference entities.py:545-581)
# ---------------------------------------------------------------------------

class ShapeType(str, enum.Enum):
    TRIANGLE = 'triangle'
    SQUARE = 'square'
    PENTAGON = 'pentagon'
    HEXAGON = 'hexagon'
    OCTAGON = 'octagon'
    CIRCLE = 'circle'
    STAR = 'star'


class ShapeColour(str, enum.Enum):
    RED = 'red'
    GREEN = 'green'
    BLUE = 'blue'
    YELLOW = 'yellow'


SHAPE_TYPES = np.asarray([ShapeType.SQUARE, ShapeType.PENTAGON,
                          ShapeType.STAR, ShapeType.CIRCLE], dtype='object')
SHAPE_COLOURS = np.asarray([ShapeColour.RED, ShapeColour.GREEN,
                            ShapeColour.BLUE, ShapeColour.YELLOW],
                           dtype='object')

# integer codes stored in the compiled scene (mg_block_t / mg_goal_t)
SHAPE_TYPE_CODE = {ShapeType.SQUARE: 0, ShapeType.PENTAGON: 1,
                   ShapeType.STAR: 2, ShapeType.CIRCLE: 3,
                   ShapeType.TRIANGLE: 4, ShapeType.HEXAGON: 5,
                   ShapeType.OCTAGON: 6}
COLOUR_CODE = {ShapeColour.RED: 0, ShapeColour.GREEN: 1, ShapeColour.BLUE: 2,
               ShapeColour.YELLOW: 3}

ROBOT_GROUP = 1
GREY = COLOURS_RGB['grey']


class Entity:
    """Something that contributes bodies/shapes/joints/geoms to a scene."""

    def setup(self, builder):
        raise NotImplementedError


# ---------------------------------------------------------------------------
# Robot (reference entities.py:217-490)
# ---------------------------------------------------------------------------

class Robot(Entity):
    finger_rot_limit_outer = math.pi / 8
    finger_rot_limit_inner = 0.0

    def __init__(self, radius, init_pos, init_angle, mass=1.0):
        self.radius = radius
        self.init_pos = (float(init_pos[0]), float(init_pos[1]))
        self.init_angle = float(init_angle)
        self.mass = mass

    def setup(self, b):
        pv = b.phys_vars
        R = self.radius
        # main body + kinematic control body (entities.py:243-254)
        body = b.add_body(self.mass,
                          gtools.moment_for_circle(self.mass, 0, R),
                          self.init_pos, self.init_angle)
        control = b.add_body(0, 0, self.init_pos, self.init_angle,
                             kind=sc.BODY_KINEMATIC)
        # position servo: pivot with no positional correction, capped force;
        # heading servo: gear with capped correction rate (entities.py:255-263)
        b.add_joint(sc.JOINT_PIVOT, control, body, max_bias=0.0,
                    max_force=pv.robot_pos_joint_max_force)
        b.add_joint(sc.JOINT_GEAR, control, body, p0=0.0, p1=1.0,
                    error_bias=0.0, max_bias=2.5,
                    max_force=pv.robot_rot_joint_max_force)
        # googly eyes on damped rotary springs (entities.py:265-277); the
        # spring ignores max_bias/max_force
        eyes = []
        for _ in (-1, 1):
            eye_mass = self.mass / 10
            eye = b.add_body(eye_mass,
                             gtools.moment_for_circle(eye_mass, 0, R),
                             (0.0, 0.0), self.init_angle)
            b.add_joint(sc.JOINT_ROTARY_SPRING, body, eye, p0=0.0, p1=0.1,
                        p2=3e-3, max_bias=3.0, max_force=0.001)
            eyes.append(eye)
        # fingers (entities.py:279-354)
        thickness = 0.25 * R
        upper_len = 1.1 * R
        lower_len = 0.7 * R
        fingers, motors, finger_verts, finger_inner_verts = [], [], [], []
        for side in (-1, 1):
            verts = gtools.make_finger_vertices(upper_len, lower_len,
                                                thickness, side)
            finger_verts.append(verts)
            inner = gtools.make_finger_vertices(
                upper_len - ROBOT_LINE_THICKNESS * 2,
                lower_len - ROBOT_LINE_THICKNESS * 2,
                thickness - ROBOT_LINE_THICKNESS * 2, side)
            finger_inner_verts.append(
                [[(x, y + ROBOT_LINE_THICKNESS) for x, y in box]
                 for box in inner])
            if side < 0:
                lower_lim = -self.finger_rot_limit_inner
                upper_lim = self.finger_rot_limit_outer
                delta_angle = upper_lim
            else:
                lower_lim = -self.finger_rot_limit_outer
                upper_lim = self.finger_rot_limit_inner
                delta_angle = lower_lim
            f_mass = self.mass / 8
            f_moment = gtools.moment_for_poly(f_mass, verts[0] + verts[1])
            rel_pos = (side * R * 0.45, R * 0.1)
            rel_rot = gtools.vrotated(rel_pos, self.init_angle)
            finger = b.add_body(f_mass, f_moment,
                                gtools.vadd(self.init_pos, rel_rot),
                                self.init_angle + delta_angle)
            fingers.append(finger)
            # zero-length pin (rest distance is measured at construction and
            # both anchors coincide), hard limit, weak motor
            b.add_joint(sc.JOINT_PIN, body, finger, anchor_a=rel_pos,
                        p0=0.0, error_bias=0.0)
            b.add_joint(sc.JOINT_ROTARY_LIMIT, body, finger, p0=lower_lim,
                        p1=upper_lim, error_bias=0.0)
            motors.append(b.add_joint(
                sc.JOINT_MOTOR, body, finger, max_bias=0.0,
                max_force=pv.robot_finger_max_force))
        # collision shapes (entities.py:356-375): all in the robot's filter
        # group so robot parts never collide with each other
        # reference `Entity.bodies` / `Entity.shapes` order (entities.py:247-374)
        self.body_ids = [body, control, *eyes, *fingers]
        self.cgroup_ids = [b.add_shapes(
            body, [(sc.SHAPE_CIRCLE, [(0.0, 0.0)], R, 0.5, ROBOT_GROUP)],
            robot_group=True)]
        finger_hulls = []
        for finger, verts in zip(fingers, finger_verts):
            hulls = [gtools.convex_hull(sub) for sub in verts]
            finger_hulls.append(hulls)
            self.cgroup_ids.append(b.add_shapes(
                finger, [(sc.SHAPE_POLY, h, 0.0, 5.0, ROBOT_GROUP)
                         for h in hulls], robot_group=True))

        # graphics (entities.py:377-437), painter's order preserved
        dark = to_u8(darken_rgb(GREY))
        light = to_u8(lighten_rgb(GREY, 4))
        grey = to_u8(GREY)
        for finger, hulls in zip(fingers, finger_hulls):
            for h in hulls:
                b.add_poly_prim(h, grey, body=finger)
        for finger, inner in zip(fingers, finger_inner_verts):
            for box in inner:
                b.add_poly_prim(box, light, body=finger)
        b.add_ngon_prim(100, R, dark, body)
        b.add_ngon_prim(100, R - ROBOT_LINE_THICKNESS, grey, body)
        for x_sign, eye in zip((-1, 1), eyes):
            centre = (x_sign * 0.4 * R, 0.3 * R)
            b.add_ngon_prim(20, 0.2 * R, (255, 255, 255), body, centre=centre)
            b.add_ngon_prim(10, 0.12 * R, to_u8((0.1, 0.1, 0.1)), body,
                            centre=centre, pupil_eye=eye,
                            pre_offset=(0.0, R * 0.07))
        b.robot = dict(robot_body=body, control_body=control,
                       finger_body=fingers, motor_joint=motors,
                       eye_body=eyes, robot_radius=R)


# ---------------------------------------------------------------------------
# Arena walls (reference entities.py:493-537)
# ---------------------------------------------------------------------------

class ArenaBoundaries(Entity):
    def __init__(self, left, right, top, bottom, seg_rad=1):
        self.left, self.right, self.top, self.bottom = left, right, top, bottom
        self.seg_rad = seg_rad

    def setup(self, b):
        rad = self.seg_rad
        pts = [(self.left - rad, self.top + rad),
               (self.right + rad, self.top + rad),
               (self.right + rad, self.bottom - rad),
               (self.left - rad, self.bottom - rad)]
        self.body_ids = []
        self.cgroup_ids = []
        for start, end in zip(pts, pts[1:] + pts[:1]):
            # each wall is its own collision group: broadphase tests them
            # individually
            self.cgroup_ids.append(b.add_shapes(
                -1, [(sc.SHAPE_SEGMENT, [start, end], rad, 0.8, 0)]))
        w = self.right - self.left
        h = self.top - self.bottom
        rect = [(-w / 2, h / 2), (w / 2, h / 2), (w / 2, -h / 2),
                (-w / 2, -h / 2)]
        b.add_poly_prim(rect, (255, 255, 255))
        # LineWidth(GOAL_LINE_THICKNESS) is enabled before the PolyLine's own
        # LineWidth(1) (attrs run in reverse), so the border is 1 px wide
        b.add_lineloop_prim(rect, to_u8(GREY), 1.0)


# ---------------------------------------------------------------------------
# Pushable shapes (reference entities.py:584-761)
# ---------------------------------------------------------------------------

_REGULAR = {ShapeType.TRIANGLE: (3, 0.8), ShapeType.PENTAGON: (5, 1.0),
            ShapeType.HEXAGON: (6, 1.0), ShapeType.OCTAGON: (8, 1.0)}


def make_rect_points(width, height):
    """gym_render.make_rect vertex order (gym_render.py:449-457)."""
    rw, rh = width / 2, height / 2
    return [(-rw, rh), (rw, rh), (rw, -rh), (-rw, -rh)]


class Shape(Entity):
    def __init__(self, shape_type, colour_name, shape_size, init_pos,
                 init_angle, mass=0.5):
        self.shape_type = ShapeType(shape_type)
        self.colour_name = ShapeColour(colour_name)
        self.colour = COLOURS_RGB[self.colour_name.value]
        self.shape_size = shape_size
        self.init_pos = (float(init_pos[0]), float(init_pos[1]))
        self.init_angle = float(init_angle)
        self.mass = mass
        self.block_index = None

    def setup(self, b):
        pv = b.phys_vars
        size = self.shape_size
        st = self.shape_type
        fric = 0.5
        inner_rgb = to_u8(self.colour)
        outer_rgb = to_u8(darken_rgb(self.colour))
        if st == ShapeType.SQUARE:
            side = math.sqrt(math.pi) * size
            hw = side / 2
            # Poly.create_box vertex order, bevel radius 1% of the side; body
            # mass/moment are derived from the shape (shape.mass = m)
            verts = [(hw, -hw), (hw, hw), (-hw, hw), (-hw, -hw)]
            moment = self.mass * gtools.moment_for_poly(1.0, verts)
            shapes = [(sc.SHAPE_POLY, verts, 0.01 * side, fric, 0)]
            outer = [('poly', make_rect_points(side, side))]
            inner = [('poly', make_rect_points(
                side - 2 * SHAPE_LINE_THICKNESS,
                side - 2 * SHAPE_LINE_THICKNESS))]
        elif st == ShapeType.CIRCLE:
            moment = gtools.moment_for_circle(self.mass, 0, size)
            shapes = [(sc.SHAPE_CIRCLE, [(0.0, 0.0)], size, fric, 0)]
            outer = [('ngon', size)]
            inner = [('ngon', size - SHAPE_LINE_THICKNESS)]
        elif st == ShapeType.STAR:
            out_rad = 1.3 * size
            in_rad = 0.5 * out_rad
            star = gtools.compute_star_verts(5, out_rad, in_rad)
            parts = gtools.convex_decomposition(star + star[:1], 0)
            hull = gtools.to_convex_hull(star, 1e-5)
            moment = gtools.moment_for_poly(self.mass, hull)
            group = b.new_group_id()
            shapes = [(sc.SHAPE_POLY, p, 0.0, fric, group) for p in parts]
            short = gtools.compute_star_verts(
                5, out_rad - SHAPE_LINE_THICKNESS,
                in_rad - SHAPE_LINE_THICKNESS)
            short_parts = gtools.convex_decomposition(short + short[:1], 0)
            outer = [('poly', p) for p in parts]
            inner = [('poly', p) for p in short_parts]
        else:
            n_sides, factor = _REGULAR[st]
            side = factor * gtools.regular_poly_circ_rad_to_side_length(
                n_sides, size)
            verts = gtools.compute_regular_poly_verts(n_sides, side)
            moment = gtools.moment_for_poly(self.mass, verts)
            shapes = [(sc.SHAPE_POLY, gtools.convex_hull(verts), 0.0, fric,
                       0)]
            apothem = gtools.regular_poly_side_length_to_apothem(n_sides,
                                                                 side)
            short_side = gtools.regular_poly_apothem_to_side_length(
                n_sides, apothem - SHAPE_LINE_THICKNESS)
            outer = [('poly', verts)]
            inner = [('poly', gtools.compute_regular_poly_verts(n_sides,
                                                                short_side))]
        body = b.add_body(self.mass, moment, self.init_pos, self.init_angle)
        cgroup = b.add_shapes(body, shapes)
        self.body_ids = [body]
        self.cgroup_ids = [cgroup]
        # table friction: force-capped pivot + gear against the static body
        # (entities.py:703-711)
        b.add_joint(sc.JOINT_PIVOT, -1, body, max_bias=0.0,
                    max_force=pv.shape_trans_joint_max_force)
        b.add_joint(sc.JOINT_GEAR, -1, body, p0=0.0, p1=1.0, max_bias=0.0,
                    max_force=pv.shape_rot_joint_max_force)
        for geoms, colour in ((outer, outer_rgb), (inner, inner_rgb)):
            for kind, data in geoms:
                if kind == 'poly':
                    b.add_poly_prim(data, colour, body=body)
                else:
                    b.add_ngon_prim(100, data, colour, body)
        self.block_index = len(b.blocks)
        b.blocks.append(dict(body=body, cgroup=cgroup,
                             shape_type=SHAPE_TYPE_CODE[st],
                             colour=COLOUR_CODE[self.colour_name], role=0,
                             label=0))


# ---------------------------------------------------------------------------
# Goal regions (reference entities.py:769-886)
# ---------------------------------------------------------------------------

class GoalRegion(Entity):
    """Axis-aligned sensor box; (x, y) is its TOP-LEFT corner."""

    def __init__(self, x, y, h, w, colour_name):
        assert h > 0 and w > 0
        self.x, self.y, self.h, self.w = float(x), float(y), float(h), float(w)
        self.colour_name = ShapeColour(colour_name)
        self.base_colour = COLOURS_RGB[self.colour_name.value]
        self.goal_index = None

    @property
    def centre(self):
        return (self.x + self.w / 2, self.y - self.h / 2)

    def _corners(self, cx, cy):
        return [(px + cx, py + cy)
                for px, py in make_rect_points(self.w, self.h)]

    def move_to(self, b, cx, cy):
        """Re-centre the region (reset-time randomisation moves the goal's
        static body, geom.py:362-384; its geoms follow in `pre_draw`,
        entities.py:883-886)."""
        cx, cy = float(cx), float(cy)
        self.x, self.y = cx - self.w / 2, cy + self.h / 2
        goal = b.goals[self.goal_index]
        goal['cx'], goal['cy'] = cx, cy
        pts = self._corners(cx, cy)
        for pi in self.prim_ids:
            v0 = b.prims[pi]['vert0']
            b.dverts[v0:v0 + 4] = pts

    def setup(self, b):
        cx, cy = self.centre
        pts = self._corners(cx, cy)
        self.body_ids = []
        self.cgroup_ids = []
        self.prim_ids = (len(b.prims), len(b.prims) + 1)
        b.add_poly_prim(pts, to_u8(lighten_rgb(self.base_colour, times=2)))
        b.add_lineloop_prim(pts, to_u8(self.base_colour),
                            250 * GOAL_LINE_THICKNESS, stipple=0x00FF)
        self.goal_index = len(b.goals)
        b.goals.append(dict(cx=cx, cy=cy, w=self.w, h=self.h,
                            colour=COLOUR_CODE[self.colour_name],
                            expect_block=-1))
