"""Compiled-scene tables: numpy mirror of `mg_scene_t` (include/magical_b200.h).

A `SceneBuilder` plays the role that `pm.Space` + `gym_render.Viewer` play for
`Entity.setup()` in the reference (magical/base_env.py:164-175): entities add
bodies, shapes, joints and draw geoms to it, in the reference's insertion
order, and `compile()` freezes the result into one fixed-size record that is
copied to the GPU (and handed to the CPU oracle in tests).
"""
import numpy as np

ABI_VERSION = 2
MAX_BODIES = 16
STATE_CACHE = 48
MAX_SHAPES = 72
MAX_CVERTS = 384
MAX_JOINTS = 32
MAX_CGROUPS = 20
MAX_BPAIRS = 160
MAX_GOALS = 3
MAX_BLOCKS = 10
MAX_PRIMS = 160
MAX_DVERTS = 704

SHAPE_CIRCLE, SHAPE_SEGMENT, SHAPE_POLY = 0, 1, 2
BODY_DYNAMIC, BODY_KINEMATIC = 0, 1
(JOINT_PIVOT, JOINT_GEAR, JOINT_ROTARY_SPRING, JOINT_PIN, JOINT_ROTARY_LIMIT,
 JOINT_MOTOR) = range(6)
PRIM_POLY, PRIM_NGON, PRIM_LINELOOP = 0, 1, 2
XFORM_WORLD, XFORM_BODY, XFORM_PUPIL = 0, 1, 2
(TASK_MOVE_TO_CORNER, TASK_MOVE_TO_REGION, TASK_MATCH_REGIONS, TASK_MAKE_LINE,
 TASK_FIND_DUPE, TASK_FIX_COLOUR, TASK_CLUSTER_COLOUR,
 TASK_CLUSTER_SHAPE) = range(8)
(OBS_LORES4E, OBS_LORES4A, OBS_LORES3EA, OBS_LORESSTACK, OBS_LORESCHW4E,
 OBS_RAW) = range(6)

INF = float('inf')

body_dt = np.dtype([('m_inv', 'f8'), ('i_inv', 'f8'), ('p0', 'f8', 2),
                    ('a0', 'f8'), ('kind', 'i4'), ('pad_', 'i4')], align=True)
shape_dt = np.dtype([('kind', 'i4'), ('body', 'i4'), ('vert0', 'i4'),
                     ('nvert', 'i4'), ('radius', 'f8'), ('friction', 'f8'),
                     ('group', 'i4'), ('pad_', 'i4')], align=True)
joint_dt = np.dtype([('kind', 'i4'), ('a', 'i4'), ('b', 'i4'), ('pad_', 'i4'),
                     ('anchor_a', 'f8', 2), ('anchor_b', 'f8', 2),
                     ('p0', 'f8'), ('p1', 'f8'), ('p2', 'f8'),
                     ('max_force', 'f8'), ('max_bias', 'f8'),
                     ('error_bias', 'f8')], align=True)
cgroup_dt = np.dtype([('shape0', 'u1'), ('nshape', 'u1'), ('body', 'i1'),
                      ('robot_group', 'u1')], align=True)
prim_dt = np.dtype([('kind', 'u1'), ('xform', 'u1'), ('body', 'u1'),
                    ('body2', 'u1'), ('rgb', 'u1', 3), ('nvert', 'u1'),
                    ('vert0', 'u2'), ('stipple', 'u2'), ('cx', 'f4'),
                    ('cy', 'f4'), ('radius', 'f4'), ('ex', 'f4'),
                    ('ey', 'f4')], align=True)
goal_dt = np.dtype([('cx', 'f8'), ('cy', 'f8'), ('w', 'f8'), ('h', 'f8'),
                    ('colour', 'i4'), ('expect_block', 'i4')], align=True)
block_dt = np.dtype([('body', 'i4'), ('cgroup', 'i4'), ('shape_type', 'i4'),
                     ('colour', 'i4'), ('role', 'i4'), ('label', 'i4')],
                    align=True)
scene_dt = np.dtype([
    ('task', 'i4'), ('max_steps', 'i4'), ('debug_reward', 'i4'),
    ('n_bodies', 'i4'), ('n_shapes', 'i4'), ('n_cverts', 'i4'),
    ('n_joints', 'i4'), ('n_cgroups', 'i4'), ('n_bpairs', 'i4'),
    ('n_goals', 'i4'), ('n_blocks', 'i4'), ('n_prims', 'i4'),
    ('n_dverts', 'i4'), ('n_labels', 'i4'),
    ('robot_body', 'i4'), ('control_body', 'i4'),
    ('finger_body', 'i4', 2), ('motor_joint', 'i4', 2), ('eye_body', 'i4', 2),
    ('pad_', 'i4', 2),
    ('robot_radius', 'f8'),
    ('bodies', body_dt, MAX_BODIES),
    ('shapes', shape_dt, MAX_SHAPES),
    ('cverts', 'f8', (MAX_CVERTS, 2)),
    ('joints', joint_dt, MAX_JOINTS),
    ('cgroups', cgroup_dt, MAX_CGROUPS),
    ('bpairs', 'u1', (MAX_BPAIRS, 2)),
    ('goals', goal_dt, MAX_GOALS),
    ('blocks', block_dt, MAX_BLOCKS),
    ('prims', prim_dt, MAX_PRIMS),
    ('dverts', 'f4', (MAX_DVERTS, 2)),
], align=True)

config_dt = np.dtype([('device', 'i4'), ('batch', 'i4'), ('n_scenes', 'i4'),
                      ('obs_mode', 'i4'), ('res', 'i4'), ('auto_reset', 'i4'),
                      ('reserved0_', 'i4'), ('reset_seed', 'i4'),
                      ('keep_scene', 'i4'), ('device_sampling', 'i4'),
                      ('reserved_', 'i4', 6)], align=True)

MAX_PLACE_ENTS, MAX_PLACE_BODIES = 16, 6
place_ent_dt = np.dtype([
    ('kind', 'i4'), ('goal', 'i4'), ('n_bodies', 'i4'), ('n_groups', 'i4'),
    ('bodies', 'i4', MAX_PLACE_BODIES), ('groups', 'i4', 4),
    ('rand_pos', 'i4'), ('rand_rot', 'i4'),
    ('pos_limit', 'f8'), ('rot_limit', 'f8'), ('orig', 'f8', 3)], align=True)
place_hw_dt = np.dtype([
    ('goal', 'i4'), ('pad_', 'i4'), ('min_side', 'f8'), ('max_side', 'f8'),
    ('cur_h', 'f8'), ('cur_w', 'f8'), ('linf', 'f8')], align=True)
placement_dt = np.dtype([
    ('n_ents', 'i4'), ('n_hw', 'i4'), ('goal_prims', 'i4', (MAX_GOALS, 2)),
    ('arena', 'f8', 4), ('ents', place_ent_dt, MAX_PLACE_ENTS),
    ('hw', place_hw_dt, MAX_GOALS)], align=True)

state_dt = np.dtype([
    ('n_bodies', 'i4'), ('n_joints', 'i4'), ('n_contacts', 'i4'),
    ('episode_steps', 'i4'), ('scene', 'i4'), ('overflow', 'i4'),
    ('n_cache', 'i4'), ('stamp', 'i4'),
    ('pos', 'f8', (MAX_BODIES, 2)), ('angle', 'f8', MAX_BODIES),
    ('vel', 'f8', (MAX_BODIES, 2)), ('angvel', 'f8', MAX_BODIES),
    ('joint_acc', 'f8', (MAX_JOINTS, 2)),
    ('contact_shapes', 'i4', (32, 2)),
    ('contact_jn', 'f8', 32), ('contact_jt', 'f8', 32),
    ('bias_vel', 'f8', (MAX_BODIES, 2)), ('bias_angvel', 'f8', MAX_BODIES),
    ('cache_shapes', 'i4', (STATE_CACHE, 2)), ('cache_hash', 'u4', STATE_CACHE),
    ('cache_age', 'i4', STATE_CACHE),
    ('cache_jn', 'f8', STATE_CACHE), ('cache_jt', 'f8', STATE_CACHE),
], align=True)


class SceneBuilder:
    """Accumulates one scene in reference insertion order."""

    def __init__(self, task, max_steps, phys_vars, debug_reward=False):
        self.task = task
        self.max_steps = max_steps
        self.debug_reward = bool(debug_reward)
        self.phys_vars = phys_vars
        self.bodies = []
        self.shapes = []
        self.cverts = []
        self.joints = []
        self.cgroups = []
        self.prims = []
        self.dverts = []
        self.goals = []
        self.blocks = []
        self.entities = []  # Entity objects in insertion order (placement)
        self.n_labels = 0
        self.robot = None
        self._star_group = 999  # Entity.generate_group_id, entities.py:60-66

    # -- physics ----------------------------------------------------------
    def add_body(self, mass, moment, pos, angle, kind=BODY_DYNAMIC):
        if kind == BODY_KINEMATIC:
            m_inv = i_inv = 0.0
        else:
            m_inv, i_inv = 1.0 / mass, 1.0 / moment
        self.bodies.append(dict(m_inv=m_inv, i_inv=i_inv,
                                p0=(float(pos[0]), float(pos[1])),
                                a0=float(angle), kind=kind))
        return len(self.bodies) - 1

    def new_group_id(self):
        self._star_group += 1
        return self._star_group

    def add_shapes(self, body, shapes, robot_group=False):
        """Add the shapes of one body as one collision group.  Each shape is
        (kind, verts, radius, friction, group)."""
        shape0 = len(self.shapes)
        for kind, verts, radius, friction, group in shapes:
            vert0 = len(self.cverts)
            self.cverts.extend((float(x), float(y)) for x, y in verts)
            self.shapes.append(dict(kind=kind, body=body, vert0=vert0,
                                    nvert=len(verts), radius=float(radius),
                                    friction=float(friction), group=group))
        self.cgroups.append(dict(shape0=shape0, nshape=len(shapes), body=body,
                                 robot_group=int(robot_group)))
        return len(self.cgroups) - 1

    def add_joint(self, kind, a, b, anchor_a=(0.0, 0.0), anchor_b=(0.0, 0.0),
                  p0=0.0, p1=0.0, p2=0.0, max_force=INF, max_bias=INF,
                  error_bias=None):
        if error_bias is None:
            error_bias = pow(1.0 - 0.1, 60.0)  # Chipmunk default
        self.joints.append(dict(kind=kind, a=a, b=b, anchor_a=anchor_a,
                                anchor_b=anchor_b, p0=p0, p1=p1, p2=p2,
                                max_force=max_force, max_bias=max_bias,
                                error_bias=error_bias))
        return len(self.joints) - 1

    # -- drawing ----------------------------------------------------------
    def add_poly_prim(self, verts, rgb_u8, body=None):
        vert0 = len(self.dverts)
        self.dverts.extend((float(x), float(y)) for x, y in verts)
        self.prims.append(dict(
            kind=PRIM_POLY,
            xform=XFORM_WORLD if body is None else XFORM_BODY,
            body=0 if body is None else body, body2=0, rgb=rgb_u8,
            nvert=len(verts), vert0=vert0, stipple=0xFFFF))

    def add_ngon_prim(self, n, radius, rgb_u8, body, centre=(0.0, 0.0),
                      pupil_eye=None, pre_offset=(0.0, 0.0)):
        self.prims.append(dict(
            kind=PRIM_NGON,
            xform=XFORM_BODY if pupil_eye is None else XFORM_PUPIL,
            body=body, body2=0 if pupil_eye is None else pupil_eye,
            rgb=rgb_u8, nvert=n, vert0=0, stipple=0xFFFF, cx=centre[0],
            cy=centre[1], radius=radius, ex=pre_offset[0], ey=pre_offset[1]))

    def add_lineloop_prim(self, verts, rgb_u8, width_px, stipple=0xFFFF):
        vert0 = len(self.dverts)
        self.dverts.extend((float(x), float(y)) for x, y in verts)
        self.prims.append(dict(kind=PRIM_LINELOOP, xform=XFORM_WORLD, body=0,
                               body2=0, rgb=rgb_u8, nvert=len(verts),
                               vert0=vert0, stipple=stipple, radius=width_px))

    # -- freeze -----------------------------------------------------------
    def _collision_pairs(self):
        """Canonical candidate list: collision-group pairs i<j in insertion
        order, minus static-static pairs and pairs inside the robot's filter
        group (QueryReject in Chipmunk: same body / same non-zero group)."""
        pairs = []
        for i, gi in enumerate(self.cgroups):
            for j in range(i + 1, len(self.cgroups)):
                gj = self.cgroups[j]
                if gi['body'] < 0 and gj['body'] < 0:
                    continue
                if gi['robot_group'] and gj['robot_group']:
                    continue
                if gi['body'] == gj['body']:
                    continue
                pairs.append((i, j))
        return pairs

    def compile(self):
        rec = np.zeros((), dtype=scene_dt)

        def fill(name, items, limit):
            if len(items) > limit:
                raise ValueError(f'scene has {len(items)} {name} > {limit}')
            rec['n_' + name] = len(items)
            arr = rec[name]
            for idx, item in enumerate(items):
                if isinstance(item, dict):
                    for k, v in item.items():
                        arr[idx][k] = v
                else:
                    arr[idx] = item

        rec['task'] = self.task
        rec['max_steps'] = self.max_steps
        rec['debug_reward'] = int(self.debug_reward)
        fill('bodies', self.bodies, MAX_BODIES)
        fill('shapes', self.shapes, MAX_SHAPES)
        fill('cverts', self.cverts, MAX_CVERTS)
        fill('joints', self.joints, MAX_JOINTS)
        fill('cgroups', self.cgroups, MAX_CGROUPS)
        fill('bpairs', self._collision_pairs(), MAX_BPAIRS)
        fill('goals', self.goals, MAX_GOALS)
        fill('blocks', self.blocks, MAX_BLOCKS)
        fill('prims', self.prims, MAX_PRIMS)
        fill('dverts', self.dverts, MAX_DVERTS)
        rec['n_labels'] = self.n_labels
        assert self.robot is not None, 'scene has no robot'
        for k, v in self.robot.items():
            rec[k] = v
        return rec
