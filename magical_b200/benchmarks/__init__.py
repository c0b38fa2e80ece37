"""Environment registry: the `<Task>-<Variant>[-<Preproc>]-v0` ids, the
demo->test map and the preprocessor table of the reference
(`magical/benchmarks/__init__.py:242-1049`), bound to the B200 vector env.

The reference implements each preprocessor as a stack of gym wrappers around a
CPU env (`FlattenFrameStack`, `EagerDictFrameStack`, `ResizeDictObservation`,
`ChannelsFirst`, :46-192); here a preprocessor is an observation layout the
rasteriser writes directly (scene.OBS_*), so no wrapper objects exist.
"""
import collections
import re

from magical_b200.benchmarks.cluster import ClusterColourEnv, ClusterShapeEnv
from magical_b200.benchmarks.find_dupe import FindDupeEnv
from magical_b200.benchmarks.fix_colour import FixColourEnv
from magical_b200.benchmarks.make_line import MakeLineEnv
from magical_b200.benchmarks.match_regions import MatchRegionsEnv
from magical_b200.benchmarks.move_to_corner import MoveToCornerEnv
from magical_b200.benchmarks.move_to_region import MoveToRegionEnv

__all__ = [
    'ALL_REGISTERED_ENVS', 'DEMO_ENVS_TO_TEST_ENVS_MAP', 'register_envs',
    'EnvName', 'update_magical_env_name', 'AVAILABLE_PREPROCESSORS',
    'DEFAULT_RES', 'ENV_SPECS',
]

DEFAULT_RES = (384, 384)
# preprocessor name -> (allo frames, ego frames, layout note); order matters:
# it is the registration order of the reference (:242-274)
DEFAULT_PREPROC_ENTRY_POINT_WRAPPERS = collections.OrderedDict([
    ('LoRes3EA', 'allo x1 + ego x3 -> (96, 96, 12)'),
    ('LoRes4E', 'ego x4 -> (96, 96, 12)'),
    ('LoRes4A', 'allo x4 -> (96, 96, 12)'),
    ('LoResStack', "{'allo','ego'} each x4 -> (96, 96, 12)"),
    ('LoResCHW4E', 'ego x4, channels first -> (12, 96, 96)'),
])
AVAILABLE_PREPROCESSORS = list(DEFAULT_PREPROC_ENTRY_POINT_WRAPPERS)
_ENV_NAME_RE = re.compile(
    r'^(?P<name_prefix>[^-]+)(?P<demo_test_spec>-(Demo|Test[^-]*))'
    r'(?P<env_name_suffix>(-[^-]+)*)(?P<version_suffix>-v\d+)$')
_REGISTERED = False
DEMO_ENVS_TO_TEST_ENVS_MAP = collections.OrderedDict()
ALL_REGISTERED_ENVS = []
# env id -> EnvSpec
ENV_SPECS = collections.OrderedDict()

EnvSpec = collections.namedtuple(
    'EnvSpec', ['id', 'entry_point', 'preproc', 'max_episode_steps', 'kwargs'])


class EnvName:
    """Parser for `<name_prefix>-<demo_test_spec>[-<suffix>]-<version>` ids
    (same fields as reference benchmarks/__init__.py:317-391)."""

    def __init__(self, env_name):
        match = _ENV_NAME_RE.match(env_name)
        if match is None:
            raise ValueError(
                f"env name '{env_name}' does not match _ENV_NAME_RE spec")
        groups = match.groupdict()
        self.name_prefix = groups['name_prefix']
        self.demo_test_spec = groups['demo_test_spec']
        self.env_name_suffix = groups['env_name_suffix']
        self.version_suffix = groups['version_suffix']
        assert env_name == self.env_name
        if not self.is_test:
            assert self.demo_env_name == self.env_name

    @property
    def env_name(self):
        return self.name_prefix + self.demo_test_spec \
            + self.env_name_suffix + self.version_suffix

    @property
    def is_test(self):
        return self.demo_test_spec.startswith('-Test')

    @property
    def demo_env_name(self):
        return self.name_prefix + '-Demo' + self.env_name_suffix \
            + self.version_suffix

    @property
    def task(self):
        return self.name_prefix

    @property
    def variant(self):
        return self.demo_test_spec.strip('-')

    @property
    def preproc(self):
        return self.env_name_suffix.strip('-') \
            if self.env_name_suffix else None

    @property
    def version(self):
        return self.version_suffix.strip('-')


def update_magical_env_name(env_name, *, task=None, variant=None,
                            preproc=None, version=None):
    ename = EnvName(env_name)
    parts = [task if task is not None else ename.task,
             variant if variant is not None else ename.variant]
    if preproc is None:
        preproc = ename.preproc
    if preproc is not None:
        parts.append(preproc)
    parts.append(version if version is not None else ename.version)
    return '-'.join(parts)


def _variants(flag_names, table):
    """Expand {variant: set of enabled flags} into full kwargs dicts."""
    out = []
    for variant, enabled in table:
        unknown = set(enabled) - set(flag_names)
        assert not unknown, unknown
        out.append((variant, {f: (f in enabled) for f in flag_names}))
    return out


_CLUSTER_FLAGS = ['rand_shape_colour', 'rand_shape_type', 'rand_layout_minor',
                  'rand_layout_full', 'rand_shape_count', 'rand_dynamics']
_CLUSTER_TABLE = [
    ('Demo', []),
    ('TestJitter', ['rand_layout_minor']),
    ('TestColour', ['rand_shape_colour']),
    ('TestShape', ['rand_shape_type']),
    ('TestLayout', ['rand_layout_full']),
    ('TestCountPlus', ['rand_shape_colour', 'rand_shape_type',
                       'rand_layout_full', 'rand_shape_count']),
    ('TestDynamics', ['rand_dynamics']),
    ('TestAll', ['rand_shape_colour', 'rand_shape_type', 'rand_layout_full',
                 'rand_shape_count', 'rand_dynamics']),
]
_BLOCKS_FLAGS = ['rand_colours', 'rand_shapes', 'rand_count',
                 'rand_layout_minor', 'rand_layout_full', 'rand_dynamics']
_BLOCKS_TABLE = [
    ('Demo', []),
    ('TestJitter', ['rand_layout_minor']),
    ('TestColour', ['rand_colours']),
    ('TestShape', ['rand_shapes']),
    ('TestLayout', ['rand_layout_full']),
    ('TestCountPlus', ['rand_colours', 'rand_shapes', 'rand_count',
                       'rand_layout_full']),
    ('TestDynamics', ['rand_dynamics']),
    ('TestAll', ['rand_colours', 'rand_shapes', 'rand_count',
                 'rand_layout_full', 'rand_dynamics']),
]
_MR_FLAGS = ['rand_target_colour', 'rand_shape_type', 'rand_shape_count',
             'rand_layout_minor', 'rand_layout_full', 'rand_dynamics']
_MR_TABLE = [
    ('Demo', []),
    ('TestJitter', ['rand_layout_minor']),
    ('TestColour', ['rand_target_colour']),
    ('TestShape', ['rand_shape_type']),
    ('TestLayout', ['rand_layout_full']),
    ('TestCountPlus', ['rand_target_colour', 'rand_shape_type',
                       'rand_shape_count', 'rand_layout_full']),
    ('TestDynamics', ['rand_dynamics']),
    ('TestAll', ['rand_target_colour', 'rand_shape_type', 'rand_shape_count',
                 'rand_layout_full', 'rand_dynamics']),
]
_MTC_FLAGS = ['rand_shape_colour', 'rand_shape_type', 'rand_poses',
              'rand_dynamics']
_MTC_TABLE = [
    ('Demo', []),
    ('TestColour', ['rand_shape_colour']),
    ('TestShape', ['rand_shape_type']),
    ('TestJitter', ['rand_poses']),
    ('TestDynamics', ['rand_dynamics']),
    ('TestAll', _MTC_FLAGS),
]
_MTR_FLAGS = ['rand_poses_minor', 'rand_poses_full', 'rand_goal_colour',
              'rand_dynamics']
_MTR_TABLE = [
    ('Demo', []),
    ('TestJitter', ['rand_poses_minor']),
    ('TestColour', ['rand_goal_colour']),
    ('TestLayout', ['rand_poses_full']),
    ('TestDynamics', ['rand_dynamics']),
    ('TestAll', ['rand_poses_full', 'rand_goal_colour', 'rand_dynamics']),
]
# (task name, entry point, episode length, flags, variant table); episode
# lengths from reference benchmarks/__init__.py:407,453,503,580,657,733,813;
# the list order is the reference's registration order (:964-972)
_TASKS = [
    ('ClusterShape', ClusterShapeEnv, 240, _CLUSTER_FLAGS, _CLUSTER_TABLE),
    ('ClusterColour', ClusterColourEnv, 240, _CLUSTER_FLAGS, _CLUSTER_TABLE),
    ('FindDupe', FindDupeEnv, 100, _BLOCKS_FLAGS, _BLOCKS_TABLE),
    ('FixColour', FixColourEnv, 60, _BLOCKS_FLAGS, _BLOCKS_TABLE),
    ('MakeLine', MakeLineEnv, 180, _BLOCKS_FLAGS, _BLOCKS_TABLE),
    ('MatchRegions', MatchRegionsEnv, 120, _MR_FLAGS, _MR_TABLE),
    ('MoveToCorner', MoveToCornerEnv, 80, _MTC_FLAGS, _MTC_TABLE),
    ('MoveToRegion', MoveToRegionEnv, 40, _MTR_FLAGS, _MTR_TABLE),
]


def _gym_module():
    """The real `gym` package when it is importable (it is not part of the
    build image; the reference pins gym==0.17.*), else None."""
    try:
        import gym
        return gym if hasattr(gym, 'register') else None
    except Exception:  # noqa: BLE001
        return None


def _register(env_id, entry_point, preproc, ep_len, kwargs):
    ALL_REGISTERED_ENVS.append(env_id)
    ENV_SPECS[env_id] = EnvSpec(env_id, entry_point, preproc, ep_len, kwargs)
    # the upper surface is kept verbatim (reference benchmarks/__init__.py:976-999):
    # with gym present, `gym.make(env_id)` builds the GPU-backed env, wrapped in
    # gym's own TimeLimit by `max_episode_steps` exactly like the reference's ids
    gym = _gym_module()
    if gym is not None:
        gym.register(env_id, entry_point='magical_b200.env:MagicalEnv',
                     max_episode_steps=ep_len, kwargs={'env_id': env_id})


def register_envs():
    """Register all default environments; idempotent, returns False when they
    were already registered (reference benchmarks/__init__.py:394-399)."""
    global _REGISTERED
    if _REGISTERED:
        return False
    _REGISTERED = True
    common_kwargs = dict(res_hw=DEFAULT_RES, fps=8, phys_steps=10,
                         phys_iter=10)
    for task, cls, ep_len, flags, table in _TASKS:
        for variant, env_kwargs in _variants(flags, table):
            env_name = f'{task}-{variant}-v0'
            kwargs = {'max_episode_steps': ep_len, **common_kwargs,
                      **env_kwargs}
            _register(env_name, cls, None, ep_len, kwargs)
            for preproc in DEFAULT_PREPROC_ENTRY_POINT_WRAPPERS:
                _register(update_magical_env_name(env_name, preproc=preproc),
                          cls, preproc, ep_len, kwargs)

    train_to_test = {}
    demo_envs = set()
    for name in ALL_REGISTERED_ENVS:
        parsed = EnvName(name)
        if parsed.is_test:
            train_to_test.setdefault(parsed.demo_env_name, []).append(
                parsed.env_name)
        else:
            demo_envs.add(parsed.env_name)
    assert demo_envs == train_to_test.keys()
    DEMO_ENVS_TO_TEST_ENVS_MAP.update(
        sorted((k, tuple(v)) for k, v in train_to_test.items()))

    # MoveToCorner with the shaped debugging reward.  As in the reference
    # (:1039-1047) the "-<Preproc>" debug ids are registered with the PLAIN
    # entry point, i.e. they return the raw two-view observation.
    debug_kwargs = {'debug_reward': True, 'max_episode_steps': 80,
                    'rand_shape_colour': False, 'rand_shape_type': False,
                    'rand_poses': False, **common_kwargs}
    _register('MoveToCorner-Demo-DebugReward-v0', MoveToCornerEnv, None, 80,
              debug_kwargs)
    for preproc in DEFAULT_PREPROC_ENTRY_POINT_WRAPPERS:
        _register(f'MoveToCorner-Demo-DebugReward-{preproc}-v0',
                  MoveToCornerEnv, None, 80, debug_kwargs)
    return True
