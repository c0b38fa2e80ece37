"""Registry surface (mirrors the reference's tests/test_rollout_preproc.py
shape checks) and the C-ABI library: loads, exports every declared symbol,
struct layouts agree.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

import magical_b200 as magical
from magical_b200 import _native
from magical_b200 import scene as sc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_registered_envs():
    assert magical.register_envs() in (True, False)
    assert magical.register_envs() is False  # idempotent
    names = magical.ALL_REGISTERED_ENVS
    assert len(names) == 366 and len(set(names)) == 366
    assert len(names) > 8
    # 60 base ids x (1 + 5 preprocessors) + 6 debug-reward ids
    assert sum(1 for n in names if 'DebugReward' in n) == 6
    assert magical.AVAILABLE_PREPROCESSORS == ['LoRes3EA', 'LoRes4E', 'LoRes4A', 'LoResStack', 'LoResCHW4E']
    assert names[0] == 'ClusterShape-Demo-v0'
    assert 'MatchRegions-TestAll-LoResStack-v0' in names
    assert 'ClusterColour-Demo-LoRes4E-v0' in names


def test_demo_to_test_map_and_env_name():
    magical.register_envs()
    m = magical.DEMO_ENVS_TO_TEST_ENVS_MAP
    assert len(m) == 48  # 8 tasks x (plain + 5 preprocessors)
    assert m['MoveToCorner-Demo-v0'] == (
        'MoveToCorner-TestColour-v0', 'MoveToCorner-TestShape-v0', 'MoveToCorner-TestJitter-v0',
        'MoveToCorner-TestDynamics-v0', 'MoveToCorner-TestAll-v0')
    assert len(m['ClusterColour-Demo-LoRes4E-v0']) == 7
    e = magical.EnvName('MatchRegions-TestAll-LoResStack-v0')
    assert (e.task, e.variant, e.preproc, e.version) == ('MatchRegions', 'TestAll', 'LoResStack', 'v0')
    assert e.is_test and e.demo_env_name == 'MatchRegions-Demo-LoResStack-v0'
    assert magical.update_magical_env_name('MoveToCorner-Demo-v0', preproc='LoRes4E') == 'MoveToCorner-Demo-LoRes4E-v0'
    with pytest.raises(ValueError):
        magical.EnvName('NotAnEnv')


def test_episode_lengths_and_kwargs():
    from magical_b200.benchmarks import ENV_SPECS
    magical.register_envs()
    want = {'MoveToCorner': 80, 'MoveToRegion': 40, 'MatchRegions': 120, 'MakeLine': 180,
            'FindDupe': 100, 'FixColour': 60, 'ClusterColour': 240, 'ClusterShape': 240}
    for name, spec in ENV_SPECS.items():
        assert spec.max_episode_steps == want[magical.EnvName(name).task]
        assert spec.kwargs['fps'] == 8 and spec.kwargs['phys_iter'] == 10
        assert spec.kwargs['res_hw'] == (384, 384)
    # the reference registers the DebugReward "-<Preproc>" ids with the plain entry point
    assert ENV_SPECS['MoveToCorner-Demo-DebugReward-LoRes4E-v0'].preproc is None
    assert ENV_SPECS['MoveToCorner-Demo-DebugReward-v0'].kwargs['debug_reward'] is True
    # every registered id builds a scene (Demo + randomised non-layout variants)
    for name in ['FixColour-TestColour-v0', 'ClusterShape-TestShape-LoRes4A-v0', 'FindDupe-TestDynamics-v0']:
        task, _ = magical.make_task(name)
        task.seed(7)
        assert int(task.build_scene()['n_bodies']) >= 6


def test_abi_library_loads_and_exports_every_declared_symbol(built):
    header = open(os.path.join(ROOT, 'include', 'magical_b200.h')).read()
    declared = set(re.findall(r'\b(mg_[a-z_]+)\s*\(', header))
    declared -= {'mg_handle'}
    assert declared == set(_native.ABI_SYMBOLS), declared ^ set(_native.ABI_SYMBOLS)
    lib = _native.load()
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.mg_version() == sc.ABI_VERSION
    assert lib.mg_sizeof_scene() == sc.scene_dt.itemsize
    assert lib.mg_sizeof_state() == sc.state_dt.itemsize
    # argument validation that needs no GPU
    assert lib.mg_bind_obs(None, None, 0) < 0
    assert b'null' in lib.mg_last_error()


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under magical_b200/ may include, import, link or
    load it (comments that cite it are fine)."""
    pkg = os.path.join(ROOT, 'magical_b200')
    bad = re.compile(r'#\s*include\s*"[^"]*(oracle|mgo)[^"]*"|^\s*(from|import)\s+\S*oracle|libmgo_oracle|oracle_lib',
                     re.MULTILINE)
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, fn), errors='replace').read()
                assert not bad.search(text), fn


def test_vec_env_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(_native.NativeError):
        magical.make_vec('MoveToCorner-Demo-LoRes4E-v0', batch=2)


def test_register_envs_registers_with_real_gym_when_importable(monkeypatch):
    """VERDICT r1 item 5 / reference benchmarks/__init__.py:976-999: with `gym`
    importable every id is registered there, so `gym.make(id)` works.  gym is
    not in the image, so a stand-in module with gym's `register` signature is
    injected; the entry point must resolve to the GPU-backed `MagicalEnv`."""
    import importlib
    import sys
    import types
    from magical_b200 import benchmarks
    calls = []
    fake = types.ModuleType('gym')
    fake.register = lambda id, entry_point=None, max_episode_steps=None, kwargs=None, **kw: \
        calls.append((id, entry_point, max_episode_steps, kwargs))
    monkeypatch.setitem(sys.modules, 'gym', fake)
    saved = (benchmarks._REGISTERED, list(benchmarks.ALL_REGISTERED_ENVS),
             dict(benchmarks.ENV_SPECS), dict(benchmarks.DEMO_ENVS_TO_TEST_ENVS_MAP))
    try:
        benchmarks._REGISTERED = False
        benchmarks.ALL_REGISTERED_ENVS.clear()
        benchmarks.ENV_SPECS.clear()
        benchmarks.DEMO_ENVS_TO_TEST_ENVS_MAP.clear()
        assert benchmarks.register_envs() is True
        assert len(calls) == 366
        ids = [c[0] for c in calls]
        assert ids == benchmarks.ALL_REGISTERED_ENVS
        by_id = {c[0]: c for c in calls}
        _, ep, steps, kwargs = by_id['ClusterColour-Demo-LoRes4E-v0']
        assert steps == 240 and kwargs == {'env_id': 'ClusterColour-Demo-LoRes4E-v0'}
        mod, cls = ep.split(':')
        assert getattr(importlib.import_module(mod), cls) is magical.MagicalEnv
        assert by_id['MoveToRegion-Demo-v0'][2] == 40
    finally:
        benchmarks._REGISTERED = saved[0]
        benchmarks.ALL_REGISTERED_ENVS[:] = saved[1]
        benchmarks.ENV_SPECS.clear(); benchmarks.ENV_SPECS.update(saved[2])
        benchmarks.DEMO_ENVS_TO_TEST_ENVS_MAP.clear(); benchmarks.DEMO_ENVS_TO_TEST_ENVS_MAP.update(saved[3])
