"""The measured evidence under profiles/ stays consistent with the tools that made it: `roofline.traffic`
is recomputable from the tracked ncu CSV, and the tracked bench lines carry the keys
of the bench.py contract."""
import importlib.util
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, 'profiles')
ALGORITHMIC_RASTER_BYTES = (110592 + 82944) * 65536   # obs write + surviving-frame read, DESIGN.md section 5


def _load(path):
    return [json.loads(l) for l in open(path) if l.startswith('{')]


def test_traffic_json_is_recomputable_from_the_tracked_csv():
    spec = importlib.util.spec_from_file_location('make_traffic', os.path.join(ROOT, 'tools', 'make_traffic.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    want = mod.compute(os.path.join(PROF, 'r02_dram_bench.csv'))
    have = json.load(open(os.path.join(PROF, 'traffic.json')))
    for k in ('k_raster', 'k_physics_tpe', 'k_finish'):
        assert have[k] == want[k], k
    assert have['_batch'] == 65536 and have['_env_id'] == 'ClusterColour-Demo-LoRes4E-v0'
    # steady-state launches only (the reset renders of the prelude must not dilute the median)
    assert have['_detail']['k_raster']['full_launches'] >= 10
    ratio = have['k_raster'] / ALGORITHMIC_RASTER_BYTES
    assert 1.0 <= ratio <= 1.25, ratio


@pytest.mark.parametrize('name', ['r02_bench_cluster65536.json', 'r02_bench_mtc4096.json',
                                  'r02_bench_mr_testall8192.json', 'r02_bench_2gpu.json',
                                  'r02_bench_4gpu.json', 'r02_bench_8gpu.json'])
def test_tracked_bench_lines_follow_the_contract(name):
    lines = _load(os.path.join(PROF, name))
    assert len(lines) == 1
    d = lines[0]
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
              'dtype', 'data', 'config', 'e2e', 'gpu_launches', 'clocks', 'roofline'):
        assert k in d, k
    assert d['metric'] == 'env_steps_per_s' and d['higher_is_better'] is True and d['data'] == 'synthetic'
    assert d['gpu_launches'] > 0 and d['overflow_envs'] == 0
    assert not set(d['clocks']['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}
    r = d['roofline']
    assert r['bound'] == 'hbm' and r['unit'] == 'GB/s'
    assert abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9
    assert d['e2e']['h2d_bytes_per_step'] > 0 and d['e2e']['d2h_bytes_per_step'] > 0
    if d['n_gpus'] == 1 and name == 'r02_bench_cluster65536.json':
        assert r['kernel'] == 'k_raster' and r['traffic'] == json.load(open(os.path.join(PROF, 'traffic.json')))['k_raster']
        assert d['cpu_baseline']['kind'] in ('port', 'reference') and d['cpu_baseline']['cores'] >= 1
    if d['n_gpus'] > 1:
        assert 'obs' in d['config']['parallelism'] and r['nvlink']['cap_gbs_per_direction'] == 900.0
