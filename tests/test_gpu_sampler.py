"""Device-side reset randomisation (csrc/mg_sample.cu; reference magical/geom.py:116-359):
10^5 sampled resets are valid layouts (inside the arena, nothing touching, jitter limits kept), every reset is
a fresh layout, the step after a sampled reset is bit-exact against the oracle stepping the sampled scene, and
the sampler sustains the reset rate BASELINE config 4 needs."""
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _world_shapes(rec):
    """WorldShape lists per collision group of a compiled scene record at its reset poses, plus the goal
    sensor boxes (host statement of what may not touch: tests/ only)."""
    import math
    from magical_b200.placement import WorldShape
    groups = []
    for g in range(int(rec['n_cgroups'])):
        grp = rec['cgroups'][g]
        shapes = []
        for si in range(int(grp['shape0']), int(grp['shape0']) + int(grp['nshape'])):
            sh = rec['shapes'][si]
            verts = [tuple(v) for v in rec['cverts'][int(sh['vert0']):int(sh['vert0']) + int(sh['nvert'])]]
            b = int(sh['body'])
            if b >= 0:
                c, s = math.cos(rec['bodies'][b]['a0']), math.sin(rec['bodies'][b]['a0'])
                px, py = rec['bodies'][b]['p0']
                verts = [(px + v[0] * c - v[1] * s, py + v[0] * s + v[1] * c) for v in verts]
            shapes.append(WorldShape(verts, float(sh['radius']), int(sh['group'])))
        groups.append((int(grp['body']), shapes))
    goals = []
    for g in range(int(rec['n_goals'])):
        cx, cy, w, h = (float(rec['goals'][g][k]) for k in ('cx', 'cy', 'w', 'h'))
        goals.append(WorldShape([(cx - w / 2, cy - h / 2), (cx + w / 2, cy - h / 2), (cx + w / 2, cy + h / 2),
                                 (cx - w / 2, cy + h / 2)], 0.0, 0))
    return groups, goals


def _assert_layout_valid(rec, prog):
    """Exact check of one sampled layout with the host overlap predicate: no entity touches a wall, another
    entity or a goal sensor."""
    from magical_b200.placement import shapes_collide
    groups, goals = _world_shapes(rec)
    ents = []   # list of shape lists, one per entity (bodies moving together = one entity)
    owner = {}
    for k in range(int(prog['n_ents'])):
        e = prog['ents'][k]
        if int(e['kind']) == 1:
            ents.append([goals[int(e['goal'])]])
        else:
            ents.append([s for gi in e['groups'][:int(e['n_groups'])] for s in groups[int(gi)][1]])
            for gi in e['groups'][:int(e['n_groups'])]:
                owner[int(gi)] = k
    fixed = [s for gi, (body, shapes) in enumerate(groups) if gi not in owner for s in shapes]
    listed_goals = {int(prog['ents'][k]['goal']) for k in range(int(prog['n_ents'])) if int(prog['ents'][k]['kind']) == 1}
    fixed += [g for gi, g in enumerate(goals) if gi not in listed_goals]
    for i, a in enumerate(ents):
        for s in a:
            assert not any(shapes_collide(s, o) for o in fixed), ('entity touches a fixed shape', i)
        for j in range(i):
            assert not any(shapes_collide(s, o) for s in a for o in ents[j]), ('entities touch', i, j)


def test_hundred_thousand_sampled_resets_are_valid_and_fresh(built):
    import magical_b200 as magical
    B, rounds = 4096, 25           # 102 400 sampled resets
    venv = magical.make_vec('MatchRegions-TestAll-LoResStack-v0', B, n_scenes=48, seed=5, device_sampling=True,
                            alloc_obs=False)
    all_pos = []
    exact_checked = 0
    for r in range(rounds):
        venv.reset()
        poses = venv.get_poses()
        all_pos.append(poses[:, :, :2].copy())
        if r % 6 == 0:
            for e in range(0, B, 41):   # 100 layouts per checked round, exact host predicate
                rec = venv.get_env_scene(e)
                tmpl = None
                for t in range(venv.n_scenes):   # the template this env drew: same structure
                    if np.array_equal(rec['shapes'], venv.scenes[t]['shapes']) and \
                            np.array_equal(rec['blocks'], venv.scenes[t]['blocks']) and \
                            np.array_equal(rec['joints'], venv.scenes[t]['joints']):
                        tmpl = t
                        break
                assert tmpl is not None
                _assert_layout_valid(rec, venv.programs[tmpl])
                # goal size within the reference's range, sensor inside the arena
                g = rec['goals'][0]
                assert 0.5 <= g['w'] <= 0.8 and 0.5 <= g['h'] <= 0.8
                assert abs(g['cx']) + g['w'] / 2 <= 1.0 + 1e-9 and abs(g['cy']) + g['h'] / 2 <= 1.0 + 1e-9
                # the state was reset from the sampled scene
                nb = int(rec['n_bodies'])
                assert np.array_equal(poses[e, :nb, 0], rec['bodies']['p0'][:nb, 0])
                exact_checked += 1
    assert venv.sampler_failures() == 0
    assert exact_checked >= 400
    pos = np.stack(all_pos).reshape(-1, 16, 2)      # [10^5 layouts, 16 bodies, 2]
    # MatchRegions bodies: blocks first, then the robot's six (body, control, eye, eye, finger, finger); unused
    # slots are exactly zero.  (The eye bodies sit at the world origin in the reference and are shifted rigidly
    # with the robot -- entities.py:267-277, geom.py:362-384 -- so they may leave the arena: only their angle
    # is ever used.)
    used = np.abs(pos).sum(axis=2) > 0
    nb = used.sum(axis=1)
    assert nb.min() >= 7 and nb.max() <= 14
    rows = np.arange(len(pos))
    robot = pos[rows, nb - 6]
    # the robot's circle (r = 0.2) clears the four walls in every one of the 10^5 layouts
    assert np.abs(robot).max() < 0.8
    for f in (1, 2):                                 # finger bodies stay inside the arena
        assert np.abs(pos[rows, nb - f]).max() < 1.0
    # every block centre is inside the arena by more than the smallest block inradius, and further from
    # the robot's centre than robot radius + that inradius
    for k in range(8):
        has = nb - 6 > k
        if not has.any():
            continue
        blk = pos[has, k]
        assert np.abs(blk).max() < 1.0 - 0.04
        assert np.linalg.norm(blk - robot[has], axis=1).min() > 0.2 + 0.04
    # fresh layouts: no robot position ever repeats, and positions spread over the free area
    assert len(np.unique(robot, axis=0)) == len(robot)
    assert np.abs(robot.mean(axis=0)).max() < 0.02 and robot.std(axis=0).min() > 0.3
    venv.close()


def test_jitter_variant_keeps_the_reference_limits(built):
    import magical_b200 as magical
    B = 2048
    venv = magical.make_vec('MoveToRegion-TestJitter-LoRes4E-v0', B, n_scenes=4, seed=2, device_sampling=True,
                            alloc_obs=False)
    prog = venv.programs[0]
    lim_p, lim_r = float(prog['ents'][1]['pos_limit']), float(prog['ents'][1]['rot_limit'])
    assert lim_p == pytest.approx(0.025) and lim_r == pytest.approx(0.05 * np.pi)
    orig = prog['ents'][1]['orig']
    robot = int(prog['ents'][1]['bodies'][0])
    seen = []
    for r in range(5):
        venv.reset()
        p = venv.get_poses()[:, robot]
        assert np.abs(p[:, 0] - orig[0]).max() <= lim_p and np.abs(p[:, 1] - orig[1]).max() <= lim_p
        assert np.abs(p[:, 2] - orig[2]).max() <= lim_r
        seen.append(p.copy())
    p = np.concatenate(seen)
    # the draws fill the allowed box (uniform: std = width / sqrt(12))
    assert p[:, 0].std() == pytest.approx(2 * lim_p / np.sqrt(12), rel=0.05)
    assert p[:, 2].std() == pytest.approx(2 * lim_r / np.sqrt(12), rel=0.05)
    rec = venv.get_env_scene(7)
    g, hw = rec['goals'][0], prog['hw'][0]
    assert abs(g['h'] - hw['cur_h']) <= hw['linf'] + 1e-12 and abs(g['w'] - hw['cur_w']) <= hw['linf'] + 1e-12
    venv.close()


@pytest.mark.parametrize('env_id', ['MatchRegions-TestAll-LoResStack-v0', 'ClusterShape-TestLayout-LoRes4E-v0'])
def test_step_after_sampled_reset_is_bit_exact_vs_oracle(built, env_id):
    """The scene an environment plays after a device-sampled reset (explicit and auto-reset) is handed to the
    oracle: poses, done / score and both rendered views must agree bit for bit over the following steps."""
    import torch
    import magical_b200 as magical
    from oracle_lib import OracleEnv
    B = 48
    venv = magical.make_vec(env_id, B, n_scenes=8, seed=9, device_sampling=True, auto_reset=True)
    obs = venv.reset()
    rng = np.random.RandomState(4)
    watch = [3, 17, 40]

    def fresh_oracles():
        return {e: OracleEnv(venv.get_env_scene(e), det_sincos=True) for e in watch}

    def check_frames(orcs, obs):
        for e, orc in orcs.items():
            if obs.dim() == 5:
                assert np.array_equal(obs[0, e, :, :, 9:].cpu().numpy(), orc.render_lores(0)), e
                assert np.array_equal(obs[1, e, :, :, 9:].cpu().numpy(), orc.render_lores(1)), e
            else:
                assert np.array_equal(obs[e, :, :, 9:].cpu().numpy(), orc.render_lores(1)), e

    orcs = fresh_oracles()
    check_frames(orcs, obs)          # the reset observation shows the sampled layout
    first_scenes = {e: venv.get_env_scene(e) for e in watch}
    T = venv.max_episode_steps
    for t in range(T + 25):
        acts = rng.randint(0, 18, size=B).astype(np.int32)
        obs, rew, done, info = venv.step(torch.from_numpy(acts).cuda())
        if t == T - 1:
            assert bool(done.all())
            for e, orc in orcs.items():
                _, d, s = orc.step(int(acts[e]))
                assert d and np.float32(s) == np.float32(info['eval_score'][e].item())
            # auto-reset: every env plays a NEW sampled layout from now on
            orcs = fresh_oracles()
            for e in watch:
                assert not np.array_equal(first_scenes[e]['bodies']['p0'], venv.get_env_scene(e)['bodies']['p0'])
            check_frames(orcs, obs)
            continue
        for e, orc in orcs.items():
            orc.step(int(acts[e]))
        if t % 10 == 0 or t > T:
            for e, orc in orcs.items():
                st, ost = venv.get_state(e), orc.state()
                nb = int(st['n_bodies'])
                assert np.array_equal(st['pos'][:nb], ost['pos'][:nb]) and np.array_equal(st['angle'][:nb], ost['angle'][:nb]), (e, t)
            check_frames(orcs, obs)
    assert venv.sampler_failures() == 0
    venv.close()


def test_sampler_sustains_config4_reset_rate(built):
    """BASELINE config 4 resets 8192 envs / 120 steps x ~1.5 M env-steps/s = 12 500 layouts/s/GPU."""
    import torch
    import magical_b200 as magical
    B = 8192
    venv = magical.make_vec('MatchRegions-TestAll-LoResStack-v0', B, n_scenes=64, seed=1, device_sampling=True,
                            alloc_obs=False)
    venv.reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        venv.reset()
    torch.cuda.synchronize()
    rate = n * B / (time.perf_counter() - t0)
    print(f'device sampler: {rate:,.0f} sampled resets/s')
    assert rate >= 12500
    assert venv.sampler_failures() == 0
    venv.close()


def test_unsupported_tasks_are_refused(built):
    import magical_b200 as magical
    with pytest.raises(NotImplementedError, match='randomise_pose|ignore_ents'):
        magical.make_vec('FindDupe-TestAll-LoRes4E-v0', 8, n_scenes=2, device_sampling=True)
