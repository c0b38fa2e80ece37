#!/usr/bin/env python3
"""Record a ground-truth trace from the REAL reference (qxcv/magical on pymunk 5.6 / pyglet / gym 0.17) so the
oracle's "parity unpinned" status can be lifted on any machine that has those packages (none does in the build
or GPU containers of this project: SURVEY.md §8c).

    xvfb-run -a python tools/dump_pymunk_trace.py MoveToRegion-Demo-v0 --steps 200 --seed 7 -o trace.json
    python tools/dump_pymunk_trace.py --compare trace.json          # needs only this repo

The trace holds, per env-step: the action, the pose (x, y, angle) and velocity of the six robot bodies and every
block in the reference's entity order, `done` and `eval_score`.  `--compare` replays the actions through the
CPU oracle (oracle/, libm sin/cos) and prints the largest pose difference per step; see DESIGN.md §4 for why the
zero-length finger PinJoints make anything beyond the first few steps sensitive to 1-ulp differences.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def record(env_id, steps, seed, out):
    try:
        import gym
        import magical  # the reference package
    except Exception as ex:  # noqa: BLE001
        raise SystemExit(f'the reference stack is not importable here ({ex}); run this where '
                         'pymunk~=5.6, pyglet==1.5.*, gym==0.17.* and magical are installed')
    magical.register_envs()
    env = gym.make(env_id)
    env.seed(seed)
    env.action_space.seed(seed + 35)
    env.reset()
    base = env.unwrapped
    robot = base._robot
    bodies = [robot.robot_body, robot.control_body, *robot.pupil_bodies, *robot.finger_bodies]
    for ent in base._entities:
        if hasattr(ent, 'shape_body'):
            bodies.append(ent.shape_body)
    trace = {'env_id': env_id, 'seed': seed, 'steps': []}
    for _ in range(steps):
        action = int(env.action_space.sample())
        _, rew, done, info = base.step(action)  # unwrapped: keeps stepping past the time limit
        trace['steps'].append({
            'action': action, 'done': bool(done), 'reward': float(rew),
            'eval_score': float(info['eval_score']),
            'pose': [[float(b.position.x), float(b.position.y), float(b.angle)] for b in bodies],
            'vel': [[float(b.velocity.x), float(b.velocity.y), float(b.angular_velocity)] for b in bodies]})
    with open(out, 'w') as fh:
        json.dump(trace, fh)
    print(f'wrote {out}: {steps} steps of {env_id}')


def compare(path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import numpy as np
    import magical_b200 as magical
    from oracle_lib import OracleEnv
    with open(path) as fh:
        trace = json.load(fh)
    task, _ = magical.make_task(trace['env_id'])
    rec = task.build_scene()
    orc = OracleEnv(rec, det_sincos=False)
    # reference body order -> scene body indices
    order = [int(rec['robot_body']), int(rec['control_body']), *map(int, rec['eye_body']),
             *map(int, rec['finger_body']), *[int(b) for b in rec['blocks']['body'][:rec['n_blocks']]]]
    worst = 0.0
    for t, step in enumerate(trace['steps']):
        _, done, score = orc.step(step['action'])
        st = orc.state()
        ref = np.asarray(step['pose'])
        got = np.stack([np.append(st['pos'][b], st['angle'][b]) for b in order[:len(ref)]])
        d = float(np.abs(got - ref).max())
        worst = max(worst, d)
        flag = '' if bool(done) == step['done'] else '  DONE MISMATCH'
        print(f'step {t:4d}  max |pose - pymunk| = {d:.3e}{flag}')
    print(f'worst over {len(trace["steps"])} steps: {worst:.3e}')


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('env_id', nargs='?', default='MoveToRegion-Demo-v0')
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--seed', type=int, default=7)
    ap.add_argument('-o', '--out', default='pymunk_trace.json')
    ap.add_argument('--compare', metavar='TRACE')
    args = ap.parse_args()
    if args.compare:
        compare(args.compare)
    else:
        record(args.env_id, args.steps, args.seed, args.out)
