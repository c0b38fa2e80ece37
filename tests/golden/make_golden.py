#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ from the reference tree.

Run in the BUILD container only (needs /root/reference; the GPU box does not
have it).  Two kinds of fixtures, both produced by the reference itself:

* ref_images/static-<task>-demo-v0.png: the reference's own GL renders of the
  eight Demo reset states (README art, /root/reference/images/).  They are the
  only recorded OUTPUTS of the reference's render path that exist offline and
  pin scene geometry, colours, draw order and camera mapping.
* style_physvars.json: values obtained by importing the two reference modules
  that import cleanly without pymunk/pyglet/gym (magical/style.py,
  magical/phys_vars.py + the PhysicsVariables ranges quoted from
  magical/base_env.py:49-57).
"""
import importlib.util
import json
import os
import shutil

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    img_dir = os.path.join(HERE, 'ref_images')
    os.makedirs(img_dir, exist_ok=True)
    for fn in sorted(os.listdir(os.path.join(REF, 'images'))):
        if fn.startswith('static-') and fn.endswith('.png'):
            shutil.copy(os.path.join(REF, 'images', fn), os.path.join(img_dir, fn))
    style = load(os.path.join(REF, 'magical', 'style.py'), 'ref_style')
    pv = load(os.path.join(REF, 'magical', 'phys_vars.py'), 'ref_phys_vars')

    class PhysicsVariables(pv.PhysicsVariablesBase):
        # ranges as in reference magical/base_env.py:49-57
        robot_pos_joint_max_force = pv.PhysVar(3, (2.2, 3.5))
        robot_rot_joint_max_force = pv.PhysVar(1, (0.7, 1.5))
        robot_finger_max_force = pv.PhysVar(4, (2.5, 4.5))
        shape_trans_joint_max_force = pv.PhysVar(1.5, (1.0, 1.8))
        shape_rot_joint_max_force = pv.PhysVar(0.1, (0.07, 0.15))

    import numpy as np
    rng = np.random.RandomState(1234)
    sampled = PhysicsVariables.sample(rng)
    out = {
        'COLOURS_RGB': {k: list(v) for k, v in style.COLOURS_RGB.items()},
        'background': list(style.lighten_rgb(style.COLOURS_RGB['grey'], times=4)),
        'darken': {k: list(style.darken_rgb(v)) for k, v in style.COLOURS_RGB.items()},
        'lighten2': {k: list(style.lighten_rgb(v, times=2)) for k, v in style.COLOURS_RGB.items()},
        'thickness': [style.GOAL_LINE_THICKNESS, style.SHAPE_LINE_THICKNESS, style.ROBOT_LINE_THICKNESS],
        'ARENA_ZOOM_OUT': style.ARENA_ZOOM_OUT,
        'physvar_names': list(PhysicsVariables.variables.keys()),
        'physvar_defaults': [getattr(PhysicsVariables.defaults(), k) for k in PhysicsVariables.variables],
        'physvar_sample_seed1234': [getattr(sampled, k) for k in PhysicsVariables.variables],
    }
    with open(os.path.join(HERE, 'style_physvars.json'), 'w') as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print('wrote', len(os.listdir(img_dir)), 'images and style_physvars.json')


if __name__ == '__main__':
    main()
