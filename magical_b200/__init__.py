"""magical_b200: B200-native batched implementation of the MAGICAL
(qxcv/magical) physics + render hot path, behind the reference's own
`register_envs()` / env-id surface.  See DESIGN.md and INTEGRATION.md.

    import magical_b200 as magical
    magical.register_envs()
    venv = magical.make_vec('ClusterColour-Demo-LoRes4E-v0', batch=65536)
    obs = venv.reset()                      # torch.uint8 [B, 96, 96, 12] on the GPU
    obs, rew, done, info = venv.step(actions)
"""
from magical_b200.benchmarks import (  # noqa: F401
    ALL_REGISTERED_ENVS, AVAILABLE_PREPROCESSORS, DEMO_ENVS_TO_TEST_ENVS_MAP,
    EnvName, register_envs, update_magical_env_name)
from magical_b200.env import (MagicalEnv, make, make_task, make_vec,  # noqa: F401
                              make_vec_mixed)
from magical_b200.vec_env import MagicalVecEnv  # noqa: F401

__version__ = '0.1.0'
