"""Physics variables (joint force limits) and their randomisation ranges.

Mirror of reference `magical/phys_vars.py:70-103` and
`magical/base_env.py:49-57`, without the metaclass: the five variables are a
fixed, ordered tuple because the compiled scene tables index them by position.
"""
import collections


class PhysVar:
    """Default + uniform sampling bounds for one scalar (phys_vars.py:90-103)."""
    def __init__(self, default, bounds):
        lower, upper = bounds
        assert lower <= default <= upper, (lower, default, upper)
        self.default = default
        self.lower = lower
        self.upper = upper

    def sample(self, rng):
        return rng.uniform(self.lower, self.upper)


class PhysicsVariables:
    """Values for one environment instance; build with defaults() / sample()."""
    variables = collections.OrderedDict([
        # order == reference class-body order == RandomState draw order
        ('robot_pos_joint_max_force', PhysVar(3, (2.2, 3.5))),
        ('robot_rot_joint_max_force', PhysVar(1, (0.7, 1.5))),
        ('robot_finger_max_force', PhysVar(4, (2.5, 4.5))),
        ('shape_trans_joint_max_force', PhysVar(1.5, (1.0, 1.8))),
        ('shape_rot_joint_max_force', PhysVar(0.1, (0.07, 0.15))),
    ])

    def __init__(self, *, _var_values):
        if _var_values.keys() != self.variables.keys():
            raise ValueError("must supply all & only given variable names")
        for k, v in _var_values.items():
            setattr(self, k, float(v))

    @classmethod
    def defaults(cls):
        return cls(_var_values={k: v.default
                                for k, v in cls.variables.items()})

    @classmethod
    def sample(cls, rng):
        return cls(_var_values={k: v.sample(rng)
                                for k, v in cls.variables.items()})

    def as_tuple(self):
        return tuple(getattr(self, k) for k in self.variables)

    def __repr__(self):
        pairs = ', '.join(f'{k}={getattr(self, k)}' for k in self.variables)
        return f'{type(self).__name__}({pairs})'
