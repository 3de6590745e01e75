"""Run unmodified mannequin2 scripts on this package:

    import mannequin2_b200.compat      # registers the alias
    from mannequin import Adam, ...    # now resolves to mannequin2_b200

Only the module names the reference scripts import are aliased
(mnist/mlp.py:7-8, benchmarks/ppo.py:8, benchmarks/policy.py:9-10).
"""
import sys

import mannequin2_b200 as _pkg
from mannequin2_b200 import _adam, _normalize, _trajectory, _utils, backprop, basicnet, distrib

sys.modules.setdefault("mannequin", _pkg)
for _name, _mod in (("basicnet", basicnet), ("backprop", backprop), ("distrib", distrib),
                    ("_adam", _adam), ("_normalize", _normalize), ("_trajectory", _trajectory),
                    ("_utils", _utils)):
    sys.modules.setdefault("mannequin." + _name, _mod)


def _alias_gym():
    # mannequin.gym (one_step, episode, NormalizedObservations, PrintRewards) needs a CUDA device behind its normaliser:
    # aliased lazily so that importing compat on a machine without one still works for everything else
    from mannequin2_b200 import gym as _gym
    sys.modules.setdefault("mannequin.gym", _gym)


try:
    _alias_gym()
except Exception:                                # pragma: no cover
    pass
