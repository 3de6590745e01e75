"""mannequin2_b200 -- B200-native implementation of mannequin2's training hot path.

The package exports the top-level names of the reference package (mannequin/__init__.py:2-7), so that
`from mannequin import Adam, bar, RunningNormalize, Trajectory` keeps working through
`mannequin2_b200.compat`.  All arithmetic runs in the sm_100a CUDA library libmq_b200.so (csrc/); there is
no CPU fallback.  Fused drivers live in the submodules ppo, mlp, es, vae, rollout.
"""
from . import _adam as _a
from . import _normalize as _n
from . import _trajectory as _t
from . import _utils as _u

Adam, Adams = _a.Adam, _a.Adams                                   # optimisers (ascent, _adam.py)
RunningMean, RunningNormalize = _n.RunningMean, _n.RunningNormalize
Trajectory = _t.Trajectory
endswith, discounted = _u.endswith, _u.discounted
bar = _u.bar

__all__ = ["Adam", "Adams", "RunningMean", "RunningNormalize", "Trajectory", "endswith", "discounted", "bar", "version"]

version = (2, 7, 2)                                               # the reference version this API mirrors
