"""mannequin2_b200 -- B200-native implementation of mannequin2's training hot path.

Same top-level names as mannequin/__init__.py:2-7.  All arithmetic runs in the
sm_100a CUDA library libmq_b200.so (csrc/); there is no CPU fallback.
"""
from ._utils import endswith, discounted, bar, plot_function_2d
from ._normalize import RunningMean, RunningNormalize
from ._trajectory import Trajectory
from ._adam import Adam, Adams

version = (2, 7, 2)
