"""CPU tests of host-side logic: Trajectory container semantics (tests/trajectory.py + .out),
bar (tests/bar.py + bar.out), endswith, sharding helpers."""
import io
from contextlib import redirect_stdout

import numpy as np
import pytest

from conftest import load_golden


def test_trajectory_py_golden():
    from mannequin2_b200 import Trajectory
    old = np.get_printoptions()
    np.set_printoptions(precision=3, linewidth=100, suppress=True, legacy="1.13")
    out = io.StringIO()
    try:
        with redirect_stdout(out):
            t = Trajectory([[1, 2], [3, 4], [5, 6]], [[10], [11], [12]], [20, 21, 22])
            print(t); print(t[1:]); print(t[1])
            t = t[::-1]
            print(t.o); print(t.a); print(t.r)
            t = t.joined(t[0])
            print(t.r)
            t = t.modified(rewards=sum)
            print(t.r)
    finally:
        np.set_printoptions(**old)
    want = """<trajectory: 3 x (2,) -> (1,)>
<trajectory: 2 x (2,) -> (1,)>
<trajectory: 1 x (2,) -> (1,)>
[[ 5.  6.]
 [ 3.  4.]
 [ 1.  2.]]
[[ 12.]
 [ 11.]
 [ 10.]]
[ 22.  21.  20.]
[ 22.  21.  20.  22.]
[ 85.  85.  85.  85.]"""
    norm = lambda s: [" ".join(l.split()) for l in s.strip().splitlines()]
    assert norm(out.getvalue()) == norm(want)


def test_trajectory_rules_vs_reference_fixture():
    from mannequin2_b200 import Trajectory
    g = load_golden("trajectory")
    t = Trajectory(g["traj_o"], g["traj_a"], g["traj_r"])
    b = t[g["traj_idx"]]
    assert np.array_equal(b.o, g["traj_take_o"]) and b.o.dtype == np.float32
    assert np.array_equal(b.a, g["traj_take_a"]) and np.array_equal(b.r, g["traj_take_r"])
    s = t[5:33:3]
    assert np.array_equal(s.o, g["traj_slice_o"]) and np.array_equal(s.r, g["traj_slice_r"])
    ti = Trajectory([3, 1, 2], [[1], [0], [1]])
    assert ti.o.dtype == g["traj_int_o"].dtype and ti.a.dtype == g["traj_int_a"].dtype
    assert np.array_equal(ti.r, np.ones(3, np.float32))
    assert np.array_equal(Trajectory.joined(t[:3], t[10:12]).r, g["traj_join_r"])
    assert len(t) == 40 and t.get_length() == 40
    with pytest.raises((IndexError, ValueError)):      # the reference indexes [0] before its length check
        Trajectory([], [])
    with pytest.raises(ValueError):
        Trajectory([[1.0]], [[1.0], [2.0]])
    with pytest.raises(ValueError):
        t.rewards = 3
    with pytest.raises(ValueError):
        t.o[0, 0] = 1.0                                   # arrays are read-only


def test_bar_golden():
    from mannequin2_b200 import bar, endswith
    # /root/reference/tests/bar.out (first and a negative case)
    assert bar(50.0, 100.0, length=10) == "  50.00 [" + chr(0x2588) * 5 + " " * 5 + "]"
    assert bar(-100.0, 100.0, length=4) == "-100.00 [" + chr(0x2501) * 4 + "]"
    assert len(bar(12.3)) == len("-100.00 ") + 52
    assert endswith((4, 5, 6), (5, 6)) and not endswith((5,), (4, 5)) and endswith((1, 2), ())


def test_shard_ranges_cover_everything():
    from mannequin2_b200.dist import shard_range
    for n, w in ((4096, 8), (10, 3), (7, 8), (1 << 20, 8)):
        parts = [shard_range(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        assert max(hi - lo for lo, hi in parts) - min(hi - lo for lo, hi in parts) <= 1


def test_gym_helpers_host_side():
    """mannequin/gym.py:86-137 without gym: one_step / episode bookkeeping (next_obs carry-over, done handling),
    PrintRewards' running mean of finished episodes and the reward-log line of benchmarks/_env.py:60-63.  Host logic
    only: no device is touched (the observation normaliser is tested on the GPU)."""
    from mannequin2_b200.gym import PrintRewards, episode, log_line, one_step

    class Counter(object):
        def __init__(self):
            self.t = 0

        def reset(self):
            self.t = 0
            return np.array([0.0, 1.0])

        def step(self, action):
            self.t += 1
            return np.array([float(self.t), 1.0]), 2.0, self.t >= 5, {}

    lines = []
    env = PrintRewards(Counter(), every=4, print=lambda s, r: lines.append(log_line(s, r)))
    traj = episode(env, lambda obs: obs[:1] * 0.5)
    assert len(traj) == 5 and np.array_equal(traj.o[:, 0], np.arange(5.0)) and np.all(traj.r == 2.0)
    assert env.next_obs is None                                   # the episode ended: the next one_step resets
    obs, act, rew, done = one_step(env, lambda obs: obs[:1])
    assert obs[0] == 0.0 and rew == 2.0 and not done and env.next_obs[0] == 1.0
    for _ in range(6):
        one_step(env, lambda obs: obs[:1])
    assert lines == ["4 0.00\n", "8 10.00\n", "12 10.00\n"]
    assert log_line(1000, -3.14159) == "1000 -3.14\n"
