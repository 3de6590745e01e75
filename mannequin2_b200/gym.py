"""Environment-side helpers of mannequin/gym.py that the benchmark scripts use, without the `gym`
dependency: any object with `reset() -> obs` and `step(action) -> (obs, reward, done, info)` works.

    one_step / episode       mannequin/gym.py:115-137 (the chunk carry-over `env.next_obs` included)
    NormalizedObservations   mannequin/gym.py:61-84 for ONE environment: the device normaliser of rollout.py with E = 1
                             (for many environments in lock step use rollout.BatchedActor instead)
    PrintRewards             mannequin/gym.py:86-113 + the log line of benchmarks/_env.py:57-66: "<steps> <reward>" with
                             the reward printed as %.2f, to a file (LOG_FILE) or through `print`

The action-space wrappers (ClippedActions, UnboundedActions, ArgmaxActions) and the environment factory are gym-specific and
stay out of scope.
"""
import numpy as np

from ._trajectory import Trajectory
from .rollout import NormalizedObservations as _DeviceNormalizer


class _Wrapper(object):
    def __init__(self, env):
        self.env = env
        for name in ("observation_space", "action_space"):
            if hasattr(env, name):
                setattr(self, name, getattr(env, name))

    def step(self, action):
        return self._step(action)

    def reset(self):
        return self._reset()

    def _step(self, action):
        return self.env.step(action)

    def _reset(self):
        return self.env.reset()


class NormalizedObservations(_Wrapper):
    def __init__(self, env, shape=None):
        super().__init__(env)
        if shape is None:
            shape = env.observation_space.shape
        normalize = _DeviceNormalizer(shape)             # update + apply + clip(-5, 5) in one kernel

        def do_step(action):
            obs, reward, done, info = self.env.step(action)
            return normalize(np.asarray(obs, dtype=np.float32)), reward, done, info

        def do_reset():
            return normalize(np.asarray(self.env.reset(), dtype=np.float32))

        self._step, self._reset = do_step, do_reset
        self.get_mean, self.get_var, self.get_std = normalize.get_mean, normalize.get_var, normalize.get_std


class PrintRewards(_Wrapper):
    def __init__(self, env, every=1000, print=print):
        super().__init__(env)
        ep_rew, finished, total_steps, mean_rew = None, [], 0, 0.0

        def do_step(action):
            nonlocal ep_rew, finished, total_steps, mean_rew
            obs, reward, done, info = self.env.step(action)
            assert ep_rew is not None
            ep_rew += reward
            total_steps += 1
            if done:
                finished.append(ep_rew)
                ep_rew = None
            if total_steps % every == 0:
                if len(finished) >= 1:
                    mean_rew = round(float(np.mean(finished)), 2)
                    finished = []
                print(total_steps, mean_rew)
            return obs, reward, done, info

        def do_reset():
            nonlocal ep_rew
            ep_rew = 0.0
            return self.env.reset()

        self._step, self._reset = do_step, do_reset


def log_line(steps, reward):
    """The line benchmarks/_env.py:60-63 appends to LOG_FILE (the format of benchmarks/results/*)."""
    return "%d %.2f\n" % (steps, reward)


def one_step(env, policy):
    obs = env.next_obs if hasattr(env, "next_obs") else None
    obs = env.reset() if obs is None else obs
    act = policy(np.reshape(obs, -1))
    next_obs, rew, done, _ = env.step(act)
    env.next_obs = None if done else next_obs
    return obs, act, float(rew), bool(done)


def episode(env, policy, *, render=False, max_steps=10000):
    env.next_obs = env.reset()
    hist = []
    for _ in range(max_steps):
        *exp, done = one_step(env, policy)
        hist.append(exp)
        if done:
            break
    return Trajectory(*zip(*hist))
