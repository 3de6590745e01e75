"""Trajectory (mannequin/_trajectory.py:6-105): immutable (obs, act, rew) container.

Container semantics (dtype rules, slicing, prefix attribute access, errors) are
host logic and follow the reference line by line in behaviour.  The arrays are
mirrored lazily to the device (`dev()`), where minibatch sampling is a row
gather (csrc/mq_ops.cu) or is fused into the MLP kernels' operand loads.
"""
import numpy as np

from . import _device as D
from ._utils import discounted as _discounted


def _standardize(values):
    if isinstance(values, D.DeviceVector):
        values = np.asarray(values)
    if isinstance(values[0], (int, np.integer)):
        return np.asarray(values, dtype=np.int32)
    return np.asarray(values, dtype=np.float32)


class Trajectory(object):
    def __init__(self, observations, actions, rewards=None):
        observations = _standardize(observations)
        length = len(observations)
        if length < 1:
            raise ValueError("Cannot create empty trajectory")
        actions = _standardize(actions)
        if len(actions) != length:
            raise ValueError("Actions don't match observations")
        if rewards is None:
            rewards = np.ones(length, dtype=np.float32)
        else:
            rewards = np.broadcast_to(np.asarray(rewards, dtype=np.float32), (length,))
        for a in (observations, actions, rewards):
            a.setflags(write=False)
        object.__setattr__(self, "_fields", (observations, actions, rewards))
        object.__setattr__(self, "_devcache", None)

    # -- reference API -----------------------------------------------------------
    def get_length(self):
        return len(self._fields[0])

    def get_observations(self):
        return self._fields[0][:]

    def get_actions(self):
        return self._fields[1][:]

    def get_rewards(self):
        return self._fields[2][:]

    def modified(self, **kwargs):
        data = dict(zip(("observations", "actions", "rewards"), self._fields))
        for k in kwargs:
            data[k] = kwargs[k](data[k]) if callable(kwargs[k]) else kwargs[k]
        return Trajectory(**data)

    def discounted(self, *, horizon):
        o, a, r = self._fields
        return Trajectory(observations=o, actions=a, rewards=_discounted(r, horizon=horizon))

    def __getattribute__(self, name):
        if not name.startswith("_"):
            if name == "observations"[:len(name)]:
                return object.__getattribute__(self, "_fields")[0][:]
            if name == "actions"[:len(name)]:
                return object.__getattribute__(self, "_fields")[1][:]
            if name == "rewards"[:len(name)]:
                return object.__getattribute__(self, "_fields")[2][:]
        return object.__getattribute__(self, name)

    def __setattr__(self, name, value):
        if name[:1] in ("o", "a", "r"):
            raise ValueError("Field '%s' is read-only" % name)
        return object.__setattr__(self, name, value)

    def __getitem__(self, idx):
        o, a, r = (f[idx] for f in self._fields)
        if len(r.shape) >= 1:
            return Trajectory(o, a, r)
        return Trajectory([o], [a], [r])

    def __len__(self):
        return self.get_length()

    def __str__(self):
        o, a, _ = self._fields
        return "<trajectory: %d x %s -> %s>" % (len(o), o.shape[1:], a.shape[1:])

    def joined(*ts):
        if len(ts) <= 1:
            ts = ts[0]
        return Trajectory(*(np.concatenate([t._fields[i] for t in ts], axis=0) for i in range(3)))

    # -- device side ---------------------------------------------------------------
    def dev(self):
        """(obs, act, rew) as device tensors (float32 / int32), uploaded once."""
        if self._devcache is None:
            object.__setattr__(self, "_devcache", tuple(D.from_host(f) for f in self._fields))
        return self._devcache

    def take_dev(self, idx):
        """Device row gather of all three fields for an int index array (t[idx])."""
        import torch
        idx_h = np.asarray(idx)
        n = self.get_length()
        if idx_h.size and (idx_h.min() < -n or idx_h.max() >= n):
            raise IndexError("index out of bounds for trajectory of length %d" % n)
        idx_d = D.from_host(idx_h.reshape(-1), np.int32)
        outs = []
        for f, t in zip(self._fields, self.dev()):
            row = int(np.prod(f.shape[1:])) if f.ndim > 1 else 1
            dst = torch.empty(idx_d.numel() * row, dtype=t.dtype, device=t.device)
            D.call("mq_gather_rows", D.ptr(t), D.ptr(idx_d), idx_d.numel(), row * f.itemsize, n,
                   D.ptr(dst), D.stream())
            outs.append(dst)
        return tuple(outs)
