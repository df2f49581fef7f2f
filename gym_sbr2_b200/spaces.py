"""`spaces.Box` as the reference uses it (gym_SBR_env2.py:64-66, gym_SBR_oneshot.py:106-113): taken from gym /
gymnasium when one is installed, otherwise a minimal stand-in with the same attributes (low, high, shape, dtype,
sample, contains).  The reference declares float32 boxes and never enforces them; neither does this package."""
import numpy as np


def _find_box():
    for name in ("gym.spaces", "gymnasium.spaces"):
        try:
            mod = __import__(name, fromlist=["Box"])
            return mod.Box
        except Exception:
            continue
    return None


class _Box(object):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.low = np.asarray(low, dtype=self.dtype)
        self.high = np.asarray(high, dtype=self.dtype)
        self.shape = tuple(self.low.shape if shape is None else shape)

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return "Box(%s, %s, %s, %s)" % (self.low, self.high, self.shape, self.dtype)


Box = _find_box() or _Box


def env_base():
    """gym.Env / gymnasium.Env when available (so isinstance checks of RL libraries pass), else object."""
    for name in ("gym", "gymnasium"):
        try:
            return __import__(name).Env
        except Exception:
            continue
    return object
