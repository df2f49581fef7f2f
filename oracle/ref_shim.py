"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the unmodified reference package (`/root/reference/gym_SBR`) in the authoring container so that
golden vectors can be generated from the reference itself.  The reference imports
`gym` and `matplotlib` at module top (e.g. gym_SBR_env2.py:3-4, sub_phases_FB.py:4, buffer_tank3.py:6-9);
neither is installed here and there is no network, so both are replaced by inert stubs in `sys.modules`.
Nothing numerical is stubbed: numpy and scipy.integrate.odeint are the real ones.

`/root/reference` does not exist on the GPU box, so nothing that runs there may import this module.
"""
import contextlib
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SBR_REFERENCE_ROOT", "/root/reference")

_registry = {}


def _install_stubs():
    if "gym" in sys.modules and getattr(sys.modules["gym"], "_sbr_stub", False):
        return
    gym = types.ModuleType("gym")
    gym._sbr_stub = True

    class Env(object):
        pass

    gym.Env = Env
    spaces = types.ModuleType("gym.spaces")

    class Box(object):
        def __init__(self, low=None, high=None, shape=None, dtype=None):
            import numpy as np
            self.low = np.asarray(low)
            self.high = np.asarray(high)
            self.shape = self.low.shape if shape is None else shape
            self.dtype = dtype

    spaces.Box = Box
    gym.spaces = spaces
    envs = types.ModuleType("gym.envs")
    registration = types.ModuleType("gym.envs.registration")

    def register(id, entry_point=None, **kw):
        _registry[id] = entry_point

    registration.register = register
    envs.registration = registration
    gym.envs = envs
    sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.envs": envs,
                        "gym.envs.registration": registration})

    mpl = types.ModuleType("matplotlib")
    pyplot = types.ModuleType("matplotlib.pyplot")
    fm = types.ModuleType("matplotlib.font_manager")
    gs = types.ModuleType("matplotlib.gridspec")

    class FontProperties(object):
        def __init__(self, *a, **k):
            pass

        def set_size(self, *a, **k):
            pass

    fm.FontProperties = FontProperties
    mpl.pyplot, mpl.font_manager, mpl.gridspec = pyplot, fm, gs
    sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": pyplot,
                        "matplotlib.font_manager": fm, "matplotlib.gridspec": gs})


@contextlib.contextmanager
def quiet():
    """The reference prints on every reset/step; swallow it."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def load_reference():
    """Return the imported reference package `gym_SBR` (takes ~3 s: gym_SBR_env0 runs a cycle at import)."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "gym_SBR")):
        raise RuntimeError("reference checkout not found at %s (it only exists in the authoring container)"
                           % REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with quiet():
        import gym_SBR  # noqa: F401
        import gym_SBR.envs  # noqa: F401
    mod = sys.modules["gym_SBR"]
    if not os.path.realpath(mod.__file__).startswith(os.path.realpath(REFERENCE_ROOT)):
        raise RuntimeError("`gym_SBR` resolved to %s, not the reference" % mod.__file__)
    return mod


def registered_ids():
    return dict(_registry)
