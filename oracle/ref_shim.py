"""Import shim for the UNMODIFIED reference tree (container only; test infrastructure).

Used by ``tests/golden/gen_golden.py``, by the container-only cross-checks in ``tests/`` and by the
reference arms of ``bench.py`` (CPU baseline, torch-eager-on-B200), which time the reference's OWN
modules when a reference tree is present (``/root/reference`` here, ``baseline/_ref`` on the GPU box).
Nothing in the product package imports this file.

The reference eagerly imports every agent / environment (``mirl/agents/__init__.py:2-19``,
``mirl/environments/__init__.py:1-6``), which drags in simulators that are not installed
here.  A ``sys.meta_path`` finder appended LAST fabricates stub modules for exactly those
third-party roots; installed modules always win.  Recipe: SURVEY.md Appendix B.
"""
import collections
import collections.abc
import importlib.abc
import importlib.machinery
import os
import sys
import types
from unittest.mock import MagicMock

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _resolve_root():
    """MIRL_REFERENCE_ROOT, else the read-only tree of the build container, else the git-ignored copy
    ``baseline/_ref`` that ``__graft_entry__.build()`` installs so that the reference's own modules can be
    timed on the GPU box ("or its copy under baseline/_ref/ on the GPU box")."""
    env = os.environ.get("MIRL_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if os.path.isdir(os.path.join(cand, "mirl")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _resolve_root()

_ROOTS = {"gym", "gtimer", "matplotlib", "faiss", "termcolor", "kornia", "imageio",
          "dm_control", "dm_env", "mujoco_py", "robosuite", "tree", "d4rl", "skvideo"}


class _Env:
    pass


class _Wrapper(_Env):
    def __init__(self, env=None):
        self.env = env


class Box:
    def __init__(self, low=None, high=None, shape=None, dtype=None):
        self.shape = shape


_REAL = {
    "gym": dict(Env=_Env, Wrapper=_Wrapper, ObservationWrapper=_Wrapper,
                ActionWrapper=_Wrapper, RewardWrapper=_Wrapper),
    "gym.spaces": dict(Box=Box),
    "gym.core": dict(Env=_Env, Wrapper=_Wrapper),
}


class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return MagicMock()


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in _ROOTS:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Stub(spec.name)
        m.__path__ = []
        for k, v in _REAL.get(spec.name, {}).items():
            setattr(m, k, v)
        return m

    def exec_module(self, m):
        pass


_installed = False


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mirl"))


def install():
    """Make ``import mirl`` resolve to the reference tree (idempotent)."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for n in ("Iterable", "Mapping", "Sequence", "Callable"):
        if not hasattr(collections, n):
            setattr(collections, n, getattr(collections.abc, n))
    sys.path.insert(0, REFERENCE_ROOT)
    sys.meta_path.append(_Finder())
    import mirl.torch_modules.utils as ptu
    ptu.set_gpu_mode(False)
    _installed = True


class FakeEnv:
    """All the hot path reads from an env (base_model.py:18-27, base_value.py:5-7)."""

    def __init__(self, env_name="cheetah", obs_dim=17, act_dim=6):
        self.env_name = env_name
        self.observation_space = Box(shape=(obs_dim,))
        self.action_space = Box(shape=(act_dim,))
