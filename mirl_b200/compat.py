"""Bridge to the reference's plug-in boundary.

MIRL selects classes through a JSON class factory that searches ``__subclasses__()`` of a base
class (``mirl/utils/initialize_utils.py:6-42``): ``component`` items must derive from
``mirl.components.Component`` (``mirl/components/__init__.py:1-3``), ``value`` items from
``mirl.values.Value`` (``mirl/values/base_value.py:4``).  When the reference is importable, our
classes derive from ITS bases, so ``"class": "B200PEModel"`` in a MIRL config resolves with no
edit to the reference.  Stand-alone (e.g. on the GPU box, where the
reference is absent) the local bases below are used.
"""
import abc


def _reference_importable():
    import importlib.util
    import sys
    if "mirl" in sys.modules:
        return True
    try:
        return importlib.util.find_spec("mirl") is not None
    except (ImportError, ValueError):
        return False


HAVE_REFERENCE = False
logger = None
if _reference_importable():
    try:
        from mirl.components import Component                  # noqa: F401
        from mirl.values.base_value import QValue, Value       # noqa: F401
        HAVE_REFERENCE = True
        try:
            from mirl.utils.logger import logger                # noqa: F401
        except Exception:                                       # pragma: no cover
            logger = None
    except ImportError:                                         # reference present, deps missing
        HAVE_REFERENCE = False

# ``pool`` items derive from mirl.pools.base_pool.Pool (mirl/pools/base_pool.py:5-32)
Pool = None
if HAVE_REFERENCE:
    try:
        from mirl.pools.base_pool import Pool                   # noqa: F401
    except Exception:                                           # pragma: no cover
        Pool = None
if Pool is None:

    class Pool(object):
        """mirl/pools/base_pool.py:5-32 (interface only)."""

        def __init__(self, env):
            pass

        def start_new_path(self, o):
            pass

        def end_a_path(self, next_o):
            pass


# ``policy`` items derive from mirl.policies.base_policy.Policy (mirl/policies/base_policy.py:10-34)
Policy = None
if HAVE_REFERENCE:
    try:
        from mirl.policies.base_policy import Policy            # noqa: F401
    except Exception:                                           # pragma: no cover
        Policy = None
if Policy is None:

    class Policy(object):
        """mirl/policies/base_policy.py:10-34 (interface only)."""

        def __init__(self, env):
            self.action_shape = env.action_space.shape
            self.observation_shape = env.observation_space.shape

        def get_diagnostics(self):
            return {}


if not HAVE_REFERENCE:

    class Component:
        """mirl/components/__init__.py:1-3."""

        def __init__(self):
            pass

    class Value(object, metaclass=abc.ABCMeta):
        """mirl/values/base_value.py:4-20."""

        def __init__(self, env):
            self.action_shape = env.action_space.shape
            self.observation_shape = env.observation_space.shape

        def save(self, save_dir=None):
            raise NotImplementedError

        def load(self, load_dir=None):
            raise NotImplementedError

        def get_diagnostics(self):
            return {}

        def get_snapshot(self):
            return {}

    class QValue(Value, metaclass=abc.ABCMeta):
        """mirl/values/base_value.py:35-46."""

        @abc.abstractmethod
        def value(self, obs, action):
            pass

        def value_np(self, obs, action, **kwargs):
            import torch
            dev = next(self.parameters()).device
            obs = torch.as_tensor(obs, device=dev).float()
            action = torch.as_tensor(action, device=dev).float()
            value, info = self.value(obs, action, **kwargs)
            return value.cpu().numpy(), {k: v.cpu().numpy() for k, v in info.items()}


def snapshot_dir():
    """``logger._snapshot_dir`` (pe_model.py:309-317) when the reference's logger is in use."""
    d = getattr(logger, "_snapshot_dir", None) if logger is not None else None
    if d is None:
        raise RuntimeError("no directory given and no reference logger snapshot dir is set")
    return d


# env-name aliases, mirl/environments/our_envs/__init__.py:58-73
SHORT_NAME = {
    "half_cheetah": "cheetah", "HalfCheetah-v2": "cheetah", "Swimmer-v2": "swimmer", "Ant-v2": "ant",
    "AntTruncatedObs-v2": "mb_ant", "Hopper-v2": "hopper", "Walker2d-v2": "walker2d",
    "Humanoid-v2": "humanoid", "HumanoidTruncatedObs-v2": "mb_humanoid",
    "InvertedPendulum-v2": "inverted_pendulum", "MultiGoal2D-v0": "2dpointer",
}


def short_env_name(name):
    return SHORT_NAME.get(name, name)
