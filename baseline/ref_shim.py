"""Import shim that lets the UNMODIFIED reference (mcgillmrl/prob_mbrl) import under Python 3.12 /
torch 2.11.  The reference is looked up at $PROB_MBRL_REFERENCE, else /root/reference (read-only mount
of the build container), else baseline/_ref (the git-ignored `pip install --target` copy made by
baseline/install_reference.sh, which travels to the GPU box with the repo snapshot).

Baseline / fixture / acceptance infrastructure only: used by `tests/golden/make_golden*.py`, by the tests
that compare against the live reference, by `bench.py --impl reference` (the reference arm) and by
`baseline/run_example.py` (the acceptance runs of examples/deep_pilco_*.py).  Nothing in
`prob_mbrl_b200/` imports it.

What it papers over:
  * `collections.Iterable` was removed in Python 3.10 (reference: utils/core.py:8,
    models/core.py:7, utils/experience_dataset.py:241);
  * third-party modules imported at package-import time that the rollout path never
    touches: matplotlib, gym, Box2D, tensorboardX.
"""
import collections
import collections.abc
import os
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    env = os.environ.get("PROB_MBRL_REFERENCE")
    for cand in ([env] if env else []) + ["/root/reference", os.path.join(_HERE, "_ref")]:
        if cand and os.path.isdir(os.path.join(cand, "prob_mbrl")):
            return cand
    return env or "/root/reference"


REFERENCE_ROOT = _find_root()


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "prob_mbrl"))


def _lenient(factory):
    def getter(name):
        if name.startswith("__"):
            raise AttributeError(name)
        return factory()
    return getter


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


class _Anything:
    """Class whose instances swallow every call/attribute (plot/Box2D stand-ins)."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return _Anything()


class _Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.shape = self.low.shape
        self.dtype = dtype

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)


class _Env:
    spec = None
    metadata = {}
    reward_range = (-float("inf"), float("inf"))

    def seed(self, seed=None):
        return [seed]


class _EzPickle:
    def __init__(self, *a, **k):
        pass


def _np_random(seed=None):
    return np.random.RandomState(seed), seed


def install():
    """Install the stubs and put the reference on sys.path. Idempotent."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    import torch  # noqa: F401  (import before the stubs exist: torch introspects sys.modules)
    if not hasattr(collections, "Iterable"):
        collections.Iterable = collections.abc.Iterable
    if "matplotlib" not in sys.modules:
        plt = _stub("matplotlib.pyplot")
        plt.__getattr__ = _lenient(_Anything)
        mpl = _stub("matplotlib", pyplot=plt)
        mpl.__getattr__ = _lenient(_Anything)
    if "gym" not in sys.modules:
        spaces = _stub("gym.spaces", Box=_Box)
        seeding = _stub("gym.utils.seeding", np_random=_np_random)
        gutils = _stub("gym.utils", seeding=seeding, EzPickle=_EzPickle)
        _stub("gym", Env=_Env, spaces=spaces, utils=gutils)
    if "Box2D" not in sys.modules:
        names = ("edgeShape", "circleShape", "fixtureDef", "polygonShape",
                 "revoluteJointDef", "contactListener")
        b2 = _stub("Box2D.b2", **{n: _Anything for n in names})
        box2d = _stub("Box2D", b2=b2, **{"b2" + n[0].upper() + n[1:]: _Anything for n in names})
        box2d.__getattr__ = _lenient(lambda: _Anything)
    if "tensorboardX" not in sys.modules:
        class SummaryWriter(_Anything):
            pass
        _stub("tensorboardX", SummaryWriter=SummaryWriter)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings
    warnings.filterwarnings("ignore", category=UserWarning)
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    import prob_mbrl  # noqa: F401
    return prob_mbrl
