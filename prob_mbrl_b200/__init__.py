"""prob_mbrl_b200 -- B200-native (sm_100a) imagined-rollout hot path of mcgillmrl/prob_mbrl.

Public surface (mirrors the reference's for this path):
    rollout(...)      drop-in for prob_mbrl.utils.rollout          (reference utils/rollout.py:62-163)
    mc_pilco(...)     drop-in for prob_mbrl.algorithms.mc_pilco    (reference algorithms/mc_pilco.py:13-267)
    models, rewards   host-side mirrors of the module API the path consumes
    install()         rebind the two functions inside an importable reference package so its
                      examples/deep_pilco_*.py run unchanged on the fused backend
"""
from . import models, operands, rewards  # noqa: F401
from .mc_pilco import FusedIteration, mc_pilco  # noqa: F401
from .operands import NotEligible  # noqa: F401
from .rollout import fused_rollout_tensors, rollout  # noqa: F401
from .train_regressor import FusedFit, train_regressor  # noqa: F401

__version__ = "0.1.0"


def install(reference_package=None):
    """Rebind ``utils.rollout`` / ``algorithms.mc_pilco`` of the reference package to this package's.

    The reference resolves ``utils.rollout`` through the package attribute at call time
    (algorithms/mc_pilco.py:101, algorithms/MBDDPG.py:154) and ``utils/core.py:111`` binds it at import,
    so all three names are patched.  Returns the originals so callers can undo the patch."""
    import sys
    if reference_package is None:
        import prob_mbrl as reference_package
    ref = reference_package
    saved = {"rollout": ref.utils.rollout, "mc_pilco": ref.algorithms.mc_pilco,
             "train_regressor": ref.utils.train_regressor}
    ref.utils.rollout = rollout
    ref.utils.train_regressor = train_regressor      # examples call utils.train_regressor(...) (deep_pilco_no_mm.py:216)
    core = sys.modules.get(ref.__name__ + ".utils.core")
    if core is not None and hasattr(core, "rollout"):
        core.rollout = rollout
    ref.algorithms.mc_pilco = mc_pilco
    return saved


def uninstall(saved, reference_package=None):
    import sys
    if reference_package is None:
        import prob_mbrl as reference_package
    ref = reference_package
    ref.utils.rollout = saved["rollout"]
    core = sys.modules.get(ref.__name__ + ".utils.core")
    if core is not None and hasattr(core, "rollout"):
        core.rollout = saved["rollout"]
    ref.algorithms.mc_pilco = saved["mc_pilco"]
    if "train_regressor" in saved:
        ref.utils.train_regressor = saved["train_regressor"]
