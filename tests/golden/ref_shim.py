"""The import shim for the unmodified reference lives in baseline/ref_shim.py (it is shared with the
reference arm of bench.py and the acceptance runner); re-exported here for the fixture generators."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "baseline"))
import importlib.util as _ilu

_spec = _ilu.spec_from_file_location(
    "pmb_baseline_ref_shim",
    os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "baseline", "ref_shim.py"))
_mod = _ilu.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
REFERENCE_ROOT = _mod.REFERENCE_ROOT
available = _mod.available
install = _mod.install
