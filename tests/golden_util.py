"""Loading helpers for tests/golden/*.npz (see tests/golden/make_golden.py)."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name, dtype=torch.float32):
    """-> (flat operand dict of torch tensors, dict of the other arrays as torch tensors)."""
    raw = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    ops, rest = {}, {}
    for k in raw.files:
        v = raw[k]
        tgt = ops if k.startswith("op_") else rest
        key = k[3:] if k.startswith("op_") else k
        if v.shape == ():
            tgt[key] = v.item()
        elif np.issubdtype(v.dtype, np.floating):
            tgt[key] = torch.from_numpy(v.copy()).to(dtype)
        else:
            tgt[key] = torch.from_numpy(v.copy())
    return ops, rest


def policy_grad_list(rest, tag, ops):
    n = 0
    for i in range(int(ops["pol_L"]) + 1):
        n += 1 + (("pol_b%d" % i) in ops)
    return [rest["%s_grad%d" % (tag, i)] for i in range(n)]


def rel_l2(a, b):
    a = torch.cat([x.reshape(-1).double() for x in a]) if isinstance(a, (list, tuple)) else a.reshape(-1).double()
    b = torch.cat([x.reshape(-1).double() for x in b]) if isinstance(b, (list, tuple)) else b.reshape(-1).double()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))
