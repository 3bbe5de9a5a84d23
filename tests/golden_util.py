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


def modules_from_ops(ops, device="cpu"):
    """Rebuild (dynamics, policy) mirror modules whose tensors equal a fixture's operand bundle."""
    from prob_mbrl_b200 import models, rewards
    D, U = int(ops["D"]), int(ops["U"])
    Lp, Ld = int(ops["pol_L"]), int(ops["dyn_L"])
    hid_p = [ops["pol_W%d" % i].shape[0] for i in range(Lp)]
    hid_d = [ops["dyn_W%d" % i].shape[0] for i in range(Ld)]
    from functools import partial
    reward = rewards.CartpoleReward() if D == 5 else rewards.DoubleCartpoleReward()
    dyn_net = models.mlp(D + U, 2 * D, hid_d, dropout_layers=[models.CDropout(0.1 * torch.ones(h)) for h in hid_d])
    dyn = models.DynamicsModel(dyn_net, reward_func=reward, output_density=models.DiagGaussianDensity(D)).float()
    pol_net = models.mlp(D, 2 * U, hid_p, dropout_layers=[models.BDropout(1.0 - float(ops["pol_p%d" % i])) for i in range(Lp)],
                         output_nonlin=partial(models.DiagGaussianDensity, U))
    maxU = (ops["act_scale"] + ops["act_bias"]).numpy()
    minU = (ops["act_bias"] - ops["act_scale"]).numpy()
    pol = models.Policy(pol_net, maxU, minU).float()
    with torch.no_grad():
        for tag, net, L in (("pol", pol.model, Lp), ("dyn", dyn.model, Ld)):
            for i in range(L + 1):
                fc = getattr(net, "fc%d" % i) if i < L else net.fc_out
                fc.weight.copy_(ops["%s_W%d" % (tag, i)])
                fc.bias.copy_(ops["%s_b%d" % (tag, i)])
        for i in range(Lp):
            drop = getattr(pol.model, "drop%d" % i)
            drop.noise.data = ops["pol_mask%d" % i].clone()
        for i in range(Ld):
            drop = getattr(dyn.model, "drop%d" % i)
            drop.noise.data = torch.rand_like(ops["dyn_mask%d" % i])
            drop.concrete_noise = ops["dyn_mask%d" % i].clone()
        pol.model.fc_nonlin.z.data = ops["pol_z"].clone()
        dyn.output_density.z.data = ops["dyn_z"].clone()
        for k in ("mx", "iSx", "my", "Sy"):
            getattr(dyn, k).data = ops[k].reshape(1, -1).clone()
        dyn.Sx.data = dyn.iSx.reciprocal()
        dyn.iSy.data = dyn.Sy.reciprocal()
        dyn.X.data = torch.zeros(1, D + U)
    dyn.eval()
    pol.train()
    return dyn.to(device), pol.to(device)


def synthetic_ops(D=3, U=2, hid=(24, 20), N=9, seed=5, dtype=torch.float32, pol_density=True, dyn_density=True):
    """Random operand bundle with the fixtures' key schema for shapes the reference's environments do not offer
    (e.g. U > 1): checked against the oracle instead of a golden file."""
    g = torch.Generator().manual_seed(seed)

    def r(*s):
        return torch.randn(*s, generator=g)

    ops = {"D": D, "U": U, "pol_L": len(hid), "dyn_L": len(hid)}
    for tag, nin, nout, dens in (("pol", D, 2 * U if pol_density else U, pol_density),
                                 ("dyn", D + U, 2 * D if dyn_density else D, dyn_density)):
        dims = [nin] + list(hid) + [nout]
        for i in range(len(dims) - 1):
            scale = (2.0 / dims[i]) ** 0.5 * (0.3 if (i == len(dims) - 2 and tag == "dyn") else 1.0)
            ops["%s_W%d" % (tag, i)] = r(dims[i + 1], dims[i]) * scale
            ops["%s_b%d" % (tag, i)] = 0.1 * r(dims[i + 1])
        for i, h in enumerate(hid):
            ops["%s_mask%d" % (tag, i)] = (torch.rand(N, h, generator=g) < 0.9).float()
            ops["%s_p%d" % (tag, i)] = 0.9 if tag == "pol" else 1.0
        ops[tag + "_has_density"] = int(dens)
        if dens:
            ops[tag + "_z"] = r(N, U if tag == "pol" else D)
        ops[tag + "_lmax"] = float(torch.tensor(5.0).log())
    ops["act_scale"] = torch.tensor([2.0, 0.5, 1.0, 3.0][:U])
    ops["act_bias"] = torch.tensor([0.0, 0.1, -0.2, 0.3][:U])
    ops["mx"] = 0.1 * r(D + U)
    ops["iSx"] = 1.0 / (0.5 + torch.rand(D + U, generator=g))
    ops["my"] = 0.01 * r(D)
    ops["Sy"] = 0.05 * (0.5 + torch.rand(D, generator=g))
    ops["rew_C"] = r(2, D)
    ops["rew_c0"] = 0.1 * r(2)
    Q = r(2, 2)
    ops["rew_Q"] = Q @ Q.T + torch.eye(2)
    R = 0.1 * r(U, U)
    ops["rew_R"] = R @ R.T + 1e-2 * torch.eye(U)
    ops["rew_scale"] = 1.0
    ops["rew_offset"] = 0.0
    x0 = 0.3 * r(N, D)
    return {k: (v.to(dtype) if torch.is_tensor(v) else v) for k, v in ops.items()}, x0.to(dtype)


def min_abs_preactivation(d, x0, H, per_particle=False):
    """Smallest |ReLU input| over all unmasked hidden units, particles and steps of the (fp64) oracle rollout.  A value
    near the fp32 rounding error of the layer sum means the ReLU gate of that unit depends on the summation order:
    two correct fp32 implementations may then differ by O(1/width) in the gradients of that particle."""
    from oracle import rollout_oracle as orc
    S, A, _, _ = orc.forward_with_saved(d, x0, H)
    best = torch.full((x0.shape[0],), float("inf"), dtype=torch.float64)
    for t in range(H):
        for tag, x in (("pol", S[t]), ("dyn", (torch.cat([S[t], A[t]], -1) - d["mx"]) * d["iSx"])):
            h = x
            for l in range(int(d[tag + "_L"])):
                pre = h @ d["%s_W%d" % (tag, l)].T + d["%s_b%d" % (tag, l)]
                m = d["%s_mask%d" % (tag, l)]
                best = torch.minimum(best, (pre.abs() + (m == 0) * 1e9).min(1).values.double())
                h = torch.relu(pre) * m / d["%s_p%d" % (tag, l)]
    return best if per_particle else float(best.min())
