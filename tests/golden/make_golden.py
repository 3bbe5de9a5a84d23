"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
The reference (mcgillmrl/prob_mbrl) has no tests or golden vectors of its own, so
these files are how parity is pinned: inputs (operands extracted from the reference's modules,
x0, z_mm, z_rr) + the outputs of the reference's own ``utils.rollout`` / ``loss.backward()`` /
``algorithms.mc_pilco`` on them.  Construction follows SURVEY.md App. C.2 verbatim so the
App. C.3 known-answer losses are reproduced (asserted below).
"""
import os
import sys
from functools import partial

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_shim  # noqa: E402

ref_shim.install()
import torch  # noqa: E402
from prob_mbrl import utils, models, envs, algorithms  # noqa: E402  (the reference)
from prob_mbrl_b200 import operands  # noqa: E402


def build(envname, hid, N, H, seed=3):
    """SURVEY.md App. C.2 bounded synthetic fixture."""
    torch.set_num_threads(4)
    env = getattr(envs, envname)()                       # BEFORE seeding (env ctor draws from torch RNG)
    D, U = env.observation_space.shape[0], env.action_space.shape[0]
    torch.manual_seed(seed)
    np.random.seed(seed)
    od = models.DiagGaussianDensity(D)
    dm = models.mlp(D + U, 2 * D, hid,
                    dropout_layers=[models.modules.CDropout(0.1 * np.ones(h)) for h in hid],
                    nonlin=torch.nn.ReLU)
    dyn = models.DynamicsModel(dm, reward_func=env.reward_func, output_density=od).float()
    pm = models.mlp(D, 2 * U, hid, dropout_layers=[models.modules.BDropout(0.1) for h in hid],
                    nonlin=torch.nn.ReLU, output_nonlin=partial(models.DiagGaussianDensity, U))
    pol = models.Policy(pm, env.action_space.high, env.action_space.low).float()
    g = torch.Generator().manual_seed(7)
    X = torch.randn(512, D + U, generator=g)
    X[:, -U:] *= float(env.action_space.high[0]) / 2
    Y = 1e-3 * torch.randn(512, D, generator=g)
    dyn.set_dataset(X, Y)
    dyn.eval()
    pol.train()
    x0 = 0.1 * torch.randn(N, D, generator=g)
    z_mm = torch.randn(H + N, D, generator=g)
    z_rr = torch.randn(H + N, 1, generator=g)
    # materialise the [N,h] masks / z buffers exactly like step 0 of the reference would
    utils.rollout(x0, dyn, pol, 1, resample_state_noise=False, resample_action_noise=False)
    pol.zero_grad()
    return env, dyn, pol, x0, z_mm, z_rr


def run_reference(dyn, pol, x0, H, mm, z_mm, z_rr, mm_groups=None):
    pol.zero_grad()
    x0 = x0.clone().requires_grad_(True)
    s, a, r = utils.rollout(x0, dyn, pol, H, resample_state_noise=False, resample_action_noise=False,
                            mm_states=mm, mm_rewards=mm, z_mm=z_mm, z_rr=z_rr, mm_groups=mm_groups)
    loss = -(torch.stack(r).sum(0) / H).mean()
    loss.backward()
    out = {
        "states": torch.stack(s).detach(), "actions": torch.stack(a).detach(),
        "rewards": torch.stack(r).detach().squeeze(-1), "loss": loss.detach(), "dx0": x0.grad.clone(),
    }
    for i, p in enumerate(pol.parameters()):
        out["grad%d" % i] = p.grad.clone()
    pol.zero_grad()
    return out


def save(name, ops, extra):
    flat = {("op_" + k): (v.cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in ops.to_flat().items()}
    for k, v in extra.items():
        flat[k] = v.cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **flat)
    print("wrote %s (%.1f KB)" % (path, os.path.getsize(path) / 1024))


def whiten_rows(z, N):
    """Zero-mean, identity-sample-covariance copy of the first N rows of z (with the raw
    table the reference's own moment-matched rollout explodes by construction, App. D-7)."""
    zz = z[:N].double()
    zz = zz - zz.mean(0, keepdim=True)
    L = torch.linalg.cholesky(zz.T @ zz / (N - 1))
    zz = torch.linalg.solve_triangular(L, zz.T, upper=False).T
    out = z.clone()
    out[:N] = zz.float()
    return out


def fixture_rollout(name, envname, hid, N, H, kat=None, with_mm=True, thin=None, mm_groups=None, whiten=False,
                    with_nomm=True):
    env, dyn, pol, x0, z_mm, z_rr = build(envname, hid, N, H)
    if whiten:
        z_mm = whiten_rows(z_mm, N)
    ops = operands.extract(dyn, pol, N)
    extra = {"x0": x0, "z_mm": z_mm, "z_rr": z_rr, "H": H, "N": N}
    res = run_reference(dyn, pol, x0, H, False, z_mm, z_rr) if with_nomm else None
    if kat is not None:
        assert abs(float(res["loss"]) - kat) < 5e-9, (float(res["loss"]), kat)
        print("  KAT (SURVEY App. C.3) reproduced: loss32 = %.11f" % float(res["loss"]))
    modes = [("nomm", res)] if with_nomm else []
    if with_mm:
        modes.append(("mm", run_reference(dyn, pol, x0, H, True, z_mm, z_rr)))
        if mm_groups:
            modes.append(("mmg", run_reference(dyn, pol, x0, H, True, z_mm, z_rr, mm_groups)))
            extra["mm_groups"] = mm_groups
    for tag, r in modes:
        for k, v in r.items():
            if thin and k in ("states", "actions", "rewards"):
                v = v[::thin]
            extra["%s_%s" % (tag, k)] = v
    if thin:
        extra["thin"] = thin
    save(name, ops, extra)
    return res


def fixture_mc_pilco(name, envname, hid, N, H, iters, lr, mm=False):
    """K iterations of the reference's own algorithms.mc_pilco (pegasus, exp=None)."""
    env, dyn, pol, x0, _, _ = build(envname, hid, N, H)
    init = [p.detach().clone() for p in pol.parameters()]
    opt = torch.optim.Adam(pol.parameters(), lr)
    seen = {}
    real_rollout = utils.rollout

    def spy(*a, **k):
        seen["z_mm"], seen["z_rr"] = k.get("z_mm"), k.get("z_rr")
        return real_rollout(*a, **k)

    losses = []
    utils.rollout = spy
    import prob_mbrl
    prob_mbrl.utils.rollout = spy
    try:
        torch.manual_seed(11)
        algorithms.mc_pilco(x0, dyn, pol, H, opt, None, iters, pegasus=True, mm_states=mm, mm_rewards=mm,
                            maximize=True, clip_grad=1.0, resampling_period=499, init_state_noise=0.0,
                            on_iteration=lambda i, loss, *a: losses.append(float(loss)))
    finally:
        utils.rollout = real_rollout
        prob_mbrl.utils.rollout = real_rollout
    final = [p.detach().clone() for p in pol.parameters()]
    # operands AFTER the call = the masks/noise every iteration used (resample() ran before it 0)
    dyn.eval()
    for p, p0 in zip(pol.parameters(), init):
        p.data.copy_(p0)
    ops = operands.extract(dyn, pol, N)
    extra = {"x0": x0, "H": H, "N": N, "iters": iters, "lr": lr, "losses": np.asarray(losses, np.float64),
             "z_mm": seen["z_mm"], "z_rr": seen["z_rr"], "mm": int(mm)}
    for i, p in enumerate(final):
        extra["final%d" % i] = p
    save(name, ops, extra)
    print("  losses:", losses)


def main_c3():
    # c3: configs[2] -- c2 with moment matching of states and rewards, z_mm[:N] whitened
    fixture_rollout("cartpole_200x2_n100_h400_mm", "Cartpole", [200, 200], 100, 400, with_mm=True, thin=25,
                    whiten=True, with_nomm=False)


if __name__ == "__main__":
    if "--c3" in sys.argv:
        main_c3()
        sys.exit(0)
    # c1: BASELINE.json configs[0] -- Cartpole 2x[200], 25 particles, H=40
    fixture_rollout("cartpole_200x2_n25_h40", "Cartpole", [200, 200], 25, 40, kat=-0.12317804247)
    # small double-pole (D=8, 3 hidden layers) with mm_groups
    fixture_rollout("dcartpole_48x3_n24_h30", "DoubleCartpole", [48, 48, 48], 24, 30, mm_groups=2)
    # ragged sizes: widths not multiples of 4/32, odd particle count, single hidden layer policy
    fixture_rollout("cartpole_37x2_n7_h12", "Cartpole", [37, 37], 7, 12)
    # c2: configs[1] -- N=100, H=400 (trajectories thinned to every 25th step to stay small)
    fixture_rollout("cartpole_200x2_n100_h400", "Cartpole", [200, 200], 100, 400, kat=-0.06295508146,
                    with_mm=False, thin=25)
    # mc_pilco iterations (reference's own loop + torch Adam)
    fixture_mc_pilco("mcpilco_cartpole_32x2_n16_h10", "Cartpole", [32, 32], 16, 10, iters=6, lr=1e-3)
    fixture_mc_pilco("mcpilco_mm_cartpole_32x2_n16_h10", "Cartpole", [32, 32], 16, 10, iters=4, lr=1e-3, mm=True)
    main_c3()
