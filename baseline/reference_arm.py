"""The reference arm: the UNMODIFIED reference's own `algorithms.mc_pilco` (reference
algorithms/mc_pilco.py:13-267) timed on the host cores, through the reference's own modules
(`models.mlp / DynamicsModel / Policy`, `envs.*Reward`) built by the SURVEY.md App. C.2 recipe.

Baseline infrastructure only (bench.py --impl reference and bench.py's cpu_baseline leg); nothing in
prob_mbrl_b200/ imports it.  The reference is imported from /root/reference when that mount exists (build
container) or from baseline/_ref (the git-ignored `pip install --target` copy, baseline/install_reference.sh)
on the GPU box -- see baseline/ref_shim.py.
"""
import os
import sys
import time
from functools import partial

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

ENV_CLASS = {"cartpole": "Cartpole", "double_cartpole": "DoubleCartpole"}


def available():
    import ref_shim
    return ref_shim.available()


def build_reference_workload(env, hid, n_global, H, seed=3):
    """SURVEY.md App. C.2 bounded synthetic fixture out of the reference's own classes."""
    import ref_shim
    ref_shim.install()
    import torch
    from prob_mbrl import envs, models, utils
    e = getattr(envs, ENV_CLASS[env])()                     # BEFORE seeding (the env ctor draws from torch's RNG)
    D, U = e.observation_space.shape[0], e.action_space.shape[0]
    torch.manual_seed(seed)
    np.random.seed(seed)
    od = models.DiagGaussianDensity(D)
    dm = models.mlp(D + U, 2 * D, hid, dropout_layers=[models.modules.CDropout(0.1 * np.ones(h)) for h in hid],
                    nonlin=torch.nn.ReLU)
    dyn = models.DynamicsModel(dm, reward_func=e.reward_func, output_density=od).float()
    pm_ = models.mlp(D, 2 * U, hid, dropout_layers=[models.modules.BDropout(0.1) for _ in hid], nonlin=torch.nn.ReLU,
                     output_nonlin=partial(models.DiagGaussianDensity, U))
    pol = models.Policy(pm_, e.action_space.high, e.action_space.low).float()
    g = torch.Generator().manual_seed(7)
    X = torch.randn(512, D + U, generator=g)
    X[:, -U:] *= float(e.action_space.high[0]) / 2
    Y = 1e-3 * torch.randn(512, D, generator=g)
    dyn.set_dataset(X, Y)
    dyn.eval()
    pol.train()
    x0 = 0.1 * torch.randn(n_global, D, generator=g)
    utils.rollout(x0, dyn, pol, 1, resample_state_noise=False, resample_action_noise=False)   # [N, h] masks / z
    pol.zero_grad()
    return dyn, pol, x0


def time_reference(env, hid, n_global, H, mm, threads, iters, warmup=1, budget_s=60.0, Hs=None):
    """rollout-steps/s of `algorithms.mc_pilco(x0, dyn, pol, Hs, opt, None, K, pegasus=True, ...)`: K consecutive
    iterations (rollout + loss + backward incl. the dynamics weight gradients the reference forms and discards +
    clip + Adam), timed between on_iteration callbacks.  The horizon is cut to Hs (throughput of the reference is
    horizon-independent; the autograd graph of c5 at full H would need ~80 GB)."""
    import ref_shim
    ref_shim.install()
    import torch
    from prob_mbrl import algorithms
    torch.set_num_threads(int(threads))
    if Hs is None:
        Hs = min(H, 40 if mm else 100)      # mm: the reference's own z_mm table explodes beyond ~40 steps (App. D-7)
    dyn, pol, x0 = build_reference_workload(env, hid, n_global, H)
    opt = torch.optim.Adam(pol.parameters(), 1e-4)
    stamps = []

    def on_iteration(i, loss, *a):
        stamps.append(time.perf_counter())

    kw = dict(pegasus=True, mm_states=bool(mm), mm_rewards=bool(mm), maximize=True, clip_grad=1.0,
              resampling_period=499, init_state_noise=0.0, on_iteration=on_iteration)
    devnull = open(os.devnull, "w")
    old_err = sys.stderr
    sys.stderr = devnull                      # tqdm progress bar of the reference
    try:
        t0 = time.perf_counter()
        algorithms.mc_pilco(x0, dyn, pol, Hs, opt, None, warmup, **kw)
        one = (time.perf_counter() - t0) / max(1, warmup)
        k = max(2, min(int(iters), int(budget_s / max(one, 1e-3))))
        stamps.clear()
        t0 = time.perf_counter()
        algorithms.mc_pilco(x0, dyn, pol, Hs, opt, None, k, **kw)
    finally:
        sys.stderr = old_err
        devnull.close()
    done = len(stamps)                        # a failed (non-PD) iteration is skipped by the reference: not counted
    dt = (stamps[-1] - t0) if done else float("nan")
    return {"value": n_global * Hs * done / dt if done else 0.0, "unit": "rollout-steps/s", "cores": int(threads),
            "kind": "reference", "iterations": done, "horizon_timed": Hs, "ms_per_iteration": 1e3 * dt / max(done, 1),
            "sample": "%d iterations of the unmodified reference algorithms.mc_pilco (pegasus=True%s), N=%d, "
                      "H=%d (of %d), torch.set_num_threads(%d)"
                      % (done, ", mm_states=mm_rewards=True" if mm else "", n_global, Hs, H, threads)}
