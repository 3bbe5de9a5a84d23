"""GPU tests of the device-resident iteration engine (FusedIteration) across mc_pilco calls: the engine is cached
between calls (the examples call mc_pilco once per episode, reference examples/deep_pilco_no_mm.py:199-264), so it
must notice everything the user does to the modules in between -- PEGASUS resample(), dynamics.set_dataset()
(re-binds mx/iSx/my/Sy, reference models/core.py:142-149), Policy.load()-style re-binding of the weights
(models/core.py:214-219), a changed learning rate -- and a failed moment-matching step must not touch the
parameters (reference algorithms/mc_pilco.py:122-131 skips the iteration)."""
import os
import sys

import pytest
import torch

import golden_util as gu

pytestmark = pytest.mark.gpu


def _episodes(backend, monkeypatch):
    """Two mc_pilco calls on the same modules with everything re-bound in between; returns losses + parameters."""
    import prob_mbrl_b200 as pm
    monkeypatch.setenv("PMB_NO_PBAR", "1")
    monkeypatch.setenv("PROB_MBRL_BACKEND", backend)
    ops, g = gu.load("cartpole_37x2_n7_h12")
    dyn, pol = gu.modules_from_ops(ops, "cuda")
    opt = torch.optim.Adam(pol.parameters(), 1e-3)
    sys.modules["prob_mbrl_b200.mc_pilco"].policy_update_counter[pol] = 0
    x0, H = g["x0"].cuda(), int(g["H"])
    torch.manual_seed(5)
    losses = []
    cb = lambda i, loss, *a: losses.append(float(loss))      # noqa: E731
    kw = dict(pegasus=True, maximize=True, clip_grad=1.0, init_state_noise=0.0, on_iteration=cb)
    # resampling_period=2: PEGASUS resample at iterations 0 and 2 of the first call -> engine.refresh()
    pm.mc_pilco(x0, dyn, pol, H, opt, None, 3, resampling_period=2, **kw)
    # new dataset => new input/output scalers (set_dataset re-binds the buffers)
    gen = torch.Generator().manual_seed(9)
    X = torch.randn(64, 6, generator=gen)
    X[:, -1] *= 4.0
    Y = 2e-3 * torch.randn(64, 5, generator=gen)
    dyn.set_dataset(X.cuda(), Y.cuda())
    # Policy.load()-style re-binding: same values, new storage
    for p in pol.parameters():
        p.data = p.data.clone()
    for grp in opt.param_groups:
        grp["lr"] = 3e-3
    pm.mc_pilco(x0, dyn, pol, H, opt, None, 3, resampling_period=10 ** 6, **kw)
    return losses, torch.cat([p.detach().flatten() for p in pol.parameters()]).cpu()


def test_engine_follows_resample_set_dataset_rebinding_and_lr_between_calls(monkeypatch):
    la, pa = _episodes("eager", monkeypatch)
    lb, pb = _episodes("fused", monkeypatch)
    assert len(la) == len(lb) == 6
    assert max(abs(a - b) for a, b in zip(la, lb)) < 2e-6, (la, lb)
    assert (pa - pb).abs().max() < 1e-5
    # the second episode really saw the new scalers: its losses differ from the first episode's
    assert abs(la[3] - la[2]) > 1e-6


def test_failed_moment_matching_step_leaves_parameters_and_adam_state_untouched(monkeypatch, capsys):
    """8 particles per group in 8 state dims: singular covariance at step 0 of every iteration.  The reference
    raises inside rollout and skips the iteration; the fused iteration predicates clip + Adam on the status word."""
    import prob_mbrl_b200 as pm
    monkeypatch.setenv("PMB_NO_PBAR", "1")
    monkeypatch.setenv("PROB_MBRL_BACKEND", "fused")
    ops, g = gu.load("dcartpole_48x3_n24_h30")
    dyn, pol = gu.modules_from_ops(ops, "cuda")
    opt = torch.optim.Adam(pol.parameters(), 1e-2)
    before = [p.detach().clone() for p in pol.parameters()]
    x0 = g["x0"][:16].cuda()
    pm.mc_pilco(x0, dyn, pol, 10, opt, None, 3, pegasus=True, mm_states=True, mm_rewards=True, mm_groups=2,
                init_state_noise=0.0, resampling_period=10 ** 6)
    out = capsys.readouterr().out
    assert out.count("RuntimeError") == 3
    for p, b in zip(pol.parameters(), before):
        assert torch.equal(p.detach(), b)
        st = opt.state[p]
        assert float(st["step"]) == 0.0 and float(st["exp_avg"].abs().max()) == 0.0


def test_library_errors_are_not_swallowed_as_numerical_failures(monkeypatch):
    """A failing library call (here: a workspace the engine did not size) must propagate, not be skipped like a
    non-PD covariance (LibraryMissing / CUDA errors were caught by `except RuntimeError`)."""
    import prob_mbrl_b200 as pm
    from prob_mbrl_b200 import _lib
    monkeypatch.setenv("PMB_NO_PBAR", "1")
    monkeypatch.setenv("PROB_MBRL_BACKEND", "fused")
    ops, g = gu.load("cartpole_37x2_n7_h12")
    dyn, pol = gu.modules_from_ops(ops, "cuda")
    opt = torch.optim.SGD(pol.parameters(), 1e-2)           # generic path (autograd.Function)

    def broken(*a, **k):
        raise _lib.LibraryError(-4, "simulated CUDA failure")
    monkeypatch.setattr(sys.modules["prob_mbrl_b200.mc_pilco"], "fused_rollout_tensors", broken)
    with pytest.raises(_lib.LibraryError):
        pm.mc_pilco(g["x0"].cuda(), dyn, pol, int(g["H"]), opt, None, 2, pegasus=True, init_state_noise=0.0)


def test_total_action_gradients_match_autograd_hooks(monkeypatch):
    """pmb_rollout_backward(..., da_total): the TOTAL dL/da_t of every step -- what a hook on actions[t] sees in the
    reference (algorithms/mc_pilco.py:160-188, prioritized replay) -- against hooks on the eager module loop (fp64)."""
    import prob_mbrl_b200 as pm
    from prob_mbrl_b200.rollout import fused_rollout_tensors
    for name, mode in (("cartpole_37x2_n7_h12", "3"), ("cartpole_37x2_n7_h12", "2"), ("dcartpole_48x3_n24_h30", "2"),
                       ("dcartpole_48x3_n24_h30", "4")):
        monkeypatch.setenv("PMB_STREAM_MODE", mode)
        ops, g = gu.load(name)
        H = int(g["H"])
        dyn, pol = gu.modules_from_ops(ops, "cuda")
        extras = {"want_action_grads": True}
        S, A, R, _ = fused_rollout_tensors(g["x0"].cuda(), dyn, pol, H, extras=extras)
        (-(R.sum(0) / H).mean()).backward()
        got = extras["action_grads"].cpu()
        monkeypatch.setenv("PROB_MBRL_BACKEND", "eager")
        dyn64, pol64 = gu.modules_from_ops(ops, "cpu")
        dyn64, pol64 = dyn64.double(), pol64.double()
        seen = {}
        Sr, Ar, Rr = pm.rollout(g["x0"].double(), dyn64, pol64, H, resample_state_noise=False, resample_action_noise=False)
        for t, a in enumerate(Ar):
            a.register_hook(lambda gr, t=t: seen.__setitem__(t, gr.detach().clone()))
        (-(torch.stack(Rr).sum(0) / H).mean()).backward()
        want = torch.stack([seen[t] for t in range(H)])
        monkeypatch.delenv("PROB_MBRL_BACKEND")
        assert gu.rel_l2(got, want) < 2e-5, (name, mode)


def test_prioritized_replay_runs_on_the_fused_rollout(monkeypatch):
    """mc_pilco(prioritized_replay=True) on the GPU: same losses / parameters as the eager module loop on the CPU with
    identical seeds (the sum tree lives on the host; priorities come from the fused reverse sweep's da_total)."""
    import numpy as np
    import prob_mbrl_b200 as pm
    from prob_mbrl_b200.replay import SumTree

    class Exp:
        def __init__(self, eps):
            self.states = eps

        def n_samples(self):
            return sum(len(e) for e in self.states)

        def n_episodes(self):
            return len(self.states)

    monkeypatch.setenv("PMB_NO_PBAR", "1")
    ops, g = gu.load("cartpole_37x2_n7_h12")
    gen = torch.Generator().manual_seed(3)
    episodes = [(0.1 * torch.randn(10, 5, generator=gen)).tolist() for _ in range(2)]
    out = []
    for dev, backend in (("cpu", "eager"), ("cuda", "fused")):
        monkeypatch.setenv("PROB_MBRL_BACKEND", backend)
        dyn, pol = gu.modules_from_ops(ops, dev)
        dyn.resample = lambda *a, **k: None          # keep the fixture's noise on both devices
        pol.resample = lambda *a, **k: None
        opt = torch.optim.Adam(pol.parameters(), 1e-3)
        mod = sys.modules["prob_mbrl_b200.mc_pilco"]
        mod.x0_tree, mod.episode_counter = SumTree(32), 0
        mod.policy_update_counter[pol] = 1
        torch.manual_seed(2)
        np.random.seed(2)
        losses = []
        pm.mc_pilco(g["x0"].to(dev), dyn, pol, int(g["H"]), opt, Exp(episodes), 4, pegasus=True, init_state_noise=0.0,
                    prioritized_replay=True, on_iteration=lambda i, loss, *a: losses.append(float(loss)))
        out.append((losses, torch.cat([p.detach().cpu().flatten() for p in pol.parameters()]), mod.x0_tree.sum_tree.copy()))
    (la, pa, ta), (lb, pb, tb) = out
    assert max(abs(a - b) for a, b in zip(la, lb)) < 2e-6
    assert (pa - pb).abs().max() < 5e-6
    assert np.allclose(ta, tb, rtol=2e-3, atol=1e-9)
