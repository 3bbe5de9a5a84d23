"""Host-side mirror of the reference API (models, rewards, rollout, mc_pilco, install) on CPU.
Tests marked `reference` import the UNMODIFIED reference from /root/reference (build container only)
and are skipped where it does not exist (the GPU box)."""
import os
import sys
from functools import partial

import numpy as np
import pytest
import torch

import golden_util as gu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import ref_shim  # noqa: E402

needs_reference = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted")


@pytest.fixture(autouse=True)
def eager_backend(monkeypatch):
    monkeypatch.setenv("PROB_MBRL_BACKEND", "eager")
    monkeypatch.setenv("PMB_NO_PBAR", "1")


def _build(models, reward_cls, D, U, hid, maxU, seed=3):
    torch.manual_seed(seed)
    np.random.seed(seed)
    od = models.DiagGaussianDensity(D)
    dm = models.mlp(D + U, 2 * D, hid, dropout_layers=[models.CDropout(0.1 * torch.ones(h)) for h in hid],
                    nonlin=torch.nn.ReLU)
    dyn = models.DynamicsModel(dm, reward_func=reward_cls(), output_density=od).float()
    pm_ = models.mlp(D, 2 * U, hid, dropout_layers=[models.BDropout(0.1) for _ in hid], nonlin=torch.nn.ReLU,
                     output_nonlin=partial(models.DiagGaussianDensity, U))
    pol = models.Policy(pm_, np.array([maxU]), np.array([-maxU])).float()
    g = torch.Generator().manual_seed(7)
    X = torch.randn(256, D + U, generator=g)
    X[:, -U:] *= maxU / 2
    Y = 1e-3 * torch.randn(256, D, generator=g)
    dyn.set_dataset(X, Y)
    dyn.eval()
    pol.train()
    return dyn, pol, g


@needs_reference
def test_mirror_modules_match_reference_modules_bit_for_bit():
    """Same seeds -> same parameters, same lazily drawn masks/noise, same rollout, same gradient."""
    ref_shim.install()
    import prob_mbrl as ref
    import prob_mbrl_b200 as pm
    out = []
    for models, rollout, reward in ((ref.models, ref.utils.rollout, ref.envs.cartpole.env.CartpoleReward),
                                    (pm.models, pm.rollout, pm.rewards.CartpoleReward)):
        dyn, pol, g = _build(models, lambda: reward(pole_length=torch.tensor(0.5)), 5, 1, [24, 24], 10.0)
        x0 = 0.1 * torch.randn(9, 5, generator=g)
        z_mm, z_rr = torch.randn(20 + 9, 5, generator=g), torch.randn(20 + 9, 1, generator=g)
        res = {}
        for mm in (False, True):
            pol.zero_grad()
            S, A, R = rollout(x0, dyn, pol, 12, resample_state_noise=False, resample_action_noise=False,
                              mm_states=mm, mm_rewards=mm, z_mm=z_mm, z_rr=z_rr)
            (-(torch.stack(R).sum(0) / 12).mean()).backward()
            res[mm] = (torch.stack(S).detach(), [p.grad.clone() for p in pol.parameters()])
        out.append((dyn.state_dict(), pol.state_dict(), res))
    (d1, p1, r1), (d2, p2, r2) = out
    for a, b in ((d1, d2), (p1, p2)):
        assert list(a.keys()) == list(b.keys())          # checkpoints interchange
        for k in a:
            assert torch.equal(a[k], b[k]), k
    for mm in (False, True):
        assert torch.equal(r1[mm][0], r2[mm][0])
        for ga, gb in zip(r1[mm][1], r2[mm][1]):
            assert torch.allclose(ga, gb, rtol=1e-5, atol=1e-9)


@needs_reference
@pytest.mark.parametrize("cls", ["CartpoleReward", "DoubleCartpoleReward", "CartAcrobotReward", "PendulumReward"])
def test_reward_mirrors_and_tip_form_match_reference(cls):
    ref_shim.install()
    import prob_mbrl as ref
    import prob_mbrl_b200 as pm
    from prob_mbrl_b200 import operands
    envmod = {"CartpoleReward": ref.envs.cartpole.env, "DoubleCartpoleReward": ref.envs.double_cartpole.env,
              "CartAcrobotReward": ref.envs.cart_acrobot.env, "PendulumReward": ref.envs.pendulum.env}[cls]
    kw = {"CartpoleReward": dict(pole_length=torch.tensor(0.5)), "PendulumReward": dict(pole_length=torch.tensor(1.0)),
          "DoubleCartpoleReward": dict(pole1_length=torch.tensor(0.6), pole2_length=torch.tensor(0.6)),
          "CartAcrobotReward": {}}[cls]
    r_ref = getattr(envmod, cls)(**kw)
    r_mir = getattr(pm.rewards, cls)(**kw)
    D = {"CartpoleReward": 5, "PendulumReward": 3}.get(cls, 8)
    g = torch.Generator().manual_seed(1)
    x, u = torch.randn(33, D, generator=g), torch.randn(33, 1, generator=g)
    want = r_ref(x, u)
    assert torch.allclose(r_mir(x, u), want, rtol=0, atol=2e-7)
    for mod in (r_ref, r_mir):      # both satisfy the duck-typed protocol read by the fused path
        ro = operands.read_reward(mod, D, 1)
        delta = x @ ro.C.t() + ro.c0
        got = (-0.5 * (((delta @ ro.Q) * delta).sum(-1, keepdim=True) + ((u @ ro.R) * u).sum(-1, keepdim=True))).exp()
        assert torch.allclose(got, want, rtol=0, atol=3e-7)


@needs_reference
def test_mc_pilco_mirror_equals_reference_mc_pilco():
    """Same seeds, pegasus: this package's mc_pilco (eager backend) == the reference's, parameter for parameter."""
    ref_shim.install()
    import prob_mbrl as ref
    import prob_mbrl_b200 as pm
    finals = []
    for models, algo, reward in ((ref.models, ref.algorithms.mc_pilco, ref.envs.cartpole.env.CartpoleReward),
                                 (pm.models, pm.mc_pilco, pm.rewards.CartpoleReward)):
        dyn, pol, g = _build(models, lambda: reward(pole_length=torch.tensor(0.5)), 5, 1, [16, 16], 10.0)
        x0 = 0.1 * torch.randn(8, 5, generator=g)
        opt = torch.optim.Adam(pol.parameters(), 1e-3)
        torch.manual_seed(5)
        losses = []
        algo(x0, dyn, pol, 6, opt, None, 4, pegasus=True, mm_states=True, mm_rewards=True, maximize=True,
             clip_grad=1.0, resampling_period=2, init_state_noise=0.01,
             on_iteration=lambda i, loss, s, a, r, d: losses.append((float(loss), len(s), len(a), len(r), tuple(r[0].shape))))
        finals.append(([p.detach().clone() for p in pol.parameters()], losses))
    (pa, la), (pb, lb) = finals
    assert la == lb and la[0][1:] == (7, 6, 6, (8, 1))
    for a, b in zip(pa, pb):
        assert torch.equal(a, b)


@needs_reference
def test_quadratic_saturating_cost_matches_reference_losses():
    """rewards.quadratic_loss / quadratic_saturating_loss / QuadraticSaturatingCost against the reference's
    losses.quadratic_loss / quadratic_saturating_loss (losses.py:67-75), and the (C, c0, Q, R, scale, offset) form the
    fused kernel evaluates against the module's own forward."""
    ref_shim.install()
    import prob_mbrl as ref
    from prob_mbrl_b200 import rewards
    from oracle import rollout_oracle as orc
    g = torch.Generator().manual_seed(4)
    x, u = torch.randn(11, 5, generator=g), torch.randn(11, 1, generator=g)
    A = torch.randn(5, 5, generator=g)
    Q, t = A @ A.T + torch.eye(5), torch.randn(1, 5, generator=g)
    assert torch.equal(rewards.quadratic_loss(x, t, Q), ref.losses.quadratic_loss(x, t, Q))
    assert torch.equal(rewards.quadratic_saturating_loss(x, t, Q), ref.losses.quadratic_saturating_loss(x, t, Q))
    cost = rewards.QuadraticSaturatingCost(t, Q)
    assert torch.allclose(cost(x, u), ref.losses.quadratic_saturating_loss(x, t, Q), rtol=0, atol=1e-7)
    for mod in (cost, rewards.QuadraticSaturatingCost(t, Q, R=torch.tensor([[0.05]]), reward=True)):
        C, c0, Qk, R, scale, offset = mod.tip_quadratic_form()
        flat = {"rew_C": C, "rew_c0": c0, "rew_Q": Qk, "rew_R": R, "rew_scale": scale, "rew_offset": offset, "D": 5, "U": 1}
        assert torch.allclose(orc.reward(flat, x, u), mod(x, u), rtol=0, atol=2e-7)


@needs_reference
def test_train_regressor_mirror_equals_reference():
    """utils.train_regressor: this package's module loop (eager backend) on the mirror modules == the reference's on
    its own, parameter for parameter, on identical numpy / torch seeds (reference utils/train_regressor.py:58-165)."""
    ref_shim.install()
    import prob_mbrl as ref
    import prob_mbrl_b200 as pm
    import numpy as np
    import tqdm
    from functools import partial
    finals = []
    for models, fit in ((ref.models, ref.utils.train_regressor), (pm.models, pm.train_regressor)):
        torch.manual_seed(11)
        np.random.seed(11)
        CD = models.modules.CDropout if hasattr(models, "modules") else models.CDropout
        net = models.mlp(6, 10, [24, 20], dropout_layers=[CD(0.1 * np.ones(h)) for h in (24, 20)], nonlin=torch.nn.ReLU)
        dyn = models.DynamicsModel(net, reward_func=None, output_density=models.DiagGaussianDensity(5)).float()
        g = torch.Generator().manual_seed(5)
        X = torch.randn(70, 6, generator=g)
        Y = 0.1 * torch.randn(70, 5, generator=g)
        dyn.set_dataset(X, Y)
        opt = torch.optim.Adam(dyn.parameters(), 1e-3)
        fit(dyn, 7, 32, True, opt, log_likelihood=dyn.output_density.log_prob, pbar_class=partial(tqdm.tqdm, disable=True))
        finals.append([p.detach().clone() for p in dyn.parameters()])
    for a, b in zip(*finals):
        assert (a - b).abs().max() < 2e-7           # same draws, same minibatches; the regulariser sums in another order


class _FakeExperience:
    """The three members mc_pilco's prioritized-replay branch touches (reference mc_pilco.py:223-231)."""

    def __init__(self, episodes):
        self.states = episodes

    def n_samples(self):
        return sum(len(e) for e in self.states)

    def n_episodes(self):
        return len(self.states)


@needs_reference
def test_prioritized_replay_mirror_equals_reference():
    """prioritized_replay=True: initial states drawn from the sum tree, importance weights on the returns, priorities
    from the per-step norms of dL/da_t -- this package's mc_pilco + SumTree (eager backend) against the reference's
    mc_pilco + utils.SumTree on identical torch / numpy seeds: same losses, same parameters, same tree."""
    ref_shim.install()
    import prob_mbrl as ref
    import prob_mbrl_b200 as pm
    from prob_mbrl_b200.replay import SumTree
    import numpy as np
    finals = []
    for which, (models, algo, reward) in enumerate(((ref.models, ref.algorithms.mc_pilco, ref.envs.cartpole.env.CartpoleReward),
                                                   (pm.models, pm.mc_pilco, pm.rewards.CartpoleReward))):
        dyn, pol, g = _build(models, lambda: reward(pole_length=torch.tensor(0.5)), 5, 1, [16, 16], 10.0)
        x0 = 0.1 * torch.randn(8, 5, generator=g)
        episodes = [(0.1 * torch.randn(12, 5, generator=g)).tolist() for _ in range(3)]
        opt = torch.optim.Adam(pol.parameters(), 1e-3)
        mod = sys.modules["prob_mbrl.algorithms.mc_pilco" if which == 0 else "prob_mbrl_b200.mc_pilco"]
        mod.x0_tree = (ref.utils.SumTree if which == 0 else SumTree)(64)
        mod.episode_counter = 0
        torch.manual_seed(5)
        np.random.seed(5)
        losses = []
        algo(x0, dyn, pol, 6, opt, _FakeExperience(episodes), 5, pegasus=True, maximize=True, clip_grad=1.0,
             resampling_period=3, init_state_noise=0.01, prioritized_replay=True, priority_alpha=0.6,
             on_iteration=lambda i, loss, *a: losses.append(float(loss)))
        finals.append(([p.detach().clone() for p in pol.parameters()], losses, mod.x0_tree))
    (pa, la, ta), (pb, lb, tb) = finals
    assert la == lb and len(la) == 5
    for a, b in zip(pa, pb):
        assert torch.equal(a, b)
    assert ta.size == tb.size == 36 and np.array_equal(ta.counts, tb.counts)
    assert np.allclose(ta.sum_tree, tb.sum_tree, rtol=0, atol=0) and ta.max_p == tb.max_p


@needs_reference
def test_install_rebinds_reference_call_sites():
    """install() patches utils.rollout / utils.core.rollout / algorithms.mc_pilco of the reference package,
    and the reference's own mc_pilco then runs through this package's rollout with identical results."""
    ref_shim.install()
    import prob_mbrl as ref
    import prob_mbrl_b200 as pm

    def run():
        dyn, pol, g = _build(ref.models, lambda: ref.envs.cartpole.env.CartpoleReward(pole_length=torch.tensor(0.5)),
                             5, 1, [16, 16], 10.0)
        x0 = 0.1 * torch.randn(8, 5, generator=g)
        opt = torch.optim.Adam(pol.parameters(), 1e-3)
        torch.manual_seed(5)
        sys.modules["prob_mbrl.algorithms.mc_pilco"].mc_pilco(x0, dyn, pol, 5, opt, None, 3, pegasus=True,
                                                             resampling_period=99)
        return [p.detach().clone() for p in pol.parameters()]

    base = run()
    calls = []
    saved = pm.install(ref)
    try:
        assert ref.utils.rollout is pm.rollout and ref.algorithms.mc_pilco is pm.mc_pilco
        assert sys.modules["prob_mbrl.utils.core"].rollout is pm.rollout
        real = pm.rollout
        ref.utils.rollout = lambda *a, **k: (calls.append(1), real(*a, **k))[1]
        patched = run()           # the reference's loop, this package's rollout underneath
    finally:
        pm.uninstall(saved, ref)
    assert len(calls) == 3
    assert ref.utils.rollout is saved["rollout"]
    for a, b in zip(base, patched):
        assert torch.equal(a, b)


def test_mirror_eager_rollout_reproduces_golden_and_extraction_round_trips():
    import prob_mbrl_b200 as pm
    for name in ("cartpole_37x2_n7_h12", "dcartpole_48x3_n24_h30"):
        ops, g = gu.load(name)
        dyn, pol = gu.modules_from_ops(ops)
        H = int(g["H"])
        x0 = g["x0"].clone().requires_grad_(True)
        S, A, R = pm.rollout(x0, dyn, pol, H, resample_state_noise=False, resample_action_noise=False)
        loss = -(torch.stack(R).sum(0) / H).mean()
        loss.backward()
        assert abs(float(loss) - float(g["nomm_loss"])) < 1e-7
        assert gu.rel_l2([p.grad for p in pol.parameters()], gu.policy_grad_list(g, "nomm", ops)) < 1e-6
        flat = pm.operands.extract(dyn, pol, int(g["N"])).to_flat()
        for k, v in ops.items():
            if torch.is_tensor(v):
                assert torch.allclose(torch.as_tensor(flat[k]).float().reshape(v.shape), v, atol=1e-7), k


def test_eligibility_errors_are_not_runtime_errors():
    """mc_pilco treats RuntimeError as a numerical failure of one iteration and skips it (reference
    algorithms/mc_pilco.py:122-131); configuration errors must not be swallowed that way."""
    import prob_mbrl_b200 as pm
    assert not issubclass(pm.NotEligible, RuntimeError)
    ops, g = gu.load("cartpole_37x2_n7_h12")
    dyn, pol = gu.modules_from_ops(ops)
    pol.model.drop0 = torch.nn.Dropout(0.1)          # not a persistent-mask dropout
    with pytest.raises(pm.NotEligible):
        pm.operands.extract(dyn, pol, 7)
    dyn2, pol2 = gu.modules_from_ops(ops)
    dyn2.train()                                     # CDropout re-relaxes every call in train mode
    with pytest.raises(pm.NotEligible):
        pm.operands.extract(dyn2, pol2, 7)
    dyn3, pol3 = gu.modules_from_ops(ops)
    with pytest.raises(pm.NotEligible):              # mask buffers smaller than the batch
        pm.operands.extract(dyn3, pol3, 64)
    pm.operands.materialize_noise(dyn3, pol3, torch.zeros(64, 5))
    assert pm.operands.extract(dyn3, pol3, 64).pol.mask[0].shape == (64, 37)
