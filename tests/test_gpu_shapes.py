"""GPU tests at the other BASELINE.json shapes (no golden files: seeded inputs, CPU oracle as checker) and
full-size property tests (particle independence, linearity of the reverse sweep, determinism)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import golden_util as gu  # noqa: E402
from oracle import rollout_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu


def _fused(dyn, pol, x0, H, cot=None):
    from prob_mbrl_b200.rollout import fused_rollout_tensors
    for p in pol.parameters():
        p.grad = None
    x = x0.cuda().clone().requires_grad_(True)
    S, A, R, _ = fused_rollout_tensors(x, dyn, pol, H)
    if cot is None:
        obj = -(R.sum(0) / H).mean()
    else:
        obj = (R * cot).sum()
    obj.backward()
    return S.detach(), A.detach(), R.detach(), [p.grad.clone() for p in pol.parameters()], x.grad.clone(), obj.detach()


@pytest.mark.parametrize("cfg,N,H", [("c4", 13, 12), ("c5", 10, 12), ("c1", 25, 40)])
def test_other_baseline_shapes_match_oracle(cfg, N, H):
    """3x[400] double-pole (D=8) and 2x[512] nets: wide layers with 2 / 1 k-split groups, three hidden
    layers, larger rings -- against the fp32 and fp64 CPU oracle on the same seeded inputs."""
    from prob_mbrl_b200 import operands
    dyn, pol, x0, _, _ = bench.build_workload(cfg, N, "cpu")
    flat = operands.extract(dyn, pol, N).to_flat()
    ops32 = {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in flat.items()}
    ops64 = {k: (v.detach().double() if torch.is_tensor(v) else v) for k, v in flat.items()}
    r32 = orc.loss_and_grads(ops32, x0, H)
    r64 = orc.loss_and_grads(ops64, x0.double(), H)
    S, A, R, grads, dx0, obj = _fused(dyn.cuda(), pol.cuda(), x0, H)
    keys = orc.policy_param_keys(ops32)
    assert (S.cpu() - torch.stack(r32["states"])).abs().max() < 2e-6
    assert (A.cpu() - torch.stack(r32["actions"])).abs().max() < 4e-5          # actions are x maxU (10 / 20)
    assert (R.cpu() - torch.stack(r32["rewards"]).squeeze(-1)).abs().max() < 2e-6
    assert abs(float(obj) - float(r64["loss"])) < 1e-6 * abs(float(r64["loss"])) + 1e-8
    g64 = [r64["grads"][k] for k in keys]
    assert gu.rel_l2([g.cpu() for g in grads], g64) < max(1e-5, 3 * gu.rel_l2([r32["grads"][k] for k in keys], g64))
    assert gu.rel_l2(dx0.cpu(), r64["dx0"]) < 1e-5


def test_full_size_c5_shard_properties():
    """BASELINE.json configs[4] per-GPU shard (Cartpole 2x[512], 250 particles, H=1000): size-independent
    properties -- particles are independent (a sub-batch reproduces its rows), the reverse sweep is linear
    in the cotangent, results are finite and deterministic."""
    cfg, N, H = "c5", 250, 1000
    dyn, pol, x0, _, _ = bench.build_workload(cfg, N, "cuda")
    S, A, R, g1, dx1, _ = _fused(dyn, pol, x0, H)
    assert torch.isfinite(S).all() and torch.isfinite(R).all() and all(torch.isfinite(g).all() for g in g1)
    assert S.abs().max() < 50 and (R >= 0).all() and (R <= 1).all()
    # determinism
    S2, _, R2, g2, _, _ = _fused(dyn, pol, x0, H)
    assert torch.equal(S, S2) and all(torch.equal(a, b) for a, b in zip(g1, g2))
    # particle independence: the first 100 particles alone (another CTA shape: P=1 instead of P=2)
    Ss, As, Rs, _, dxs, _ = _fused(dyn, pol, x0[:100], H)
    assert (Ss - S[:, :100]).abs().max() < 1e-4 and (Rs - R[:, :100]).abs().max() < 1e-5
    # dL/dx0 of the mean-return loss scales with 1/N: rows of the sub-batch gradient = rows of the full one x N/100
    assert gu.rel_l2(dxs.cpu(), dx1[:100].cpu() * (N / 100.0)) < 1e-3
    # linearity of the reverse sweep in the reward cotangent
    gen = torch.Generator().manual_seed(1)
    c1 = torch.randn(H, N, generator=gen).cuda() / (H * N)
    c2 = torch.randn(H, N, generator=gen).cuda() / (H * N)
    _, _, _, ga, _, _ = _fused(dyn, pol, x0, H, cot=c1)
    _, _, _, gb, _, _ = _fused(dyn, pol, x0, H, cot=c2)
    _, _, _, gc, _, _ = _fused(dyn, pol, x0, H, cot=2.0 * c1 - 3.0 * c2)
    lin = [2.0 * a - 3.0 * b for a, b in zip(ga, gb)]
    assert gu.rel_l2([g.cpu() for g in gc], [g.cpu() for g in lin]) < 1e-4


def test_full_size_c4_shard_runs_and_is_finite():
    """configs[3] per-GPU shard (DoubleCartpole 3x[400], 125 particles, H=600)."""
    dyn, pol, x0, _, _ = bench.build_workload("c4", 125, "cuda")
    S, A, R, g, dx, _ = _fused(dyn, pol, x0, 600)
    assert torch.isfinite(S).all() and torch.isfinite(R).all() and all(torch.isfinite(t).all() for t in g)
    assert A.abs().max() <= 20.0 + 1e-4 and (R >= 0).all() and (R <= 1).all()
    assert sum(float(t.abs().sum()) for t in g) > 0


def test_per_step_noise_tables_match_oracle():
    """rollout() defaults (resample_state_noise / resample_action_noise = True, reference
    utils/rollout.py:68-69): fresh output noise every step.  The fused path pre-draws [H, N, .] tables in
    the reference's per-step order and the kernels read them with a step stride."""
    from prob_mbrl_b200 import operands
    from prob_mbrl_b200.rollout import FusedRolloutFunction
    ops, g = gu.load("cartpole_37x2_n7_h12")
    H, N = int(g["H"]), int(g["N"])
    gen = torch.Generator().manual_seed(3)
    zp = torch.randn(H, N, 1, generator=gen)
    zd = torch.randn(H, N, 5, generator=gen)
    o = operands.RolloutOperands.from_flat(ops, device="cuda")
    o.pol.z, o.dyn.z = zp.cuda(), zd.cuda()
    params = [p.requires_grad_(True) for p in o.policy_parameters()]
    x0 = g["x0"].cuda().requires_grad_(True)
    mm = dict(mm_states=False, mm_rewards=False, mm_groups=None, z_mm=None, z_rr=None)
    S, A, R, _ = FusedRolloutFunction.apply(x0, (o, N, H, mm), *params)
    loss = -(R.sum(0) / H).mean()
    grads = torch.autograd.grad(loss, params + [x0])
    d = dict(ops)
    d["pol_z"], d["dyn_z"] = zp, zd
    ref = orc.loss_and_grads(d, g["x0"], H)
    keys = orc.policy_param_keys(ops)
    assert (S.detach().cpu() - torch.stack(ref["states"])).abs().max() < 2e-6
    assert abs(float(loss) - float(ref["loss"])) < 1e-7
    assert gu.rel_l2([x.cpu() for x in grads[:-1]], [ref["grads"][k] for k in keys]) < 1e-5
    assert gu.rel_l2(grads[-1].cpu(), ref["dx0"]) < 1e-5


def test_rollout_default_flags_draw_fresh_noise_and_leave_buffers_like_the_reference():
    import prob_mbrl_b200 as pm
    ops, g = gu.load("cartpole_37x2_n7_h12")
    dyn, pol = gu.modules_from_ops(ops, "cuda")
    torch.manual_seed(0)
    S1, A1, R1 = pm.rollout(g["x0"].cuda(), dyn, pol, 6)           # defaults: resample_*_noise=True
    z_after = dyn.output_density.z.clone()
    S2, A2, R2 = pm.rollout(g["x0"].cuda(), dyn, pol, 6)
    assert len(S1) == 7 and not torch.equal(torch.stack(S1), torch.stack(S2))     # fresh noise each call
    assert not torch.equal(z_after, dyn.output_density.z) and dyn.output_density.z.shape == (7, 5)
    Sp1 = torch.stack(pm.rollout(g["x0"].cuda(), dyn, pol, 6, resample_state_noise=False, resample_action_noise=False)[0])
    Sp2 = torch.stack(pm.rollout(g["x0"].cuda(), dyn, pol, 6, resample_state_noise=False, resample_action_noise=False)[0])
    assert torch.equal(Sp1, Sp2)                                                  # PEGASUS: frozen noise


@pytest.mark.parametrize("variant", ["cvar", "sgd_discount", "value_func"])
def test_mc_pilco_generic_path_matches_eager_loop(variant, monkeypatch):
    """Loss variants that go through autograd on the fused rollout's outputs (reference
    algorithms/mc_pilco.py:136-188,193-194): CVaR subset, callable optimiser other than Adam + discount,
    value-function tail on states[-1].  Checked against the same mc_pilco on the eager module loop (CPU)."""
    import prob_mbrl_b200 as pm
    monkeypatch.setenv("PMB_NO_PBAR", "1")
    ops, g = gu.load("cartpole_37x2_n7_h12")
    results = []
    for dev, backend in (("cpu", "eager"), ("cuda", "fused")):
        monkeypatch.setenv("PROB_MBRL_BACKEND", backend)
        dyn, pol = gu.modules_from_ops(ops, dev)
        kw = dict(pegasus=True, maximize=True, clip_grad=1.0, resampling_period=10 ** 6, init_state_noise=0.0)
        vf = None
        if variant == "cvar":
            opt = torch.optim.Adam(pol.parameters(), 1e-3)
            kw.update(cvar_eps=0.5)
        elif variant == "sgd_discount":
            opt = torch.optim.SGD(pol.parameters(), 1e-2, momentum=0.9)
            kw.update(discount=0.97)
        else:
            opt = torch.optim.Adam(pol.parameters(), 1e-3)
            torch.manual_seed(1)
            lin = torch.nn.Linear(5, 1).to(dev)

            class V(torch.nn.Module):
                def forward(self, x, **kwargs):
                    return lin(x)

                def resample(self, **kwargs):
                    pass
            vf = V()
            kw.update(value_func=vf)
        sys.modules["prob_mbrl_b200.mc_pilco"].policy_update_counter[pol] = 1     # keep the fixture's noise
        # the initial resample() inside mc_pilco redraws masks: neutralise it so both runs use the fixture's
        dyn.resample = lambda *a, **k: None
        pol.resample = lambda *a, **k: None
        losses = []
        pm.mc_pilco(g["x0"].to(dev), dyn, pol, int(g["H"]), opt, None, 3,
                    on_iteration=lambda i, loss, *a: losses.append(float(loss)), **kw)
        results.append((torch.cat([p.detach().cpu().flatten() for p in pol.parameters()]), losses))
    (pa, la), (pb, lb) = results
    assert max(abs(a - b) for a, b in zip(la, lb)) < 2e-6
    assert (pa - pb).abs().max() < 5e-6
    assert (pa - torch.cat([ops[k].flatten() for k in orc.policy_param_keys(ops)])).abs().max() > 1e-4


@pytest.mark.parametrize("mode", ["2", "3", "4"])
@pytest.mark.parametrize("as_reward", [False, True])
def test_quadratic_saturating_cost_is_fused(mode, as_reward, monkeypatch):
    """north_star: "the quadratic/saturating cost in prob_mbrl.losses" -- a full D x D quadratic form on the state
    (reference losses.py:67-75) as reward_func, evaluated inside every sweep variant; vs the fp64 oracle."""
    from prob_mbrl_b200 import operands, rewards
    monkeypatch.setenv("PMB_STREAM_MODE", mode)
    ops, g = gu.load("cartpole_37x2_n7_h12")
    dyn, pol = gu.modules_from_ops(ops, "cuda")
    gen = torch.Generator().manual_seed(1)
    A = torch.randn(5, 5, generator=gen)
    dyn.reward_func = rewards.QuadraticSaturatingCost(torch.tensor([0.1, 0.0, 0.0, 0.2, -1.0]), A @ A.T + torch.eye(5),
                                                      R=torch.tensor([[1e-2]]), reward=as_reward).cuda()
    H = int(g["H"])
    S, A_, R, grads, dx0, obj = _fused(dyn, pol, g["x0"], H)
    flat = {k: (v.double().cpu() if torch.is_tensor(v) and v.is_floating_point() else v)
            for k, v in operands.extract(dyn, pol, 7).to_flat().items()}
    ref = orc.loss_and_grads(flat, g["x0"].double(), H)
    keys = orc.policy_param_keys(flat)
    assert (R.cpu().double() - torch.stack(ref["rewards"]).squeeze(-1)).abs().max() < 2e-6
    assert abs(float(obj) - float(ref["loss"])) < 1e-6
    assert gu.rel_l2([x.cpu() for x in grads], [ref["grads"][k] for k in keys]) < 2e-5
    assert gu.rel_l2(dx0.cpu(), ref["dx0"]) < 2e-5
