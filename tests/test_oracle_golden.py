"""The CPU oracle (oracle/rollout_oracle.py) held to the golden fixtures produced by the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import pytest
import torch

import golden_util as gu
from oracle import rollout_oracle as orc

torch.set_num_threads(4)

ROLLOUT_FIXTURES = ["cartpole_200x2_n25_h40", "dcartpole_48x3_n24_h30", "cartpole_37x2_n7_h12"]


@pytest.mark.parametrize("name", ROLLOUT_FIXTURES)
def test_rollout_nomm_matches_reference(name):
    ops, g = gu.load(name)
    res = orc.loss_and_grads(ops, g["x0"], int(g["H"]))
    # same op sequence as the reference for policy+dynamics; the reward folds the constant tip
    # map into (C, c0) => <= a few ulp (SURVEY App. C.3 second table)
    assert (torch.stack(res["states"]) - g["nomm_states"]).abs().max() < 5e-7
    assert (torch.stack(res["actions"]) - g["nomm_actions"]).abs().max() < 5e-6
    assert (torch.stack(res["rewards"]).squeeze(-1) - g["nomm_rewards"]).abs().max() < 3e-7
    assert abs(float(res["loss"]) - float(g["nomm_loss"])) < 1e-7
    keys = orc.policy_param_keys(ops)
    assert gu.rel_l2([res["grads"][k] for k in keys], gu.policy_grad_list(g, "nomm", ops)) < 2e-6
    assert gu.rel_l2(res["dx0"], g["nomm_dx0"]) < 2e-6


@pytest.mark.parametrize("name", ROLLOUT_FIXTURES)
def test_rollout_mm_matches_reference(name):
    ops, g = gu.load(name)
    res = orc.loss_and_grads(ops, g["x0"], int(g["H"]), mm_states=True, mm_rewards=True,
                             z_mm=g["z_mm"], z_rr=g["z_rr"])
    # moment matching amplifies rounding (3.4e-4 on states for a re-ordered fp32)
    assert (torch.stack(res["states"]) - g["mm_states"]).abs().max() < 2e-3
    assert abs(float(res["loss"]) - float(g["mm_loss"])) < 2e-6
    keys = orc.policy_param_keys(ops)
    assert gu.rel_l2([res["grads"][k] for k in keys], gu.policy_grad_list(g, "mm", ops)) < 2e-3


def test_rollout_mm_groups_matches_reference():
    ops, g = gu.load("dcartpole_48x3_n24_h30")
    res = orc.loss_and_grads(ops, g["x0"], int(g["H"]), mm_states=True, mm_rewards=True,
                             z_mm=g["z_mm"], z_rr=g["z_rr"], mm_groups=int(g["mm_groups"]))
    assert abs(float(res["loss"]) - float(g["mmg_loss"])) < 2e-6
    keys = orc.policy_param_keys(ops)
    assert gu.rel_l2([res["grads"][k] for k in keys], gu.policy_grad_list(g, "mmg", ops)) < 2e-3


def test_known_answers_survey_c3():
    """fp32 reference losses on the bounded fixture."""
    ops, g = gu.load("cartpole_200x2_n25_h40")
    assert abs(float(g["nomm_loss"]) - (-0.12317804247)) < 5e-9
    assert abs(float(g["mm_loss"]) - (-0.17601011693)) < 5e-8
    ops, g = gu.load("cartpole_200x2_n100_h400")
    assert abs(float(g["nomm_loss"]) - (-0.06295508146)) < 5e-9


def test_c2_full_horizon_matches_reference():
    """BASELINE.json configs[1] (N=100, H=400): oracle vs thinned golden trajectory + gradient."""
    ops, g = gu.load("cartpole_200x2_n100_h400")
    res = orc.loss_and_grads(ops, g["x0"], int(g["H"]))
    thin = int(g["thin"])
    assert (torch.stack(res["states"])[::thin] - g["nomm_states"]).abs().max() < 2e-5
    assert abs(float(res["loss"]) - float(g["nomm_loss"])) < 1e-7
    keys = orc.policy_param_keys(ops)
    assert gu.rel_l2([res["grads"][k] for k in keys], gu.policy_grad_list(g, "nomm", ops)) < 1e-5


def test_c3_full_horizon_moment_matching_matches_reference():
    """BASELINE.json configs[2] (c2 with mm_states + mm_rewards, z_mm[:N] whitened per SURVEY.md section 8d):
    oracle vs the reference's thinned trajectory + gradient.  (fp32-vs-fp64 on this fixture: states 1.8e-6,
    gradient 3e-6.)"""
    ops, g = gu.load("cartpole_200x2_n100_h400_mm")
    res = orc.loss_and_grads(ops, g["x0"], int(g["H"]), mm_states=True, mm_rewards=True, z_mm=g["z_mm"],
                             z_rr=g["z_rr"])
    thin = int(g["thin"])
    assert (torch.stack(res["states"])[::thin] - g["mm_states"]).abs().max() < 2e-5
    assert abs(float(res["loss"]) - float(g["mm_loss"])) < 1e-7
    keys = orc.policy_param_keys(ops)
    assert gu.rel_l2([res["grads"][k] for k in keys], gu.policy_grad_list(g, "mm", ops)) < 2e-5
    assert float(g["mm_states"].abs().max()) < 5.0          # bounded (the unwhitened table reaches 1e10)


def test_fp64_twin_error_budget():
    """fp32 oracle vs fp64 oracle reproduces the reference's own fp32-vs-fp64 error scale
    (grad rel-L2 ~8.5e-8)."""
    ops, g = gu.load("cartpole_200x2_n25_h40")
    ops64, g64 = gu.load("cartpole_200x2_n25_h40", torch.float64)
    r32 = orc.loss_and_grads(ops, g["x0"], int(g["H"]))
    r64 = orc.loss_and_grads(ops64, g64["x0"], int(g["H"]))
    keys = orc.policy_param_keys(ops)
    err = gu.rel_l2([r32["grads"][k] for k in keys], [r64["grads"][k] for k in keys])
    assert err < 1e-6
    assert abs(float(r64["loss"]) - (-0.12317804490)) < 1e-9     # App. C.3 fp64 loss


@pytest.mark.parametrize("name", ["mcpilco_cartpole_32x2_n16_h10", "mcpilco_mm_cartpole_32x2_n16_h10"])
def test_mc_pilco_iterations_match_reference(name):
    """Oracle loop (rollout+loss+autograd+clip+Adam) vs the reference's own algorithms.mc_pilco."""
    ops, g = gu.load(name)
    mm = {}
    if int(g["mm"]):
        mm = dict(mm_states=True, mm_rewards=True, z_mm=g["z_mm"], z_rr=g["z_rr"])
    final, losses = orc.mc_pilco_iterations(ops, g["x0"], int(g["H"]), int(g["iters"]), float(g["lr"]), **mm)
    assert torch.allclose(torch.tensor(losses, dtype=torch.float64), g["losses"].double(), rtol=0, atol=2e-7)
    keys = orc.policy_param_keys(ops)
    tol = 5e-6 if int(g["mm"]) else 1e-6
    for i, k in enumerate(keys):
        assert (final[k] - g["final%d" % i]).abs().max() < tol, k
        # and the parameters really moved (Adam lr=1e-3 for 4-6 steps)
        assert (final[k] - ops[k]).abs().max() > 1e-4
