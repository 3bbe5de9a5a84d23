"""Acceptance on the GPU (north_star: "examples/deep_pilco_* run unchanged against it"): the REFERENCE's own
nn.Modules, moved to the GPU, go through operands.extract / the fused rollout and reproduce the golden fixture the
reference produced on the CPU; and the reference's example scripts run unchanged under install() with the
policy-gradient iterations on the device-resident engine.  Needs the reference package: /root/reference (build
container) or the git-ignored baseline/_ref copy that baseline/install_reference.sh makes and gpurun ships."""
import json
import os
import subprocess
import sys
from functools import partial

import numpy as np
import pytest
import torch

import golden_util as gu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import ref_shim  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_shim.available(), reason="reference package not installed (baseline/_ref)")]


def _reference_modules(envname, hid, N, seed=3):
    """SURVEY.md App. C.2 recipe with the reference's classes (the construction of tests/golden/make_golden.py)."""
    ref_shim.install()
    from prob_mbrl import envs, models, utils
    torch.set_num_threads(4)
    env = getattr(envs, envname)()
    D, U = env.observation_space.shape[0], env.action_space.shape[0]
    torch.manual_seed(seed)
    np.random.seed(seed)
    od = models.DiagGaussianDensity(D)
    dm = models.mlp(D + U, 2 * D, hid, dropout_layers=[models.modules.CDropout(0.1 * np.ones(h)) for h in hid],
                    nonlin=torch.nn.ReLU)
    dyn = models.DynamicsModel(dm, reward_func=env.reward_func, output_density=od).float()
    pm_ = models.mlp(D, 2 * U, hid, dropout_layers=[models.modules.BDropout(0.1) for h in hid], nonlin=torch.nn.ReLU,
                     output_nonlin=partial(models.DiagGaussianDensity, U))
    pol = models.Policy(pm_, env.action_space.high, env.action_space.low).float()
    g = torch.Generator().manual_seed(7)
    X = torch.randn(512, D + U, generator=g)
    X[:, -U:] *= float(env.action_space.high[0]) / 2
    Y = 1e-3 * torch.randn(512, D, generator=g)
    dyn.set_dataset(X, Y)
    dyn.eval()
    pol.train()
    x0 = 0.1 * torch.randn(N, D, generator=g)
    utils.rollout(x0, dyn, pol, 1, resample_state_noise=False, resample_action_noise=False)
    pol.zero_grad()
    return dyn, pol, x0


@pytest.mark.parametrize("name,envname,hid", [("cartpole_200x2_n25_h40", "Cartpole", [200, 200]),
                                              ("dcartpole_48x3_n24_h30", "DoubleCartpole", [48, 48, 48])])
def test_reference_modules_on_the_gpu_reproduce_the_golden_fixture(name, envname, hid, monkeypatch):
    import prob_mbrl_b200 as pm
    monkeypatch.setenv("PROB_MBRL_BACKEND", "fused")
    ops, g = gu.load(name)
    H, N = int(g["H"]), int(g["N"])
    dyn, pol, x0 = _reference_modules(envname, hid, N)
    assert torch.equal(x0, g["x0"])
    dyn, pol = dyn.cuda(), pol.cuda()
    x = x0.cuda().requires_grad_(True)
    S, A, R = pm.rollout(x, dyn, pol, H, resample_state_noise=False, resample_action_noise=False)
    loss = -(torch.stack(R).sum(0) / H).mean()
    loss.backward()
    assert (torch.stack(S).cpu() - g["nomm_states"]).abs().max() < 2e-6
    assert (torch.stack(A).cpu() - g["nomm_actions"]).abs().max() < 2e-5
    assert (torch.stack(R).squeeze(-1).cpu() - g["nomm_rewards"]).abs().max() < 2e-6
    assert abs(float(loss) - float(g["nomm_loss"])) <= 1e-6 * abs(float(g["nomm_loss"])) + 1e-8
    grads = [p.grad.cpu() for p in pol.parameters()]
    assert gu.rel_l2(grads, gu.policy_grad_list(g, "nomm", ops)) < 1e-5
    assert gu.rel_l2(x.grad.cpu(), g["nomm_dx0"]) < 1e-5


@pytest.mark.parametrize("script", ["deep_pilco_no_mm.py", "deep_pilco_mm.py"])
def test_reference_example_runs_unchanged_on_the_fused_backend(script, tmp_path):
    env = dict(os.environ, PMB_NO_PBAR="1", PROB_MBRL_BACKEND="fused")
    cmd = [sys.executable, os.path.join(ROOT, "baseline", "run_example.py"), script, "--use_cuda", "--ps_iters", "2",
           "--pol_opt_iters", "20", "--dyn_opt_iters", "50", "--n_initial_epi", "1", "-o", str(tmp_path)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("ACCEPTANCE ")][-1]
    st = json.loads(line[len("ACCEPTANCE "):])
    assert st["cuda"] and st["backend"] == "fused"
    assert st["mc_pilco_calls"] == 2 and st["engine_steps"] == 40        # every iteration on the device engine
    assert st["fit_steps"] == 2 * 51                                     # ... and every dynamics-fit iteration (iters + 1)
    assert st["plans"] and all("N=100 H=15" in p for p in st["plans"])
    assert all(np.isfinite(v) for v in st["losses"])
    assert "Traceback" not in out.stderr
