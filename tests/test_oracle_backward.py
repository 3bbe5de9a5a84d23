"""The hand-derived reverse sweep in the oracle (the executable spec csrc/rollout_bwd.cu
transliterates) against autograd, fp64 so the comparison is exact to ~1e-12."""
import pytest
import torch

import golden_util as gu
from oracle import rollout_oracle as orc


@pytest.mark.parametrize("name", ["cartpole_37x2_n7_h12", "dcartpole_48x3_n24_h30"])
@pytest.mark.parametrize("cots", ["loss", "generic"])
def test_manual_backward_equals_autograd(name, cots):
    ops, g = gu.load(name, torch.float64)
    H, N, D, U = int(g["H"]), int(g["N"]), int(ops["D"]), int(ops["U"])
    keys = orc.policy_param_keys(ops)
    d = dict(ops)
    for k in keys:
        d[k] = d[k].clone().requires_grad_(True)
    x0 = g["x0"].clone().requires_grad_(True)
    states, actions, rewards = orc.rollout(d, x0, H)
    gen = torch.Generator().manual_seed(0)
    if cots == "loss":
        gS = gA = None
        gR = torch.full((H, N), -1.0 / (H * N), dtype=torch.float64)
        obj = (torch.stack(rewards).squeeze(-1) * gR).sum()
    else:
        gS = torch.randn(H + 1, N, D, generator=gen, dtype=torch.float64)
        gA = torch.randn(H, N, U, generator=gen, dtype=torch.float64)
        gR = torch.randn(H, N, generator=gen, dtype=torch.float64)
        obj = ((torch.stack(states) * gS).sum() + (torch.stack(actions) * gA).sum()
               + (torch.stack(rewards).squeeze(-1) * gR).sum())
    auto = torch.autograd.grad(obj, [d[k] for k in keys] + [x0])
    with torch.no_grad():
        s, a, r, saved = orc.forward_with_saved(ops, g["x0"], H)
        grads, dx0 = orc.manual_backward(ops, s, a, r, saved, gS, gA, gR)
    for k, ga in zip(keys, auto[:-1]):
        assert torch.allclose(grads[k], ga, rtol=1e-9, atol=1e-12), k
    assert torch.allclose(dx0, auto[-1], rtol=1e-9, atol=1e-12)


def test_mm_adjoint_equals_autograd():
    """The hand-derived Cholesky/moment-matching adjoint (spec of the CUDA mm reverse step)."""
    g = torch.Generator().manual_seed(5)
    for M, D in ((25, 5), (12, 8), (9, 1)):
        x = torch.randn(M, D, generator=g, dtype=torch.float64).requires_grad_(True)
        z = torch.randn(M, D, generator=g, dtype=torch.float64)
        go = torch.randn(M, D, generator=g, dtype=torch.float64)
        y = orc.mm_resample(x, z)
        (auto,) = torch.autograd.grad((y * go).sum(), x)
        with torch.no_grad():
            y2, m, L, zh = orc.mm_forward_parts(x, z)
            man = orc.mm_backward(go, x, m, L, zh)
        assert torch.allclose(y2, y, atol=1e-12)
        assert torch.allclose(man, auto, rtol=1e-8, atol=1e-10)
