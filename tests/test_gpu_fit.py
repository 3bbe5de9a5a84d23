"""Dynamics-model fit on the GPU (reference utils/train_regressor.py:58-165) through the
C ABI (pmb_fit_gradient + pmb_clip_adam_step):
  * against the trace the UNMODIFIED reference produced (tests/golden/train_regressor_cartpole_64x48.npz, made by
    tests/golden/make_golden_train.py): fed the minibatch indices and the dropout noise the reference drew, the
    per-iteration log-likelihoods and every trained tensor (weights, biases, dropout logits) agree to 2e-6;
  * the drop-in train_regressor() on the fused path against the module loop on the same device and seeds (same numpy
    and torch random streams), including ragged last minibatches."""
import os
from functools import partial

import numpy as np
import pytest
import torch
import tqdm

import golden_util as gu  # noqa: F401  (path set-up)
from test_oracle_train_regressor import _load

pytestmark = pytest.mark.gpu


def _model_from(hid, P, lmax, D=5, U=1):
    from prob_mbrl_b200 import models
    net = models.mlp(D + U, 2 * D, hid, dropout_layers=[models.CDropout(0.1 * torch.ones(h)) for h in hid])
    dyn = models.DynamicsModel(net, reward_func=None, output_density=models.DiagGaussianDensity(D)).float()
    with torch.no_grad():
        for i in range(len(hid)):
            getattr(net, "fc%d" % i).weight.copy_(P["W%d" % i])
            getattr(net, "fc%d" % i).bias.copy_(P["b%d" % i])
            getattr(net, "drop%d" % i).logit_p.copy_(P["logit_p%d" % i])
        net.fc_out.weight.copy_(P["W%d" % len(hid)])
        net.fc_out.bias.copy_(P["b%d" % len(hid)])
    return dyn


def test_fit_iterations_match_reference_golden_trace():
    import prob_mbrl_b200 as pm
    g, hid, P0, Pf, batches, noises = _load()
    dyn = _model_from(hid, P0, float(g["lmax"]))
    assert [float(getattr(dyn.model, "drop%d" % i).temp) for i in range(len(hid))] == [float(t) for t in g["temp"]]
    # the whitened dataset of the fixture, verbatim: identity scalers
    Xw, Yw = torch.from_numpy(g["Xw"]), torch.from_numpy(g["Yw"])
    dyn.set_dataset(Xw, Yw)
    for k in ("mx", "my"):
        getattr(dyn, k).data = torch.zeros_like(getattr(dyn, k))
    for k in ("Sx", "iSx", "Sy", "iSy"):
        getattr(dyn, k).data = torch.ones_like(getattr(dyn, k))
    dyn = dyn.cuda()
    dyn.train()
    opt = torch.optim.Adam(dyn.parameters(), float(g["lr"]))
    fit = pm.FusedFit(dyn, opt, len(batches[0]), 1.0)
    lls = []
    for idx, noise in zip(batches, noises):
        lls.append(float(fit.step(idx.numpy(), noise=[(u.cuda(), b.cuda()) for u, b in noise])))
    ref = g["lls"]
    assert np.abs(np.array(lls) - ref).max() < 2e-6 * np.abs(ref).max()
    names = {"W": "fc%d.weight", "b": "fc%d.bias", "logit_p": "drop%d.logit_p"}
    sd = dict(dyn.model.named_parameters())
    L = len(hid)
    for k, want in Pf.items():
        kind = k.rstrip("0123456789")
        i = int(k[len(kind):])
        name = ("fc_out." + ("weight" if kind == "W" else "bias")) if (i == L and kind != "logit_p") else names[kind] % i
        assert (sd[name].detach().cpu() - want).abs().max() < 2e-6, k
    assert float(opt.state[dyn.model.fc0.weight]["step"]) == len(batches)


@pytest.mark.parametrize("N,M,hid", [(250, 100, [200, 200]), (96, 32, [48, 40, 24])])
def test_train_regressor_fused_equals_module_loop(N, M, hid, monkeypatch):
    import prob_mbrl_b200 as pm
    from prob_mbrl_b200 import models
    out = []
    for backend in ("eager", "fused"):
        monkeypatch.setenv("PROB_MBRL_BACKEND", backend)
        torch.manual_seed(3)
        np.random.seed(3)
        net = models.mlp(6, 10, hid, dropout_layers=[models.CDropout(0.1 * torch.ones(h)) for h in hid])
        dyn = models.DynamicsModel(net, reward_func=None, output_density=models.DiagGaussianDensity(5)).float()
        g = torch.Generator().manual_seed(5)
        X = torch.randn(N, 6, generator=g)
        Y = torch.tanh(X @ (0.3 * torch.randn(6, 5, generator=g))) * 0.1 + 0.01 * torch.randn(N, 5, generator=g)
        dyn = dyn.cuda()
        dyn.set_dataset(X.cuda(), Y.cuda())
        opt = torch.optim.Adam(dyn.parameters(), 1e-3)
        torch.manual_seed(9)
        pm.train_regressor(dyn, 12, M, True, opt, log_likelihood=dyn.output_density.log_prob,
                           pbar_class=partial(tqdm.tqdm, disable=True))
        assert not dyn.training
        out.append((torch.cat([p.detach().flatten() for p in dyn.parameters()]).cpu(),
                    dyn.model.drop0.concrete_noise.detach().cpu(), dyn.model.drop0.p.detach().cpu(),
                    float(opt.state[dyn.model.fc0.weight]["step"])))
    (pa, ma, qa, sa), (pb, mb, qb, sb) = out
    assert sa == sb == 13                       # the reference runs iters + 1 steps (train_regressor.py:160-162)
    assert (pa - pb).abs().max() < 5e-6
    assert ma.shape == mb.shape and (ma - mb).abs().max() < 1e-6 and (qa - qb).abs().max() < 1e-6
    # the rollout afterwards runs on the fused sweeps with the fitted model (buffers left consistent)
    monkeypatch.setenv("PROB_MBRL_BACKEND", "fused")
