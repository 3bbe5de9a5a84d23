"""CPU oracle of the dynamics-model fit (the next path to move onto the GPU).

TEST INFRASTRUCTURE ONLY -- nothing on the product path may import this file (see oracle/rollout_oracle.py).
There is no CUDA path for this row yet; the oracle and its golden fixture
(tests/golden/train_regressor_cartpole_64x48.npz, produced by the UNMODIFIED reference through
tests/golden/make_golden_train.py) are gate 3 of that work: the restatement below reproduces the reference's
`utils.train_regressor` iteration by iteration when it is fed the minibatch indices and the dropout noise the
reference drew.

Restates, on plain tensors:
  * minibatch objective of `train_regressor` (reference utils/train_regressor.py:58-165, default branch:
    no prioritized sampling, no decoupled regulariser):
        loss = -mean_b log N(y_b | mean_b, exp(log_std_b)^2) + reg_weight * R(theta) / N
    on the WHITENED dataset  X = (model.X - mx) * iSx,  Y = (model.Y - my) * iSy  (train_regressor.py:75-76);
  * `Regressor.forward(x, normalize=False, resample=True)` in train mode (models/core.py:169-187):
    Linear -> ReLU -> CDropout (x L) -> Linear -> DiagGaussianDensity without output scaling;
  * train-mode `CDropout` (models/modules.py:95-118,120-160): fresh uniform noise u per call,
        probs = sigmoid((logit_p + log((u + 1e-7) / (1 - (u - 1e-7)))) / temp),   b ~ Bernoulli(probs),
        mask  = (b - probs).detach() + probs          (hard sample forward, concrete relaxation backward);
  * `DiagGaussianDensity.forward` (models/densities.py:87-121): log_std = lmax - softplus(lmax - raw);
  * `DiagGaussianDensity.log_prob` (densities.py:123-144);
  * `BSequential.regularization_loss` + `CDropout.weights_regularizer` (modules.py:234-274, 87-93): for every
    dropout layer and the Linear layer that FOLLOWS it
        sum_j [ scale * p_j * sum_i W[i, j]^2 + dropout_regularizer * (p_j log p_j + (1 - p_j) log(1 - p_j)) ]
        + scale * sum_i bias_i^2                                   (BDropout.biases_regularizer, modules.py:32-33),
    p = sigmoid(logit_p)  (the keep probability; refreshed by the forward pass, modules.py:118),
    scale = 0.5 * regularizer_scale (modules.py:21-22);
  * `torch.optim.Adam.step()` on every trainable tensor (weights, biases, logit_p).
"""
import math

import torch

HALF_LOG_TWO_PI = 0.5 * math.log(2.0 * math.pi)


def param_keys(n_hidden):
    """Trainable tensors in `model.parameters()` order of a reference dynamics model built by models.mlp with
    CDropout layers: fc0.weight, fc0.bias, drop0.logit_p, fc1.weight, ..., fc_out.weight, fc_out.bias."""
    keys = []
    for i in range(n_hidden):
        keys += ["W%d" % i, "b%d" % i, "logit_p%d" % i]
    keys += ["W%d" % n_hidden, "b%d" % n_hidden]
    return keys


def concrete_mask(logit_p, temp, u, b):
    """Train-mode CDropout mask (models/modules.py:102-118) from the uniform noise `u` and the hard Bernoulli
    sample `b` the reference drew from it."""
    concrete_p = logit_p + ((u + 1e-7) / (1 - (u - 1e-7))).log()
    probs = (concrete_p / temp).sigmoid()
    return (b - probs).detach() + probs


def forward_train(P, x, noise, n_hidden, temp, lmax):
    """(mean, log_std) of the whitened targets for the whitened inputs x [M, D+U]; noise[i] = (u, b) of
    dropout layer i, each [M, h_i]."""
    h = x
    for i in range(n_hidden):
        h = torch.relu(torch.nn.functional.linear(h, P["W%d" % i], P["b%d" % i]))
        u, b = noise[i]
        h = h * concrete_mask(P["logit_p%d" % i], temp[i], u, b)
    o = torch.nn.functional.linear(h, P["W%d" % n_hidden], P["b%d" % n_hidden])
    D = o.shape[-1] // 2
    mean, raw = o.split(D, -1)
    log_std = -torch.nn.functional.softplus(-raw + lmax) + lmax
    return mean, log_std


def log_prob(y, mean, log_std):
    D = mean.shape[-1]
    deltas = mean - y
    return -0.5 * ((deltas * log_std.exp().reciprocal()) ** 2).sum(-1) - log_std.sum(-1) - D * HALF_LOG_TWO_PI


def regularization_loss(P, n_hidden, reg_scale, drop_reg):
    reg = 0
    for i in range(n_hidden):
        p = P["logit_p%d" % i].sigmoid()
        Wn = P["W%d" % (i + 1)]                       # the Linear layer AFTER dropout i ([out, in = h_i])
        r = reg_scale[i] * (p * (Wn ** 2).sum(0))
        r = r + drop_reg[i] * (p * p.log() + (1 - p) * (1 - p).log())
        reg = reg + r.sum() + reg_scale[i] * (P["b%d" % (i + 1)] ** 2).sum()
    return reg


def objective(P, x, y, noise, N, n_hidden, temp, lmax, reg_scale, drop_reg, reg_weight=1.0):
    mean, log_std = forward_train(P, x, noise, n_hidden, temp, lmax)
    enlml = -log_prob(y, mean, log_std).mean()
    return enlml + reg_weight * regularization_loss(P, n_hidden, reg_scale, drop_reg) / N, enlml


def train_iterations(P0, Xw, Yw, batches, noises, n_hidden, temp, lmax, reg_scale, drop_reg, lr, reg_weight=1.0):
    """Run len(batches) iterations of the reference loop on the whitened dataset (Xw, Yw).
       batches[i] : int64 indices of minibatch i (what iterate_minibatches yielded)
       noises[i]  : per dropout layer (u, b) drawn by the reference in iteration i
    Returns the trained tensors and the per-iteration mean log-likelihood (the progress-bar value)."""
    keys = param_keys(n_hidden)
    P = {k: P0[k].detach().clone().requires_grad_(True) for k in keys}
    opt = torch.optim.Adam([P[k] for k in keys], lr)
    N = Xw.shape[0]
    lls = []
    for idx, noise in zip(batches, noises):
        opt.zero_grad()
        loss, enlml = objective(P, Xw[idx], Yw[idx], noise, N, n_hidden, temp, lmax, reg_scale, drop_reg, reg_weight)
        loss.backward()
        opt.step()
        lls.append(float(-enlml.detach()))
    return {k: v.detach() for k, v in P.items()}, lls
