"""Golden fixture for the dynamics-model fit, produced by the UNMODIFIED reference.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_train.py
Runs the reference's `utils.train_regressor` on a small synthetic dataset and records what it consumed from
the random number generators (minibatch indices, the uniform noise and the hard Bernoulli samples of every
train-mode CDropout call) next to what it produced (mean log-likelihood per iteration, trained tensors).
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
import tqdm  # noqa: E402
from prob_mbrl import utils, models  # noqa: E402  (the reference)

D, U, HID, NDATA, BATCH, ITERS, LR = 5, 1, [64, 48], 256, 32, 9, 1e-3


def main():
    torch.set_num_threads(1)
    torch.manual_seed(11)
    np.random.seed(11)
    od = models.DiagGaussianDensity(D)
    net = models.mlp(D + U, 2 * D, HID, dropout_layers=[models.modules.CDropout(0.1 * np.ones(h)) for h in HID],
                     nonlin=torch.nn.ReLU)
    dyn = models.DynamicsModel(net, reward_func=None, output_density=od).float()
    g = torch.Generator().manual_seed(5)
    X = torch.randn(NDATA, D + U, generator=g)
    X[:, -U:] *= 5.0
    W = 0.3 * torch.randn(D + U, D, generator=g)
    Y = torch.tanh(X @ W) * 0.1 + 0.01 * torch.randn(NDATA, D, generator=g)
    dyn.set_dataset(X, Y)
    keys = [k for k, _ in dyn.named_parameters()]
    init = {k: v.detach().clone() for k, v in dyn.named_parameters()}
    Xw = ((dyn.X - dyn.mx) * dyn.iSx).detach().clone()
    Yw = ((dyn.Y - dyn.my) * dyn.iSy).detach().clone()

    # ---- recorders -------------------------------------------------------------------------------
    mod = sys.modules["prob_mbrl.utils.train_regressor"]
    batches, us, bs, lls = [], [], [], []
    orig_iter, orig_rand_like, orig_bern = mod.iterate_minibatches, torch.rand_like, torch.bernoulli

    def rec_iter(inputs, targets, batchsize):
        for x, y, idx in orig_iter(inputs, targets, batchsize):
            batches.append(np.asarray(idx).copy())
            yield x, y, idx

    def rec_rand_like(x, *a, **k):
        out = orig_rand_like(x, *a, **k)
        us.append(out.detach().clone())
        return out

    def rec_bern(p, *a, **k):
        out = orig_bern(p, *a, **k)
        bs.append(out.detach().clone())
        return out

    def rec_ll(y, mean, log_std=None):
        out = dyn.output_density.log_prob(y, mean, log_std)
        lls.append(float(out.mean()))
        return out

    opt = torch.optim.Adam(dyn.parameters(), LR)
    mod.iterate_minibatches, torch.rand_like, torch.bernoulli = rec_iter, rec_rand_like, rec_bern
    try:
        utils.train_regressor(dyn, ITERS, BATCH, True, opt, log_likelihood=rec_ll,
                              pbar_class=partial(tqdm.tqdm, disable=True))
    finally:
        mod.iterate_minibatches, torch.rand_like, torch.bernoulli = orig_iter, orig_rand_like, orig_bern
    n_it = len(lls)                       # the reference runs iters + 1 steps (train_regressor.py:160-162)
    L = len(HID)
    assert len(us) == L * n_it and len(bs) == L * n_it and len(batches) >= n_it, (len(us), len(bs), len(batches), n_it)
    out = {"D": D, "U": U, "hid": np.array(HID), "lr": LR, "n_it": n_it, "N": NDATA,
           "Xw": Xw.numpy(), "Yw": Yw.numpy(), "lls": np.array(lls, dtype=np.float64),
           "lmax": float(od.max_log_std),
           "temp": np.array([float(getattr(net, "drop%d" % i).temp) for i in range(L)]),
           "reg_scale": np.array([float(getattr(net, "drop%d" % i).regularizer_scale) for i in range(L)]),
           "drop_reg": np.array([float(getattr(net, "drop%d" % i).dropout_regularizer) for i in range(L)])}
    for k in keys:
        out["init." + k] = init[k].numpy()
        out["final." + k] = dict(dyn.named_parameters())[k].detach().numpy()
    for it in range(n_it):
        out["idx%d" % it] = batches[it].astype(np.int64)
        for l in range(L):
            out["u%d_%d" % (it, l)] = us[it * L + l].numpy()
            out["b%d_%d" % (it, l)] = np.packbits(bs[it * L + l].numpy().astype(np.uint8), axis=None)
    out["param_names"] = np.array(keys)
    path = os.path.join(HERE, "train_regressor_cartpole_64x48.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", n_it, "iterations; log-likelihoods", lls[:3], "...", lls[-1])


if __name__ == "__main__":
    main()
