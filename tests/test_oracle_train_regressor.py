"""Oracle of the dynamics-model fit vs the golden fixture produced by the UNMODIFIED
reference (tests/golden/make_golden_train.py): fed the minibatch indices and the dropout noise the reference drew,
the restatement reproduces its log-likelihood trace and the trained tensors.  CPU only; the CUDA path for this row
does not exist yet -- this is the parity gate it will be built against."""
import os

import numpy as np
import torch

from oracle import train_regressor_oracle as tro

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = os.path.join(HERE, "golden", "train_regressor_cartpole_64x48.npz")

NAME_MAP = {"model.fc%d.weight": "W%d", "model.fc%d.bias": "b%d", "model.drop%d.logit_p": "logit_p%d"}


def _load(dtype=torch.float32):
    g = np.load(FIX, allow_pickle=False)
    hid = [int(h) for h in g["hid"]]
    L = len(hid)

    def key(name):
        for pat, out in NAME_MAP.items():
            for i in range(L):
                if name == pat % i:
                    return out % i
        return {"model.fc_out.weight": "W%d" % L, "model.fc_out.bias": "b%d" % L}[name]

    names = [str(n) for n in g["param_names"]]
    P0 = {key(n): torch.from_numpy(g["init." + n]).to(dtype) for n in names}
    Pf = {key(n): torch.from_numpy(g["final." + n]).to(dtype) for n in names}
    n_it, M = int(g["n_it"]), len(g["idx0"])
    batches = [torch.from_numpy(g["idx%d" % i]) for i in range(n_it)]
    noises = []
    for i in range(n_it):
        per_layer = []
        for l, h in enumerate(hid):
            u = torch.from_numpy(g["u%d_%d" % (i, l)]).to(dtype)
            b = np.unpackbits(g["b%d_%d" % (i, l)])[:M * h].reshape(M, h)
            per_layer.append((u, torch.from_numpy(b.astype(np.float32)).to(dtype)))
        noises.append(per_layer)
    return g, hid, P0, Pf, batches, noises


def test_parameter_order_matches_the_reference_model():
    g, hid, *_ = _load()
    names = [str(n) for n in g["param_names"]]
    assert names == ["model.fc0.weight", "model.fc0.bias", "model.drop0.logit_p", "model.fc1.weight", "model.fc1.bias",
                     "model.drop1.logit_p", "model.fc_out.weight", "model.fc_out.bias"]
    assert tro.param_keys(len(hid)) == ["W0", "b0", "logit_p0", "W1", "b1", "logit_p1", "W2", "b2"]


def test_oracle_reproduces_the_reference_training_trace():
    g, hid, P0, Pf, batches, noises = _load()
    Xw, Yw = torch.from_numpy(g["Xw"]), torch.from_numpy(g["Yw"])
    P, lls = tro.train_iterations(P0, Xw, Yw, batches, noises, len(hid), [float(t) for t in g["temp"]],
                                  float(g["lmax"]), [float(x) for x in g["reg_scale"]],
                                  [float(x) for x in g["drop_reg"]], float(g["lr"]))
    ref = g["lls"]
    assert len(lls) == int(g["n_it"]) == 10          # the reference runs iters + 1 steps (train_regressor.py:160-162)
    assert np.abs(np.array(lls) - ref).max() < 2e-6 * np.abs(ref).max()
    for k in P:
        assert (P[k] - Pf[k]).abs().max() < 2e-6, k
    # the dropout probabilities are trained too (concrete relaxation + entropy regulariser)
    assert (Pf["logit_p0"] - P0["logit_p0"]).abs().max() > 1e-3


def test_oracle_fp64_stays_within_the_fp32_error_budget():
    """The same trace in fp64: the fp32 reference agrees with it to a few 1e-6 relative, which is the budget a CUDA
    implementation of this row will be held to."""
    g, hid, P0, Pf, batches, noises = _load(torch.float64)
    Xw, Yw = torch.from_numpy(g["Xw"]).double(), torch.from_numpy(g["Yw"]).double()
    P, lls = tro.train_iterations(P0, Xw, Yw, batches, noises, len(hid), [float(t) for t in g["temp"]],
                                  float(g["lmax"]), [float(x) for x in g["reg_scale"]],
                                  [float(x) for x in g["drop_reg"]], float(g["lr"]))
    assert np.abs(np.array(lls) - g["lls"]).max() < 1e-5 * np.abs(g["lls"]).max()
    for k in P:
        assert (P[k] - Pf[k]).abs().max() < 1e-5, k
