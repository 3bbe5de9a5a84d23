"""Small driver for ncu: a few fused dynamics-model fit iterations (2x[200], batch 100)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from functools import partial
import numpy as np, torch, tqdm
import prob_mbrl_b200 as pm
from prob_mbrl_b200 import models

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 4
torch.manual_seed(0)
np.random.seed(0)
hid = [200, 200]
net = models.mlp(6, 10, hid, dropout_layers=[models.CDropout(0.1 * torch.ones(h)) for h in hid])
dyn = models.DynamicsModel(net, reward_func=None, output_density=models.DiagGaussianDensity(5)).float().cuda()
X = torch.randn(2000, 6, device="cuda")
Y = torch.tanh(X[:, :5]) * 0.1
dyn.set_dataset(X, Y)
opt = torch.optim.Adam(dyn.parameters(), 1e-3)
os.environ["PROB_MBRL_BACKEND"] = "fused"
pm.train_regressor(dyn, iters, 100, True, opt, log_likelihood=dyn.output_density.log_prob,
                   pbar_class=partial(tqdm.tqdm, disable=True))
torch.cuda.synchronize()
print("fit ok", float(opt.state[dyn.model.fc0.weight]["step"]))
