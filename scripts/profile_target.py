"""Small driver for ncu: a few un-graphed fused iterations of the bench workload."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PMB_CUDA_GRAPH"] = "0"
os.environ["PMB_NO_PBAR"] = "1"
import torch
import bench
import prob_mbrl_b200 as pm

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = bench.CONFIGS[cfg][5]
dyn, pol, x0, H, mm = bench.build_workload(cfg, n, "cuda")
opt = torch.optim.Adam(pol.parameters(), 1e-4)
g_r = torch.full((H, n), -1.0 / (H * n), device="cuda")
eng = pm.FusedIteration(dyn, pol, x0.cuda(), H, opt, g_r, 1.0, mm)
for _ in range(iters):
    eng.step(x0.cuda())
torch.cuda.synchronize()
print("loss", float(eng.loss))
