"""Debug aid: wide cluster-resident sweeps vs streaming sweeps vs fp64 oracle, per gradient tensor."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import golden_util as gu
from oracle import rollout_oracle as orc
import test_gpu_parity as T

hid = tuple(int(x) for x in sys.argv[1].split(",")) if len(sys.argv) > 1 else (512, 512)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 80
H = int(sys.argv[3]) if len(sys.argv) > 3 else 12
kw = dict(D=4, U=1, hid=hid, N=N)
ops, x0 = gu.synthetic_ops(**kw)
ops64, x064 = gu.synthetic_ops(dtype=torch.float64, **kw)
a = T._run(ops, x0, H, env=T.SWEEPS["ring"])
b = T._run(ops, x0, H, env=T.SWEEPS["cw"])
ref = orc.loss_and_grads(ops64, x064, H)
keys = orc.policy_param_keys(ops64)
print("S", (a["S"] - b["S"]).abs().max().item(), "R", (a["R"] - b["R"]).abs().max().item())
for k, ga, gb in zip(keys, a["grads"], b["grads"]):
    r = ref["grads"][k]
    print("%-10s ring-vs-oracle %.2e  cw-vs-oracle %.2e  ring-vs-cw %.2e  |g| %.3e" % (
        k, gu.rel_l2([ga], [r]), gu.rel_l2([gb], [r]), gu.rel_l2([ga], [gb]), r.norm().item()))
print("dx0 ring %.2e cw %.2e" % (gu.rel_l2([a["dx0"]], [ref["dx0"]]), gu.rel_l2([b["dx0"]], [ref["dx0"]])))
d = (b["dx0"].double() - ref["dx0"]).norm(dim=1) / ref["dx0"].norm(dim=1)
bad = [i for i in range(N) if d[i] > 1e-5]
print("particles with dx0 error > 1e-5:", bad, [float("%.2e" % d[i]) for i in bad][:12])
