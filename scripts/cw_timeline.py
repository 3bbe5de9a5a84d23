"""Per-warp clock64 arrival marks of one forward step of the wide cluster-resident sweeps (cluster 0 / rank 0)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PMB_CUDA_GRAPH"] = "0"; os.environ["PMB_NO_PBAR"] = "1"
import torch, bench
import prob_mbrl_b200 as pm
cfg = sys.argv[1] if len(sys.argv) > 1 else "c5"
n = bench.CONFIGS[cfg][5]
dyn, pol, x0, H, mm = bench.build_workload(cfg, n, "cuda")
opt = torch.optim.Adam(pol.parameters(), 1e-4)
g_r = torch.full((H, n), -1.0 / (H * n), device="cuda")
eng = pm.FusedIteration(dyn, pol, x0.cuda(), H, opt, g_r, 1.0)
eng.step(x0.cuda()); eng.step(x0.cuda())
dbg = torch.zeros(1024, dtype=torch.int64, device="cuda")
ptr = dbg.data_ptr()
eng.tune.reserved[2] = ptr & 0xffffffff if (ptr & 0xffffffff) < 2**31 else (ptr & 0xffffffff) - 2**32
eng.tune.reserved[3] = ptr >> 32
eng.step(x0.cuda()); torch.cuda.synchronize()
d = dbg.cpu().tolist()
names = {0: "step top", 1: "pol thin done", 2: "pol wide accum done", 3: "pol partials parked", 4: "pol epilogue+send done",
         5: "  owner: pol partials arrived", 6: "before wait(actions)", 7: "actions arrived", 8: "dyn thin done",
         9: "dyn wide accum done", 10: "dyn partials parked", 11: "dyn epilogue+send done", 12: "  owner: dyn partials arrived",
         13: "before wait(states)", 14: "states arrived"}
names_b = {32: "step top", 33: "xd arrived", 34: "dyn thin done", 35: "dyn accum done", 36: "dyn parked", 37: "dyn epilogue+send done",
           38: "  owner: dyn partials arrived", 39: "before wait(xp)", 40: "xp arrived", 41: "pol thin done", 42: "pol accum done",
           43: "pol parked", 44: "pol epilogue+send done", 45: "  owner: pol partials arrived", 46: "step end"}
for title, nm in (("forward", names), ("backward", names_b)):
    ks = sorted(nm)
    base = [x for x in d[16 * ks[0]: 16 * ks[0] + 16] if x]
    if not base:
        continue
    t0 = min(base)
    print(title, "step: per-warp arrival (cycles after the first warp entered the step)")
    for k in ks:
        v = [x - t0 if x else -1 for x in d[16 * k: 16 * k + 16]]
        vv = [x for x in v if x >= 0]
        if vv:
            print("  %-32s min %6d max %6d | %s" % (nm[k], min(vv), max(vv), " ".join("%6d" % x for x in v)))
