"""Per-warp clock64() arrival marks of ONE step (t = H/2) of cluster 0 / rank 0 of the tensor-core forward sweep.
    python scripts/tc_timeline.py c5"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PMB_CUDA_GRAPH"] = "0"; os.environ["PMB_NO_PBAR"] = "1"; os.environ["PMB_STREAM_MODE"] = "4"
import torch, bench
import prob_mbrl_b200 as pm
cfg = sys.argv[1] if len(sys.argv) > 1 else "c5"
n = bench.CONFIGS[cfg][5]
dyn, pol, x0, H, mm = bench.build_workload(cfg, n, "cuda")
opt = torch.optim.Adam(pol.parameters(), 1e-4)
g_r = torch.full((H, n), -1.0 / (H * n), device="cuda")
eng = pm.FusedIteration(dyn, pol, x0.cuda(), H, opt, g_r, 1.0, mm)
eng.step(x0.cuda()); eng.step(x0.cuda())
dbg = torch.zeros(1024, dtype=torch.int64, device="cuda")
ptr = dbg.data_ptr()
eng.tune.reserved[2] = ptr & 0xffffffff if (ptr & 0xffffffff) < 2**31 else (ptr & 0xffffffff) - 2**32
eng.tune.reserved[3] = ptr >> 32
eng.step(x0.cuda()); torch.cuda.synchronize()
d = dbg.cpu().tolist()
names = {0: "pass top", 1: "first layer + epilogue + image stores issued", 3: "cluster barrier passed",
         4: "image -> TMEM + MMAs done (last wide layer)", 5: "tcgen05.ld + epilogue done", 6: "projection partials -> global",
         8: "cluster barrier passed", 9: "partials reduced", 11: "(mm start)", 10: "per-particle stage done"}
order = [0, 1, 2, 3, 4, 5, 6, 8, 9, 11, 10]
t0 = min(x for x in d[0:8] if x)
print("tensor-core forward sweep, %s: per-warp arrival (cycles after the step began), cluster 0 / rank 0" % cfg)
for which, nm in ((0, "policy"), (1, "dynamics")):
    for k in order:
        row = d[8 * (k + 16 * which): 8 * (k + 16 * which) + 8]
        if not any(row):
            continue
        v = [x - t0 if x else -1 for x in row]
        print("  %-9s %-46s | %s" % (nm, names[k], " ".join("%6d" % x for x in v)))

print("dynamics wide layer, per compute warp: cycles from entry to [loads issued, chunk 0 parked, chunk 1, ..., accumulator done]")
for w in range(12):
    v = [x for x in d[512 + 24 * w: 512 + 24 * w + 24] if x]
    if v:
        print("   warp %d: %s" % (w, " ".join("%6d" % (x - v[0]) for x in v)))
