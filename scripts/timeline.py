"""Phase timeline (clock64 marks) of one forward and one backward step of the bench workload."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PMB_CUDA_GRAPH"] = "0"; os.environ["PMB_NO_PBAR"] = "1"
import torch, bench
import prob_mbrl_b200 as pm
from prob_mbrl_b200 import _lib
cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = bench.CONFIGS[cfg][5]
dyn, pol, x0, H, mm = bench.build_workload(cfg, n, "cuda")
opt = torch.optim.Adam(pol.parameters(), 1e-4)
g_r = torch.full((H, n), -1.0 / (H * n), device="cuda")
eng = pm.FusedIteration(dyn, pol, x0.cuda(), H, opt, g_r, 1.0, mm)
eng.step(x0.cuda()); eng.step(x0.cuda())
dbg = torch.zeros(512, dtype=torch.int64, device="cuda")
ptr = dbg.data_ptr()
eng.tune.reserved[2] = ptr & 0xffffffff if (ptr & 0xffffffff) < 2**31 else (ptr & 0xffffffff) - 2**32
eng.tune.reserved[3] = ptr >> 32
eng.step(x0.cuda()); torch.cuda.synchronize()
d = dbg.cpu().tolist()
names_f = {0:"top",1:"pol L0",2:"pol L1",3:"pol out(narrow)",7:"sync",8:"squash stage",9:"dyn L0",10:"dyn L1",11:"dyn out(narrow)",15:"sync",16:"density"}
prev = d[0]
print("forward step (cycles):")
for k in sorted(names_f):
    if d[k]: print("  %-18s +%6d  (t=%6d)" % (names_f[k], d[k]-prev, d[k]-d[0])); prev = d[k]
names_b = {32:"top",33:"density+wait sav",34:"dyn net bwd",35:"sync",36:"scaler+squash+preA",37:"pol net bwd",38:"sync",39:"gs+preB",40:"sync/end"}
prev = d[32]
print("backward step (cycles):")
for k in sorted(names_b):
    if d[k]: print("  %-18s +%6d  (t=%6d)" % (names_b[k], d[k]-prev, d[k]-d[32])); prev = d[k]

print("wide-layer internals (fwd): marks = entry, [acquire/sync, accumulate]*, partial write, reduce sync, epilogue")
for nm, base in (("pol L1", 2), ("dyn L1", 10)):
    v = d[64 + 12 * base: 64 + 12 * base + 12]
    v = [x for x in v if x]
    print("  %-7s" % nm, [v[i + 1] - v[i] for i in range(len(v) - 1)])


if os.environ.get("PMB_STREAM_MODE", "0") in ("0", "3"):
    print("cluster-resident forward sweep: per-warp arrival (cycles after the first warp entered the step), cluster 0 / rank 0")
    names = {0: "step top", 1: "pol thin done", 2: "pol wide accum done", 11: "  pol reduce+epilogue done",
             9: "  pol butterfly+gather done", 3: "pol epilogue+narrow+send done", 20: "  pol exchange arrived",
             4: "squash role done", 5: "dyn thin done", 6: "dyn wide accum done", 15: "  dyn reduce+epilogue done",
             13: "  dyn butterfly+gather done", 7: "dyn epilogue+narrow+send done", 21: "  dyn exchange arrived",
             8: "density role done"}
    names.update({24: "mm: tile record written", 25: "mm: grid barrier passed", 26: "mm: records combined, mean",
                  27: "mm: covariance done", 29: "mm: Cholesky done (thread 0)", 30: "mm: tile barrier passed",
                  28: "mm: statistics stored, resampled"})
    order = [0, 1, 2, 11, 9, 3, 20, 4, 5, 6, 15, 13, 7, 21, 24, 25, 26, 27, 29, 30, 28, 8]
    t0 = min(x for x in d[0:8] if x)
    for k in order:
        v = [x - t0 if x else -1 for x in d[8 * k: 8 * k + 8]]
        print("    %-30s | %s" % (names[k], " ".join("%6d" % x for x in v)))
    print("cluster-resident backward sweep: per-warp arrival")
    names = {32: "step top", 33: "density adjoint done", 34: "dyn thin done", 35: "dyn wide accum done",
             36: "dyn epilogue+narrow+send done", 37: "scaler/squash role done", 38: "pol thin done",
             39: "pol wide accum done", 40: "pol epilogue+narrow+send done", 41: "gs role done"}
    t0 = min(x for x in d[256:264] if x)
    for k in sorted(names):
        v = [x - t0 for x in d[8 * k: 8 * k + 8]]
        print("    %-30s min %6d max %6d  | %s" % (names[k], min(v), max(v), " ".join("%6d" % x for x in v)))
