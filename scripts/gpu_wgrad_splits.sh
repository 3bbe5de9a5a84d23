#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/wg.py <<'PY'
import os, sys, ctypes as C
sys.path.insert(0, os.getcwd())
os.environ["PMB_NO_PBAR"] = "1"
import torch, bench
import prob_mbrl_b200 as pm
from prob_mbrl_b200 import _lib
dyn, pol, x0, H = bench.build_workload("c2", 100, "cuda")
opt = torch.optim.Adam(pol.parameters(), 1e-4)
g_r = torch.full((H, 100), -1.0 / (H * 100), device="cuda")
eng = pm.FusedIteration(dyn, pol, x0.cuda(), H, opt, g_r, 1.0)
eng.step(x0.cuda()); torch.cuda.synchronize()
lib = eng.lib; st = _lib.current_stream_ptr(); pb = C.byref(eng.prob)
def tp(fn, reps=20):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
def bwd(ph):
    tune = _lib.make_tuning(phases=ph)
    return lambda: _lib.check(lib.pmb_rollout_backward(pb, C.byref(tune), eng.states.data_ptr(), eng.actions.data_ptr(), eng.rewards.data_ptr(), None, None, eng.g_rewards.data_ptr(), eng.grad_flat.data_ptr(), eng.dx0.data_ptr(), eng.ws.data_ptr(), eng.nbytes, st))
t0 = tp(lambda: eng.step(x0.cuda()))
print("splits", os.environ.get("PMB_WGRAD_SPLITS"), "umma", os.environ.get("PMB_WGRAD_UMMA"), "wgrad_ms %.4f iter_ms %.4f" % (tp(bwd(4)), t0))
PY
for s in 32 64 128 256; do PMB_WGRAD_SPLITS=$s timeout 100 python /tmp/wg.py 2>&1 | tail -1; done | tee gpurun_out/wgrad_splits.log
PMB_WGRAD_SPLITS=128 PMB_WGRAD_UMMA=2 timeout 100 python /tmp/wg.py 2>&1 | tail -1 | tee -a gpurun_out/wgrad_splits.log
