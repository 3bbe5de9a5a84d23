#!/bin/bash
mkdir -p gpurun_out
echo "== parity (cluster subset)"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "particles_per_cluster or sweep_variants or golden or cotangents" 2>&1 | tail -8 | tee gpurun_out/cluster_small.log
echo "== timeline"
timeout 300 python scripts/timeline.py c2 2>&1 | tail -24 | tee gpurun_out/timeline.log
echo "== bench quick"
for m in 3; do
  PMB_STREAM_MODE=$m timeout 300 python bench.py --steps 10 --warmup 3 --quick 2>&1 | tail -1
done | tee gpurun_out/ab.log
timeout 300 python - <<'PY' 2>&1 | tail -3 | tee -a gpurun_out/ab.log
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
def tp(fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
def fwd(ph):
    tune = _lib.make_tuning(phases=ph)
    return lambda: _lib.check(lib.pmb_rollout_forward(pb, C.byref(tune), eng.x0.data_ptr(), eng.states.data_ptr(), eng.actions.data_ptr(), eng.rewards.data_ptr(), eng.ws.data_ptr(), eng.nbytes, eng.status.data_ptr(), st))
def bwd(ph):
    tune = _lib.make_tuning(phases=ph)
    return lambda: _lib.check(lib.pmb_rollout_backward(pb, C.byref(tune), eng.states.data_ptr(), eng.actions.data_ptr(), eng.rewards.data_ptr(), None, None, eng.g_rewards.data_ptr(), eng.grad_flat.data_ptr(), eng.dx0.data_ptr(), eng.ws.data_ptr(), eng.nbytes, st))
print("fwd_sweep_ms %.3f bwd_sweep_ms %.3f wgrad_ms %.3f" % (tp(fwd(2)), tp(bwd(2)), tp(bwd(4))))
PY
