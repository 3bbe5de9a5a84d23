"""Acceptance runner: execute one of the reference's own example scripts UNCHANGED
(examples/deep_pilco_no_mm.py, examples/deep_pilco_mm.py) with prob_mbrl_b200.install() active, i.e. with
`prob_mbrl.utils.rollout` / `prob_mbrl.algorithms.mc_pilco` / `prob_mbrl.utils.train_regressor` re-bound to the fused
sm_100a paths while everything else (environments, apply_controller, the nn.Modules themselves) stays the reference's.

    python baseline/run_example.py deep_pilco_no_mm.py --use_cuda --ps_iters 2 --pol_opt_iters 20 ...

Prints one summary line `ACCEPTANCE {...}` (JSON) with how many policy-gradient iterations ran on the device-resident
engine, the sweep variant / kernels the planner chose, and the predicted returns.  Baseline / test infrastructure.
"""
import json
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def main():
    import ref_shim
    ref = ref_shim.install()
    import torch
    import prob_mbrl_b200 as pm
    from prob_mbrl_b200 import _lib
    script = sys.argv[1]
    for cand in (script, os.path.join(ref_shim.REFERENCE_ROOT, "examples", script)):
        if os.path.exists(cand):
            script = cand
            break
    else:
        raise SystemExit("example %s not found (run baseline/install_reference.sh)" % sys.argv[1])
    pm.install(ref)
    stats = {"script": os.path.basename(script), "engine_steps": 0, "mc_pilco_calls": 0, "fit_steps": 0, "plans": [],
             "losses": []}
    real_fit_step = pm.FusedFit.step

    def counting_fit_step(self, idx, noise=None):
        stats["fit_steps"] += 1
        return real_fit_step(self, idx, noise)

    pm.FusedFit.step = counting_fit_step
    real_step = pm.FusedIteration.step

    def counting_step(self, x0):
        out = real_step(self, x0)
        stats["engine_steps"] += 1
        if stats["engine_steps"] % 10 == 1:
            stats["losses"].append(round(float(out), 6))
        plan = _lib.describe_plan(self.prob, self.tune)
        tag = {0: "streaming (rollout_fwd/bwd_kernel)", 1: "cluster-resident (cluster_fwd/bwd_kernel)",
               2: "tensor-core cluster (tc_fwd/bwd_kernel)"}.get(plan["variant"], str(plan["variant"]))
        desc = "%s N=%d H=%d ctas=%d" % (tag, self.N, self.H, plan["ctas"])
        if desc not in stats["plans"]:
            stats["plans"].append(desc)
        return out

    pm.FusedIteration.step = counting_step
    real_mc = ref.algorithms.mc_pilco

    def counting_mc(*a, **k):
        stats["mc_pilco_calls"] += 1
        return real_mc(*a, **k)

    ref.algorithms.mc_pilco = counting_mc
    assert ref.utils.rollout is pm.rollout
    sys.argv = [script] + sys.argv[2:]
    os.environ.setdefault("PMB_PBAR_EVERY", "10")
    try:
        runpy.run_path(script, run_name="__main__")
    finally:
        stats["cuda"] = bool(torch.cuda.is_available())
        stats["backend"] = os.environ.get("PROB_MBRL_BACKEND", "fused")
        print("ACCEPTANCE " + json.dumps(stats))


if __name__ == "__main__":
    main()
