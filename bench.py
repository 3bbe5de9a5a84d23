#!/usr/bin/env python
"""Benchmark of the imagined-rollout hot path (BASELINE.json metric: rollout-steps/sec).

  python bench.py --gpus N --steps K --warmup W            fused sm_100a path (this repo)
  python bench.py --impl reference --gpus N --steps K ...  CPU baseline: the oracle port of the reference
                                                           algorithm on the box's host cores

A "step" is ONE mc_pilco policy-gradient iteration (weight packing, forward sweep over H imagined steps,
reverse sweep, batched policy weight gradient, gradient clip, Adam) on a batch of particles;
rollout-steps/sec = particles x horizon x K / time.  Workload = BASELINE.json configs[1]: Cartpole
swing-up, 2x[200] BNN policy + dynamics, 100 particles PER GPU (weak scaling: the particle axis shards,
one NCCL all-reduce of the 41,802-float policy gradient per iteration), H = 400, synthetic bounded fixture
of SURVEY.md section 8d with random-init weights.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CONFIGS = {
    # name: (env, D, U, maxU, hidden, particles per GPU, horizon)
    "c1": ("cartpole", 5, 1, 10.0, [200, 200], 25, 40),
    "c2": ("cartpole", 5, 1, 10.0, [200, 200], 100, 400),
    "c4": ("double_cartpole", 8, 1, 20.0, [400, 400, 400], 125, 600),
    "c5": ("cartpole", 5, 1, 10.0, [512, 512], 250, 1000),
}


def build_workload(cfg, n_global, device, seed=3):
    """Bounded synthetic fixture, built from this package's mirror modules."""
    from functools import partial
    import numpy as np
    from prob_mbrl_b200 import models, rewards
    env, D, U, maxU, hid, _, H = CONFIGS[cfg]
    torch.manual_seed(seed)
    np.random.seed(seed)
    reward = rewards.CartpoleReward() if env == "cartpole" else rewards.DoubleCartpoleReward()
    dyn_net = models.mlp(D + U, 2 * D, hid, dropout_layers=[models.CDropout(0.1 * torch.ones(h)) for h in hid])
    dyn = models.DynamicsModel(dyn_net, reward_func=reward, output_density=models.DiagGaussianDensity(D)).float()
    pol_net = models.mlp(D, 2 * U, hid, dropout_layers=[models.BDropout(0.1) for _ in hid],
                         output_nonlin=partial(models.DiagGaussianDensity, U))
    pol = models.Policy(pol_net, np.array([maxU]), np.array([-maxU])).float()
    g = torch.Generator().manual_seed(7)
    X = torch.randn(512, D + U, generator=g)
    X[:, -U:] *= maxU / 2
    Y = 1e-3 * torch.randn(512, D, generator=g)
    dyn.set_dataset(X, Y)
    dyn.eval()
    pol.train()
    x0 = 0.1 * torch.randn(n_global, D, generator=g)
    # allocate the [N, h] masks / z exactly like step 0 of a rollout would (on the CPU generator)
    from prob_mbrl_b200 import operands
    operands.materialize_noise(dyn, pol, x0)
    return dyn.to(device), pol.to(device), x0, H


def flop_model(cfg):
    """Algorithmic FLOPs per particle-step: fwd, bwd-data, policy wgrad."""
    _, D, U, _, hid, _, _ = CONFIGS[cfg]
    dims_p = [D] + hid + [2 * U]
    dims_d = [D + U] + hid + [2 * D]
    m_pol = sum(a * b for a, b in zip(dims_p[:-1], dims_p[1:]))
    m_dyn = sum(a * b for a, b in zip(dims_d[:-1], dims_d[1:]))
    return {"fwd": 2 * (m_pol + m_dyn), "bwd_data": 2 * (m_pol + m_dyn), "wgrad": 2 * m_pol,
            "total": 4 * (m_pol + m_dyn) + 2 * m_pol}


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            sm, mx, reasons = [], [], set()
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                for name, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            if sm:
                sm.sort()
                out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                       "samples": len(sm)}
        except Exception:
            pass
        finally:
            try:
                os.unlink(self.path)
            except Exception:
                pass
        return out


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"bf16_tflops": d.get("bf16_tflops"), "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d.get("hbm_gbs"), "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


# ----------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference algorithm
# ----------------------------------------------------------------------------------------------
def cpu_baseline(cfg, budget_s=15.0, threads=None, iters_min=2):
    """rollout-steps/s of the CPU oracle (plain PyTorch restatement of the reference's mc_pilco
    iteration: rollout + loss + autograd backward + clip + Adam) on a bounded sample of the workload:
    same networks and particle count, horizon cut to H_s so a few iterations fit the time budget
    (throughput of the reference is horizon-independent)."""
    from oracle import rollout_oracle as orc
    from prob_mbrl_b200 import operands
    _, D, U, _, hid, n, H = CONFIGS[cfg]
    dyn, pol, x0, _ = build_workload(cfg, n, "cpu")
    ops = operands.extract(dyn, pol, n).to_flat()
    ops = {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in ops.items()}
    best = None
    for nt in ([threads] if threads else sorted({1, os.cpu_count() or 1})):
        torch.set_num_threads(nt)
        Hs = min(H, 40)
        t0 = time.perf_counter()
        orc.mc_pilco_iterations(ops, x0, Hs, 1, 1e-4)            # warm-up
        one = time.perf_counter() - t0
        iters = max(iters_min, min(20, int(budget_s / 2 / max(one, 1e-3))))
        t0 = time.perf_counter()
        orc.mc_pilco_iterations(ops, x0, Hs, iters, 1e-4)
        dt = time.perf_counter() - t0
        val = n * Hs * iters / dt
        if best is None or val > best["value"]:
            best = {"value": val, "unit": "rollout-steps/s", "cores": nt, "kind": "port",
                    "sample": "%d iterations of the oracle mc_pilco step, N=%d, H=%d (of %d), %d thread(s)"
                              % (iters, n, Hs, H, nt)}
    return best


def gpu_eager_baseline(cfg, device, iters=3, Hs=40):
    """The same mc_pilco iteration as a plain PyTorch module loop + autograd on `device` (what the reference does
    with `--use_cuda`: ~220 ATen kernels per imagined step), on a bounded sample: same nets and particle count,
    horizon cut to Hs.  Context for the fused numbers, not a target."""
    import prob_mbrl_b200 as pm
    _, D, U, _, hid, n, H = CONFIGS[cfg]
    dyn, pol, x0, _ = build_workload(cfg, n, device)
    opt = torch.optim.Adam(pol.parameters(), 1e-4)
    old = os.environ.get("PROB_MBRL_BACKEND")
    os.environ["PROB_MBRL_BACKEND"] = "eager"
    try:
        kw = dict(pegasus=True, mm_states=False, mm_rewards=False, maximize=True, clip_grad=1.0,
                  resampling_period=10 ** 9, init_state_noise=0.0)
        x0 = x0.to(device)
        pm.mc_pilco(x0, dyn, pol, Hs, opt, None, 1, **kw)
        if torch.device(device).type == "cuda":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        pm.mc_pilco(x0, dyn, pol, Hs, opt, None, iters, **kw)
        if torch.device(device).type == "cuda":
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    finally:
        if old is None:
            os.environ.pop("PROB_MBRL_BACKEND", None)
        else:
            os.environ["PROB_MBRL_BACKEND"] = old
    return {"value": n * Hs * iters / dt, "unit": "rollout-steps/s", "kind": "PyTorch eager module loop + autograd on the GPU",
            "sample": "%d mc_pilco iterations, N=%d, H=%d (of %d)" % (iters, n, Hs, H)}


def run_reference_arm(args, cfg, rank, world):
    if rank != 0:
        return
    n = CONFIGS[cfg][5]
    cb = cpu_baseline(cfg, budget_s=max(10.0, min(120.0, 6.0 * (args.steps + args.warmup))))
    line = {
        "impl": "reference", "metric": "rollout-steps/sec (particles x horizon per mc_pilco iteration)",
        "value": cb["value"], "unit": "rollout-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, n, args.gpus),
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "rollout-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(cfg, n_per_gpu, gpus):
    env, D, U, maxU, hid, _, H = CONFIGS[cfg]
    return {"workload": "%s: %s swing-up MC-PILCO iteration, %dx[%d] BNN policy+dynamics, %d particles/GPU, H=%d"
                        % (cfg, env, len(hid), hid[0], n_per_gpu, H),
            "particles_global": n_per_gpu * gpus, "horizon": H, "parallelism": "particles sharded x%d" % gpus,
            "cache": "no explicit L2 flush: the per-iteration working set (activations kept for the reverse "
                     "sweep + deltas) is larger than the 126 MB L2"}


# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="fused", choices=["fused", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="value + per-kernel times only (tuning runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    cfg = args.config
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, cfg, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: the fused path needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    os.environ.setdefault("PMB_NO_PBAR", "1")

    import prob_mbrl_b200 as pm
    from prob_mbrl_b200 import _lib, dist

    n_per = CONFIGS[cfg][5]
    n_global = n_per * world
    dyn, pol, x0_all, H = build_workload(cfg, n_global, dev)
    opt = torch.optim.Adam(pol.parameters(), 1e-4)
    sharder = dist.ShardedNoise(dyn, pol, n_global, rank, world)
    sharder.narrow()
    row0 = sharder.row0
    x0_host = x0_all[row0:row0 + n_per].contiguous().pin_memory()
    x0_dev = x0_host.to(dev)
    g_r = torch.full((H, n_per), -1.0 / (H * n_global), device=dev)
    sync = dist.allreduce_gradient if world > 1 else None
    eng = pm.FusedIteration(dyn, pol, x0_dev, H, opt, g_r, 1.0, None, sync)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---------------- value: inputs resident in HBM ----------------
    clocks = ClockSampler(local_rank)
    clocks.start()
    t_spin = time.perf_counter()
    for _ in range(args.warmup):
        eng.step(x0_dev)
    while time.perf_counter() - t_spin < 1.0:      # let nvidia-smi start streaming; keeps the GPU under load
        eng.step(x0_dev)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        eng.step(x0_dev)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    t = torch.tensor([ms], device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t.item())
    value = n_global * H * args.steps / (ms * 1e-3)
    loss_val = float(eng.loss)

    if args.quick:
        print(json.dumps({"value": value, "ms_per_step": ms / args.steps, "n_gpus": world,
                          "tuning": {k: os.environ.get(k) for k in ("PMB_STREAM_MODE", "PMB_PARTICLES_PER_CTA")}}))
        return

    # ---------------- e2e: public API, host buffers, H2D + D2H inside the timed region ----------------
    class HostStates:
        """Stand-in for ExperienceDataset.sample_states (reference utils/experience_dataset.py:236-249):
        hands mc_pilco a fresh batch of initial particles in pinned HOST memory every iteration."""

        def __init__(self, x):
            self.x = x

        def sample_states(self, n, timestep=0):
            return self.x

    os.environ["PMB_PBAR_EVERY"] = "1"      # progress read-back of the predicted return every iteration
    x0_glob_host = x0_all.contiguous().pin_memory()
    host = HostStates(x0_glob_host)
    sharder.widen()
    kw = dict(pegasus=True, mm_states=False, mm_rewards=False, maximize=True, clip_grad=1.0,
              resampling_period=10 ** 9, init_state_noise=0.0)
    # no PEGASUS resample inside the timed window (the counter persists across calls like the reference's)
    sys.modules["prob_mbrl_b200.mc_pilco"].policy_update_counter[pol] = 1
    pm.mc_pilco(x0_glob_host, dyn, pol, H, opt, host, args.warmup, **kw)
    barrier()
    t0 = time.perf_counter()
    pm.mc_pilco(x0_glob_host, dyn, pol, H, opt, host, args.steps, **kw)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    e2e_value = n_global * H * args.steps / float(t.item())

    # ---------------- roofline of the dominant kernel (rank 0) ----------------
    roof, kern = None, {}
    if rank == 0:
        fm = flop_model(cfg)
        peaks = measured_peaks()

        def time_phase(fn, reps=5):
            fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps

        lib = eng.lib
        st = _lib.current_stream_ptr()
        pb = C.byref(eng.prob)

        def fwd(ph):
            tune = _lib.make_tuning(phases=ph)
            return lambda: _lib.check(lib.pmb_rollout_forward(
                pb, C.byref(tune), eng.x0.data_ptr(), eng.states.data_ptr(), eng.actions.data_ptr(),
                eng.rewards.data_ptr(), eng.ws.data_ptr(), eng.nbytes, eng.status.data_ptr(), st))

        def bwd(ph):
            tune = _lib.make_tuning(phases=ph)
            return lambda: _lib.check(lib.pmb_rollout_backward(
                pb, C.byref(tune), eng.states.data_ptr(), eng.actions.data_ptr(), eng.rewards.data_ptr(), None,
                None, eng.g_rewards.data_ptr(), eng.grad_flat.data_ptr(), eng.dx0.data_ptr(), eng.ws.data_ptr(),
                eng.nbytes, st))

        fwd(7)()
        kern = {"pack_ms": time_phase(fwd(1)), "fwd_sweep_ms": time_phase(fwd(2)),
                "bwd_sweep_ms": time_phase(bwd(2)), "wgrad_ms": time_phase(bwd(4))}
        work = n_per * H
        dom = "bwd_sweep_ms" if kern["bwd_sweep_ms"] >= kern["fwd_sweep_ms"] else "fwd_sweep_ms"
        flops = fm["bwd_data" if dom == "bwd_sweep_ms" else "fwd"] * work
        achieved = flops / (kern[dom] * 1e-3) / 1e12
        peak = peaks["bf16_tflops"]
        plan = _lib.describe_plan(eng.prob, eng.tune)       # sweep variant + launch geometry the planner chose
        ctas = plan["ctas"]
        kname = ("cluster_" if plan["variant"] == 1 else "rollout_") + ("bwd_kernel" if dom == "bwd_sweep_ms" else "fwd_kernel")
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_dram_traffic.json")      # dram__bytes_read+write per launch, from
        if cfg == "c2" and os.path.exists(tpath):                             # the committed `ncu --set full` capture
            traffic = json.load(open(tpath)).get(kname)
        roof = {"bound": "tensor", "kernel": kname,
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_source": "%s bf16 burst (kernel timed alone)" % peaks["source"], "traffic": traffic,
                "algorithmic_flops_per_launch": flops, "launch_ms": kern[dom],
                "sweeps": {"variant": "cluster-resident" if plan["variant"] == 1 else "streaming", "ctas": ctas,
                           "cluster_size": plan["cluster_size"], "particles_per_group": plan["particles_per_group"],
                           "smem_bytes_per_cta": plan["smem_bwd_bytes" if dom == "bwd_sweep_ms" else "smem_fwd_bytes"]},
                "sms_occupied": min(ctas, 148),
                "frac_of_occupied_sm_peak": achieved / (peak * min(ctas, 148) / 148.0),
                "arithmetic": "fp32 FFMA (CUDA cores); fp32 SIMT peak of the occupied SMs = %.2f TFLOP/s"
                              % (min(ctas, 148) * 128 * 2 * 1.9e9 / 1e12)}

    launches_per_iter = None
    try:
        pi = _lib.describe_plan(eng.prob, eng.tune)
        launches_per_iter = pi["launches_fwd"] + pi["launches_bwd"] + 2
    except Exception:
        pass
    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline(cfg)

    if rank == 0:
        nlin = len(CONFIGS[cfg][4]) + 1
        line = {
            "metric": "rollout-steps/sec (particles x horizon per mc_pilco iteration)",
            "value": value, "unit": "rollout-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(cfg, n_per, world),
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "rollout-steps/s",
                    "h2d_bytes_per_step": int(x0_glob_host.numel() * 4), "d2h_bytes_per_step": 4,
                    "api": "prob_mbrl_b200.mc_pilco(x0_host, dynamics, policy, H, opt, exp, K, pegasus=True)"},
            # per iteration: pack + forward sweep; [adjoint-factor pre-pass] + reverse sweep + one weight-gradient
            # kernel per policy layer + partial reduction; gradient norm + Adam
            "gpu_launches": args.steps * (launches_per_iter if launches_per_iter else 3 + nlin + 3),
            "kernels_ms": kern, "roofline": roof, "loss": loss_val,
        }
        if cb is not None:
            line["cpu_baseline"] = cb
            try:
                line["gpu_eager_baseline"] = gpu_eager_baseline(cfg, dev)
            except Exception as e:                      # context only: never fail the bench line over it
                line["gpu_eager_baseline"] = {"unavailable": "%s: %s" % (type(e).__name__, e)}
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
