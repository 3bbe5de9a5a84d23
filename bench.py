#!/usr/bin/env python
"""Benchmark of the imagined-rollout hot path (BASELINE.json metric: rollout-steps/sec).

  python bench.py --gpus N --steps K --warmup W            fused sm_100a path (this repo)
  python bench.py --impl reference --gpus N --steps K ...  the UNMODIFIED reference's algorithms.mc_pilco on the
                                                           box's host cores (baseline/reference_arm.py)

A "step" is ONE mc_pilco policy-gradient iteration (weight packing, forward sweep over H imagined steps,
reverse sweep, batched policy weight gradient, gradient clip, Adam) on a batch of particles;
rollout-steps/sec = particles x horizon x K / time.  Workload = BASELINE.json configs[1] (c2): Cartpole
swing-up, 2x[200] BNN policy + dynamics, 100 particles PER GPU (weak scaling: the particle axis shards,
one NCCL all-reduce of the 41,802-float policy gradient per iteration), H = 400, synthetic bounded fixture
of SURVEY.md section 8d with random-init weights.  The other BASELINE configs (c1, c3, one GPU's shard of c4 and
c5; the full c4 / c5 when launched on 4 / 8 GPUs) are measured after it and reported under `other_configs`.
"""
import argparse
import ctypes as C
import datetime
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
    # name: (env, D, U, maxU, hidden, particles per GPU, horizon, moment matching)
    "c1": ("cartpole", 5, 1, 10.0, [200, 200], 25, 40, False),
    "c2": ("cartpole", 5, 1, 10.0, [200, 200], 100, 400, False),
    "c3": ("cartpole", 5, 1, 10.0, [200, 200], 100, 400, True),
    "c4": ("double_cartpole", 8, 1, 20.0, [400, 400, 400], 125, 600, False),
    "c5": ("cartpole", 5, 1, 10.0, [512, 512], 250, 1000, False),
}
# world size at which BASELINE.json quotes the config (c4: 500 particles on 4 GPUs, c5: 2000 on 8)
NATIVE_WORLD = {"c1": 1, "c2": 1, "c3": 1, "c4": 4, "c5": 8}


def whiten_rows(z, N):
    """Zero-mean, identity-sample-covariance copy of the first N rows (the reference's
    moment-matched rollout is explosive by construction with a raw table, App. D-7)."""
    zz = z[:N].double()
    zz = zz - zz.mean(0, keepdim=True)
    L = torch.linalg.cholesky(zz.T @ zz / (N - 1))
    zz = torch.linalg.solve_triangular(L, zz.T, upper=False).T
    out = z.clone()
    out[:N] = zz.float()
    return out


def build_workload(cfg, n_global, device, seed=3):
    """Bounded synthetic fixture, built from this package's mirror modules.
    Returns (dynamics, policy, x0 [n_global, D] on the CPU, H, mm) with mm = None or the moment-matching tables."""
    from functools import partial
    import numpy as np
    from prob_mbrl_b200 import models, rewards
    env, D, U, maxU, hid, _, H, with_mm = CONFIGS[cfg]
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
    mm = None
    if with_mm:
        z_mm = whiten_rows(torch.randn(H + n_global, D, generator=g), n_global)
        z_rr = torch.randn(H + n_global, 1, generator=g)
        mm = dict(mm_states=True, mm_rewards=True, mm_groups=None, z_mm=z_mm.to(device), z_rr=z_rr.to(device))
    # allocate the [N, h] masks / z exactly like step 0 of a rollout would (on the CPU generator)
    from prob_mbrl_b200 import operands
    operands.materialize_noise(dyn, pol, x0)
    return dyn.to(device), pol.to(device), x0, H, mm


def flop_model(cfg):
    """Algorithmic FLOPs per particle-step: fwd, bwd-data, policy wgrad."""
    _, D, U, _, hid = CONFIGS[cfg][:5]
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
# CPU baselines
# ----------------------------------------------------------------------------------------------
def port_baseline(cfg, n_global, budget_s=10.0, iters_min=2):
    """rollout-steps/s of the CPU ORACLE PORT (oracle/rollout_oracle.py: plain PyTorch restatement of the iteration on
    extracted operands; it skips the dynamics weight gradients the reference wastes time on and issues ~4x fewer ATen
    ops per step, i.e. a FASTER baseline than the reference itself), same nets and particle count, horizon cut."""
    from oracle import rollout_oracle as orc
    from prob_mbrl_b200 import operands
    H, with_mm = CONFIGS[cfg][6], CONFIGS[cfg][7]
    dyn, pol, x0, _, mm = build_workload(cfg, n_global, "cpu")
    ops = operands.extract(dyn, pol, n_global).to_flat()
    ops = {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in ops.items()}
    kw = dict(mm_states=True, mm_rewards=True, z_mm=mm["z_mm"], z_rr=mm["z_rr"]) if with_mm else {}
    best = None
    for nt in sorted({1, os.cpu_count() or 1}):
        torch.set_num_threads(nt)
        Hs = min(H, 40)
        t0 = time.perf_counter()
        orc.mc_pilco_iterations(ops, x0, Hs, 1, 1e-4, **kw)            # warm-up
        one = time.perf_counter() - t0
        iters = max(iters_min, min(20, int(budget_s / 2 / max(one, 1e-3))))
        t0 = time.perf_counter()
        orc.mc_pilco_iterations(ops, x0, Hs, iters, 1e-4, **kw)
        dt = time.perf_counter() - t0
        val = n_global * Hs * iters / dt
        if best is None or val > best["value"]:
            best = {"value": val, "unit": "rollout-steps/s", "cores": nt, "kind": "port",
                    "sample": "%d iterations of the oracle mc_pilco step, N=%d, H=%d (of %d), %d thread(s)"
                              % (iters, n_global, Hs, H, nt)}
    return best


def reference_baseline(cfg, n_global, iters, budget_s=25.0):
    """The unmodified reference's algorithms.mc_pilco on the host cores at torch.set_num_threads(1) (the examples'
    default, reference examples/deep_pilco_no_mm.py:21,65) AND at all cores, reported separately; `value` = the
    faster of the two.  None when the reference package is not present (baseline/install_reference.sh)."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import reference_arm
    if not reference_arm.available():
        return None
    env, _, _, _, hid, _, H, with_mm = CONFIGS[cfg]
    runs = []
    for nt in sorted({1, os.cpu_count() or 1}):
        runs.append(reference_arm.time_reference(env, hid, n_global, H, with_mm, nt, iters, budget_s=budget_s / 2))
    best = max(runs, key=lambda r: r["value"])
    out = dict(best)
    out["threads_1"] = runs[0]["value"]
    out["threads_all"] = runs[-1]["value"]
    out["host_cores"] = os.cpu_count()
    out["runs"] = [{k: r[k] for k in ("cores", "value", "iterations", "horizon_timed", "ms_per_iteration")} for r in runs]
    return out


def gpu_eager_baseline(cfg, device, iters=3, Hs=40):
    """The same mc_pilco iteration as a plain PyTorch module loop + autograd on `device` (what the reference does
    with `--use_cuda`: ~220 ATen kernels per imagined step), on a bounded sample: same nets and particle count,
    horizon cut to Hs.  Context for the fused numbers, not a target."""
    import prob_mbrl_b200 as pm
    n, H = CONFIGS[cfg][5], CONFIGS[cfg][6]
    dyn, pol, x0, _, _ = build_workload(cfg, n, device)
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


def workload_config(cfg, n_per_gpu, gpus):
    env, D, U, maxU, hid, _, H, with_mm = CONFIGS[cfg]
    return {"workload": "%s: %s swing-up MC-PILCO iteration%s, %dx[%d] BNN policy+dynamics, %d particles/GPU, H=%d"
                        % (cfg, env, " with moment matching" if with_mm else "", len(hid), hid[0], n_per_gpu, H),
            "particles_global": n_per_gpu * gpus, "horizon": H, "parallelism": "particles sharded x%d" % gpus,
            "cache": "no explicit L2 flush: the per-iteration working set (activations kept for the reverse "
                     "sweep + deltas) is larger than the 126 MB L2"}


def run_reference_arm(args, cfg, rank, world):
    """`--impl reference`: the reference's own CPU implementation at the SAME global particle count as the fused arm
    (n_per_gpu x --gpus), on rank 0 only."""
    if rank != 0:
        return
    n_per = CONFIGS[cfg][5]
    n_global = n_per * args.gpus
    budget = max(20.0, min(150.0, 8.0 * (args.steps + args.warmup)))
    cb = reference_baseline(cfg, n_global, iters=args.steps, budget_s=budget)
    port = None
    if cb is None:       # reference package absent: fall back to the oracle port and say so
        cb = port_baseline(cfg, n_global, budget_s=budget / 2)
    else:
        try:
            port = port_baseline(cfg, n_global, budget_s=10.0)
        except Exception as e:
            port = {"unavailable": "%s: %s" % (type(e).__name__, e)}
    conf = workload_config(cfg, n_per, args.gpus)
    conf["timed"] = cb["sample"]
    line = {
        "impl": "reference", "metric": "rollout-steps/sec (particles x horizon per mc_pilco iteration)",
        "value": cb["value"], "unit": "rollout-steps/s", "n_gpus": args.gpus, "steps": cb.get("iterations", args.steps),
        "warmup": args.warmup, "ms_per_step": cb.get("ms_per_iteration"), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": conf,
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "rollout-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if port is not None:
        line["cpu_port"] = port
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# the fused arm
# ----------------------------------------------------------------------------------------------
class Harness:
    def __init__(self, args, rank, world, local_rank, dev):
        self.args, self.rank, self.world, self.local_rank, self.dev = args, rank, world, local_rank, dev

    def barrier(self):
        if self.world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = torch.tensor([float(x)], device=self.dev)
        if self.world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    def make_engine(self, cfg, n_per, world):
        """Engine on this rank's shard of a `n_per x world` particle batch (world = 1: no collective)."""
        import prob_mbrl_b200 as pm
        from prob_mbrl_b200 import dist
        n_global = n_per * world
        dyn, pol, x0_all, H, mm = build_workload(cfg, n_global, self.dev)
        opt = torch.optim.Adam(pol.parameters(), 1e-4)
        rank = self.rank if world > 1 else 0
        sharder = dist.ShardedNoise(dyn, pol, n_global, rank, world)
        sharder.narrow()
        row0 = sharder.row0
        x0_dev = x0_all[row0:row0 + n_per].contiguous().to(self.dev)
        g_r = torch.full((H, n_per), -1.0 / (H * n_global), device=self.dev)
        sync = "auto" if world > 1 else None       # peer-memory exchange (PMB_GRAD_SYNC=nccl: NCCL all-reduce)
        eng = pm.FusedIteration(dyn, pol, x0_dev, H, opt, g_r, 1.0, mm, sync)
        return dict(eng=eng, dyn=dyn, pol=pol, opt=opt, x0_all=x0_all, x0_dev=x0_dev, H=H, mm=mm, sharder=sharder,
                    n_per=n_per, n_global=n_global, world=world)

    def time_engine(self, w, steps, warmup):
        """K graph-replayed iterations with x0 resident, CUDA events, max over ranks.  Every rank issues exactly the
        same number of iterations (hence collectives): the spin-up count that lets nvidia-smi start streaming is
        derived from an all-reduced step time, never from a rank-local wall clock."""
        eng, x0 = w["eng"], w["x0_dev"]
        clocks = ClockSampler(self.local_rank)
        clocks.start()
        for _ in range(warmup):
            eng.step(x0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.step(x0)
        b.record()
        torch.cuda.synchronize()
        one_ms = self.max_over_ranks(a.elapsed_time(b)) if w["world"] > 1 else a.elapsed_time(b)
        for _ in range(int(min(2000, max(1, 1000.0 / max(one_ms, 1e-3))))):      # ~1 s under load, rank-agreed count
            eng.step(x0)
        if w["world"] > 1:
            self.barrier()
        else:
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            eng.step(x0)
        e1.record()
        if w["world"] > 1:
            self.barrier()
        else:
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        clk = clocks.stop()
        if w["world"] > 1:
            ms = self.max_over_ranks(ms)
        if w["mm"] is not None and int(eng.status.item()) != 0:
            raise RuntimeError("moment matching failed (status %d) inside the timed window" % int(eng.status.item()))
        return ms, clk

    def kernel_times(self, w):
        """Per-phase CUDA-event times (pack / forward sweep / reverse sweep / weight gradient) on this rank."""
        from prob_mbrl_b200 import _lib
        eng = w["eng"]
        lib = eng.lib
        st = _lib.current_stream_ptr()
        pb = C.byref(eng.prob)

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

        def fwd(ph):
            tune = _lib.make_tuning(phases=ph)
            return lambda: _lib.check(lib.pmb_rollout_forward(
                pb, C.byref(tune), eng.x0.data_ptr(), eng.states.data_ptr(), eng.actions.data_ptr(),
                eng.rewards.data_ptr(), eng.ws.data_ptr(), eng.nbytes, eng.status.data_ptr(), st))

        def bwd(ph):
            tune = _lib.make_tuning(phases=ph)
            return lambda: _lib.check(lib.pmb_rollout_backward(
                pb, C.byref(tune), eng.states.data_ptr(), eng.actions.data_ptr(), eng.rewards.data_ptr(), None,
                None, eng.g_rewards.data_ptr(), eng.grad_flat.data_ptr(), eng.dx0.data_ptr(), None, eng.ws.data_ptr(),
                eng.nbytes, st))

        fwd(7)()
        return {"pack_ms": time_phase(fwd(1)), "fwd_sweep_ms": time_phase(fwd(2)),
                "bwd_sweep_ms": time_phase(bwd(2)), "wgrad_ms": time_phase(bwd(4))}

    def roofline(self, cfg, w, kern):
        from prob_mbrl_b200 import _lib
        eng = w["eng"]
        fm = flop_model(cfg)
        peaks = measured_peaks()
        work = w["n_per"] * w["H"]
        dom = "bwd_sweep_ms" if kern["bwd_sweep_ms"] >= kern["fwd_sweep_ms"] else "fwd_sweep_ms"
        flops = fm["bwd_data" if dom == "bwd_sweep_ms" else "fwd"] * work
        achieved = flops / (kern[dom] * 1e-3) / 1e12
        peak = peaks["bf16_tflops"]
        plan = _lib.describe_plan(eng.prob, eng.tune)       # sweep variant + launch geometry the planner chose
        ctas = plan["ctas"]
        vname = {0: "streaming", 1: "cluster-resident", 2: "tensor-core cluster",
                 3: "wide cluster-resident"}.get(plan["variant"], str(plan["variant"]))
        kprefix = {0: "rollout_", 1: "cluster_", 2: "tc_", 3: "cw_"}.get(plan["variant"], "")
        kname = kprefix + ("bwd_kernel" if dom == "bwd_sweep_ms" else "fwd_kernel")
        traffic = None
        for tname in ("r02_dram_traffic.json", "r01_dram_traffic.json"):   # dram__bytes_read+write per launch, from
            tpath = os.path.join(ROOT, "profiles", tname)                   # the committed `ncu --set full` captures
            if os.path.exists(tpath):
                traffic = json.load(open(tpath)).get("%s:%s" % (cfg, kname))
                if traffic is None and cfg == "c2":
                    traffic = json.load(open(tpath)).get(kname)
                if traffic is not None:
                    break
        sms = min(ctas, 148)
        simt_peak = sms * 128 * 2 * 1.9e9 / 1e12         # fp32 FMA lanes of the occupied SMs at the nominal boost clock
        return {"bound": "tensor", "kernel": kname,
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_source": "%s bf16 burst (kernel timed alone)" % peaks["source"], "traffic": traffic,
                "algorithmic_flops_per_launch": flops, "launch_ms": kern[dom],
                "sweeps": {"variant": vname, "ctas": ctas,
                           "cluster_size": plan["cluster_size"], "particles_per_group": plan["particles_per_group"],
                           "smem_bytes_per_cta": plan["smem_bwd_bytes" if dom == "bwd_sweep_ms" else "smem_fwd_bytes"]},
                "sms_occupied": sms,
                "frac_of_occupied_sm_peak": achieved / (peak * sms / 148.0),
                "frac_of_fp32_simt_peak_of_occupied_sms": achieved / simt_peak,
                "arithmetic": ("3xTF32 split on tcgen05 (fp32 accumulate in TMEM) for the hidden x hidden layers, fp32 FFMA elsewhere"
                               if plan["variant"] == 2 else "fp32 FFMA2 (CUDA cores)")
                              + "; fp32 SIMT peak of the occupied SMs = %.2f TFLOP/s" % simt_peak}

    def launches_per_iter(self, w):
        from prob_mbrl_b200 import _lib
        try:
            pi = _lib.describe_plan(w["eng"].prob, w["eng"].tune)
            return pi["launches_fwd"] + pi["launches_bwd"] + 2
        except Exception:
            return None

    def measure_config(self, cfg, world, steps, warmup):
        """value / ms_per_step / roofline / clocks of one config (device-resident inputs)."""
        n_per = CONFIGS[cfg][5]
        w = self.make_engine(cfg, n_per, world)
        ms, clk = self.time_engine(w, steps, warmup)
        out = {"config": workload_config(cfg, n_per, world), "n_gpus": world, "steps": steps,
               "value": w["n_global"] * w["H"] * steps / (ms * 1e-3), "unit": "rollout-steps/s",
               "ms_per_step": ms / steps, "clocks": clk, "loss": float(w["eng"].loss)}
        # per-kernel times are rank 0's; with moment matching across ranks the sweeps of every rank take part in the
        # per-step exchange, so every rank has to run the same phase launches
        all_ranks = w["mm"] is not None and world > 1
        if self.rank == 0 or all_ranks:
            kern = self.kernel_times(w)
            if all_ranks:
                self.barrier()
        if self.rank == 0:
            out["kernels_ms"] = kern
            out["roofline"] = self.roofline(cfg, w, kern)
        w["sharder"].widen()
        return out, w


def e2e_through_mc_pilco(h, w, steps, warmup):
    """The public mc_pilco(x0_host, dynamics, policy, H, opt, exp, K) call: per iteration x0 comes from pinned HOST
    memory (H2D, like exp.sample_states(...).to(device)) and the predicted return is read back for the progress line
    (D2H, like reference algorithms/mc_pilco.py:215-216)."""
    import prob_mbrl_b200 as pm

    class HostStates:
        """Stand-in for ExperienceDataset.sample_states (reference utils/experience_dataset.py:236-249):
        hands mc_pilco a fresh batch of initial particles in pinned HOST memory every iteration."""

        def __init__(self, x):
            self.x = x

        def sample_states(self, n, timestep=0):
            return self.x

    os.environ["PMB_PBAR_EVERY"] = "1"      # progress read-back of the predicted return every iteration
    x0_glob_host = w["x0_all"].contiguous().pin_memory()
    host = HostStates(x0_glob_host)
    w["sharder"].widen()
    dyn, pol, opt, H = w["dyn"], w["pol"], w["opt"], w["H"]
    kw = dict(pegasus=True, mm_states=False, mm_rewards=False, maximize=True, clip_grad=1.0,
              resampling_period=10 ** 9, init_state_noise=0.0)
    # no PEGASUS resample inside the timed window (the counter persists across calls like the reference's)
    sys.modules["prob_mbrl_b200.mc_pilco"].policy_update_counter[pol] = 1
    pm.mc_pilco(x0_glob_host, dyn, pol, H, opt, host, warmup, **kw)
    h.barrier()
    t0 = time.perf_counter()
    pm.mc_pilco(x0_glob_host, dyn, pol, H, opt, host, steps, **kw)
    h.barrier()
    e2e_s = h.max_over_ranks(time.perf_counter() - t0)
    return {"value": w["n_global"] * H * steps / e2e_s, "unit": "rollout-steps/s",
            "h2d_bytes_per_step": int(x0_glob_host.numel() * 4), "d2h_bytes_per_step": 4,
            "api": "prob_mbrl_b200.mc_pilco(x0_host, dynamics, policy, H, opt, exp, K, pegasus=True)"}


def e2e_through_rollout(h, w, steps, warmup):
    """c3 (moment matching): mc_pilco draws its own z_mm table, with which the reference algorithm is explosive at
    H = 400, so the end-to-end number goes through the public rollout() call on the whitened
    table, as SURVEY.md section 8d prescribes: x0 from pinned host memory, rollout(), loss, backward (autograd
    through the fused rollout), clip, Adam, loss read back."""
    import prob_mbrl_b200 as pm
    dyn, pol, opt, H, mm = w["dyn"], w["pol"], w["opt"], w["H"], w["mm"]
    w["sharder"].widen()
    x0_host = w["x0_all"].contiguous().pin_memory()

    def one():
        x0 = x0_host.to(h.dev, non_blocking=True)
        opt.zero_grad()
        S, A, R = pm.rollout(x0, dyn, pol, H, resample_state_noise=False, resample_action_noise=False,
                             mm_states=True, mm_rewards=True, z_mm=mm["z_mm"], z_rr=mm["z_rr"])
        loss = -(torch.stack(R).sum(0) / H).mean()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(pol.parameters(), 1.0)
        opt.step()
        return float(loss)

    for _ in range(warmup):
        one()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"value": w["n_global"] * H * steps / dt, "unit": "rollout-steps/s",
            "h2d_bytes_per_step": int(x0_host.numel() * 4), "d2h_bytes_per_step": 4,
            "api": "prob_mbrl_b200.rollout(x0_host.to(dev), dynamics, policy, H, mm_states=True, mm_rewards=True, "
                   "z_mm=whitened, z_rr=...) + loss.backward() + clip_grad_norm_ + Adam.step()"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="fused", choices=["fused", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--quick", action="store_true", help="value + per-kernel times + clocks only (tuning runs)")
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
        # a hung collective must fail in two minutes with a stack, not sit on the GPUs for ten
        os.environ.setdefault("TORCH_NCCL_DUMP_ON_TIMEOUT", "1")
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "1")
        torch.distributed.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    os.environ.setdefault("PMB_NO_PBAR", "1")
    h = Harness(args, rank, world, local_rank, dev)

    # ---------------- value: inputs resident in HBM ----------------
    main_res, w = h.measure_config(cfg, world, args.steps, args.warmup)
    if args.quick:
        if rank == 0:
            print(json.dumps({"value": main_res["value"], "ms_per_step": main_res["ms_per_step"], "n_gpus": world,
                              "config": cfg, "clocks": main_res["clocks"], "kernels_ms": main_res.get("kernels_ms"),
                              "sweeps": (main_res.get("roofline") or {}).get("sweeps"),
                              "tuning": {k: os.environ.get(k) for k in ("PMB_STREAM_MODE", "PMB_PARTICLES_PER_CTA")}}))
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    # ---------------- e2e: public API, host buffers, H2D + D2H inside the timed region ----------------
    if CONFIGS[cfg][7]:
        e2e = e2e_through_rollout(h, w, args.steps, args.warmup)
    else:
        e2e = e2e_through_mc_pilco(h, w, args.steps, args.warmup)

    # ---------------- the other BASELINE configs ----------------
    others = {}
    if not args.no_other_configs:
        todo = []
        if world == 1:
            todo = [c for c in ("c1", "c3", "c4", "c5") if c != cfg]        # c4 / c5: one GPU's shard
        else:
            todo = [c for c in ("c4", "c5") if NATIVE_WORLD[c] == world and c != cfg]   # the full sharded config
        for c in todo:
            try:
                res, wc = h.measure_config(c, world, max(3, min(args.steps, 10)), 3)
                del wc
                torch.cuda.empty_cache()
                if NATIVE_WORLD[c] != world:
                    res["note"] = "one GPU's shard of the %d-GPU config" % NATIVE_WORLD[c]
                others[c] = res
            except Exception as e:          # never lose the main line over an extra
                if world > 1:
                    raise                   # a rank-local skip would desynchronise the collectives
                others[c] = {"unavailable": "%s: %s" % (type(e).__name__, e)}

    cb, port, eager = None, None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = reference_baseline(cfg, CONFIGS[cfg][5], iters=10, budget_s=24.0)
        port = port_baseline(cfg, CONFIGS[cfg][5], budget_s=8.0)
        if cb is None:
            cb, port = port, None
        try:
            eager = gpu_eager_baseline(cfg, dev)
        except Exception as e:                      # context only: never fail the bench line over it
            eager = {"unavailable": "%s: %s" % (type(e).__name__, e)}

    if rank == 0:
        nlin = len(CONFIGS[cfg][4]) + 1
        lpi = h.launches_per_iter(w)
        line = {
            "metric": "rollout-steps/sec (particles x horizon per mc_pilco iteration)",
            "value": main_res["value"], "unit": "rollout-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": main_res["config"],
            "clocks": main_res["clocks"],
            "e2e": e2e,
            # per iteration: pack + forward sweep; [adjoint-factor pre-pass] + reverse sweep + one weight-gradient
            # kernel per policy layer + partial reduction; gradient norm + Adam
            "gpu_launches": args.steps * (lpi if lpi else 3 + nlin + 3),
            "kernels_ms": main_res.get("kernels_ms"), "roofline": main_res.get("roofline"), "loss": main_res["loss"],
        }
        if others:
            line["other_configs"] = others
        if cb is not None:
            line["cpu_baseline"] = cb
        if port is not None:
            line["cpu_port"] = port
        if eager is not None:
            line["gpu_eager_baseline"] = eager
        print(json.dumps(line))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
