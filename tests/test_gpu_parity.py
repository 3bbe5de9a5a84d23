"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, entered through the C ABI
(libpmb_b200.so via ctypes), against (a) the golden fixtures produced by the unmodified reference and
(b) the CPU oracle on the same seeded inputs.  Tolerances are the fp32 budgets of SURVEY.md App. C.3
(no-mm H<=40: states/rewards atol 2e-6, actions atol 2e-5, loss rtol 1e-6, policy-grad rel-L2 1e-5;
H=400 bounded fixture: grad rel-L2 1e-4)."""
import os

import pytest
import torch

import golden_util as gu
from oracle import rollout_oracle as orc

pytestmark = pytest.mark.gpu


def _ops_cuda(ops):
    from prob_mbrl_b200.operands import RolloutOperands
    return RolloutOperands.from_flat(ops, device="cuda")


def _run(ops, x0, H, cot="loss", env=None, gen_seed=0, mm=None, cot_mask=None):
    """forward + backward through the fused autograd.Function; returns dict of cpu tensors."""
    from prob_mbrl_b200.rollout import FusedRolloutFunction
    old = {}
    for k, v in (env or {}).items():
        old[k] = os.environ.get(k)
        os.environ[k] = str(v)
    try:
        o = _ops_cuda(ops)
        params = [p.requires_grad_(True) for p in o.policy_parameters()]
        x = x0.cuda().clone().requires_grad_(True)
        N = x.shape[0]
        if mm is None:
            mm = dict(mm_states=False, mm_rewards=False, mm_groups=None, z_mm=None, z_rr=None)
        else:
            mm = dict(mm, z_mm=mm["z_mm"].cuda().contiguous(), z_rr=mm["z_rr"].cuda().contiguous())
        S, A, R, status = FusedRolloutFunction.apply(x, (o, N, H, mm), *params)
        if cot == "loss":
            obj = -(R.sum(0) / H).mean()
            cots = None
        else:
            g = torch.Generator().manual_seed(gen_seed)
            gS = torch.randn(H + 1, N, o.D, generator=g, dtype=torch.float64)
            gA = torch.randn(H, N, o.U, generator=g, dtype=torch.float64)
            gR = torch.randn(H, N, generator=g, dtype=torch.float64)
            if cot_mask is not None:     # particles whose cotangents are zeroed: they contribute to no gradient
                gS, gA, gR = gS * cot_mask[None, :, None], gA * cot_mask[None, :, None], gR * cot_mask[None, :]
            obj = (S * gS.float().cuda()).sum() + (A * gA.float().cuda()).sum() + (R * gR.float().cuda()).sum()
            cots = (gS, gA, gR)
        grads = torch.autograd.grad(obj, params + [x])
        torch.cuda.synchronize()
        return {"status": int(status.item()),
                "S": S.detach().cpu(), "A": A.detach().cpu(), "R": R.detach().cpu(), "obj": obj.detach().cpu(),
                "grads": [g.cpu() for g in grads[:-1]], "dx0": grads[-1].cpu(), "cots": cots}
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


# sweep variants behind the same C ABI: "ring" = streaming sweeps (hidden x hidden weights through a TMA ring),
# "cluster" = cluster-resident sweeps (weights in the shared memory of a thread-block cluster; two hidden layers)
# "tc" = tensor-core cluster sweeps (tcgen05 3xTF32 hidden x hidden layers, 16-CTA cluster per 128-particle tile)
# "cw" = wide cluster-resident sweeps (two-hidden-layer nets up to 512 wide resident in a 16-CTA cluster, up to 36
#        particles per cluster, masks and gates as bit words)
SWEEPS = {"ring": {"PMB_STREAM_MODE": 2}, "cluster": {"PMB_STREAM_MODE": 3}, "tc": {"PMB_STREAM_MODE": 4},
          "cw": {"PMB_STREAM_MODE": 5}}


@pytest.mark.parametrize("sweeps", ["ring", "cluster", "tc", "cw"])
@pytest.mark.parametrize("name", ["cartpole_200x2_n25_h40", "dcartpole_48x3_n24_h30", "cartpole_37x2_n7_h12"])
def test_rollout_and_gradient_match_reference_golden(name, sweeps):
    if sweeps in ("cluster", "cw") and "x3" in name:
        pytest.skip("three hidden layers: streaming sweeps only")
    ops, g = gu.load(name)
    H = int(g["H"])
    r = _run(ops, g["x0"], H, env=SWEEPS[sweeps])
    assert (r["S"] - g["nomm_states"]).abs().max() < 2e-6
    assert (r["A"] - g["nomm_actions"]).abs().max() < 2e-5
    assert (r["R"] - g["nomm_rewards"]).abs().max() < 2e-6
    assert abs(float(r["obj"]) - float(g["nomm_loss"])) <= 1e-6 * abs(float(g["nomm_loss"])) + 1e-8
    assert gu.rel_l2(r["grads"], gu.policy_grad_list(g, "nomm", ops)) < 1e-5
    assert gu.rel_l2(r["dx0"], g["nomm_dx0"]) < 1e-5


@pytest.mark.parametrize("sweeps", ["ring", "cluster", "tc", "cw"])
def test_c2_full_size_matches_reference_golden(sweeps):
    """BASELINE.json configs[1]: Cartpole 2x[200], 100 particles, H=400."""
    ops, g = gu.load("cartpole_200x2_n100_h400")
    H, thin = int(g["H"]), int(g["thin"])
    r = _run(ops, g["x0"], H, env=SWEEPS[sweeps])
    assert (r["S"][::thin] - g["nomm_states"]).abs().max() < 5e-5
    assert abs(float(r["obj"]) - float(g["nomm_loss"])) <= 1e-6 * abs(float(g["nomm_loss"]))
    assert gu.rel_l2(r["grads"], gu.policy_grad_list(g, "nomm", ops)) < 1e-4
    # error budget against the fp64 oracle: within a small multiple of the reference's own fp32 error
    ops64, g64 = gu.load("cartpole_200x2_n100_h400", torch.float64)
    r64 = orc.loss_and_grads(ops64, g64["x0"], H)
    keys = orc.policy_param_keys(ops64)
    # (the tensor-core sweeps' 3xTF32 split + TMEM accumulation measure 2.5e-5 here, the FFMA2 variants 2e-5;
    # the stated H=400 tolerance is 1e-4)
    assert gu.rel_l2(r["grads"], [r64["grads"][k] for k in keys]) < (5e-5 if sweeps == "tc" else 2e-5)


@pytest.mark.parametrize("name,sweeps", [("cartpole_37x2_n7_h12", "ring"), ("cartpole_37x2_n7_h12", "cluster"),
                                         ("cartpole_200x2_n25_h40", "cluster"), ("dcartpole_48x3_n24_h30", "ring"),
                                         ("cartpole_37x2_n7_h12", "tc"), ("cartpole_200x2_n25_h40", "tc"),
                                         ("dcartpole_48x3_n24_h30", "tc"), ("cartpole_37x2_n7_h12", "cw"),
                                         ("cartpole_200x2_n25_h40", "cw")])
def test_generic_cotangents_match_oracle_autograd(name, sweeps):
    """Arbitrary cotangents on states/actions/rewards (value-function tails, CVaR, callbacks)."""
    ops, g = gu.load(name)
    H = int(g["H"])
    r = _run(ops, g["x0"], H, cot="generic", env=SWEEPS[sweeps])
    ops64, g64 = gu.load(name, torch.float64)
    keys = orc.policy_param_keys(ops64)
    d = dict(ops64)
    for k in keys:
        d[k] = d[k].clone().requires_grad_(True)
    x0 = g64["x0"].clone().requires_grad_(True)
    S, A, R = orc.rollout(d, x0, H)
    gS, gA, gR = r["cots"]
    obj = (torch.stack(S) * gS).sum() + (torch.stack(A) * gA).sum() + (torch.stack(R).squeeze(-1) * gR).sum()
    auto = torch.autograd.grad(obj, [d[k] for k in keys] + [x0])
    assert gu.rel_l2(r["grads"], list(auto[:-1])) < 2e-5
    assert gu.rel_l2(r["dx0"], auto[-1]) < 2e-5


def test_ring_depths_agree():
    """The TMA weight ring with 2 or 4 stages (different k-chunk sizes, hence a different grouping of the
    k-split partial sums) gives the same trajectory and gradient to fp32 rounding."""
    ops, g = gu.load("cartpole_200x2_n25_h40")
    a = _run(ops, g["x0"], int(g["H"]), env={"PMB_STAGES": 2, "PMB_STREAM_MODE": 2})
    b = _run(ops, g["x0"], int(g["H"]), env={"PMB_STAGES": 4, "PMB_STREAM_MODE": 2})
    assert (a["S"] - b["S"]).abs().max() < 1e-6 and (a["R"] - b["R"]).abs().max() < 1e-6
    assert gu.rel_l2(a["grads"], b["grads"]) < 2e-6
    for r in (a, b):
        assert gu.rel_l2(r["grads"], gu.policy_grad_list(g, "nomm", ops)) < 1e-5


@pytest.mark.parametrize("fixture", ["cartpole_200x2_n25_h40", "cartpole_200x2_n100_h400", "cartpole_37x2_n7_h12"])
def test_tensor_core_weight_gradient_matches_fp32_kernel(fixture):
    """PMB_WGRAD_UMMA=1 (0 = auto, 2 = off) routes the hidden x hidden weight gradient through the tcgen05 kernel (TF32 hi/lo split,
    3 MMAs, fp32 accumulation in TMEM).  It must agree with the FFMA2 split-K kernel to fp32 rounding and
    meet the same 1e-5 budget against the reference's gradient."""
    ops, g = gu.load(fixture)
    a = _run(ops, g["x0"], int(g["H"]), env={"PMB_WGRAD_UMMA": 2})
    b = _run(ops, g["x0"], int(g["H"]), env={"PMB_WGRAD_UMMA": 1})
    ref = gu.policy_grad_list(g, "nomm", ops)
    ea, eb = gu.rel_l2(a["grads"], ref), gu.rel_l2(b["grads"], ref)
    print(f"{fixture}: FFMA2 vs reference {ea:.2e}, tcgen05 vs reference {eb:.2e}, "
          f"mutual {gu.rel_l2(a['grads'], b['grads']):.2e}")
    assert gu.rel_l2(a["grads"], b["grads"]) < 5e-6      # two fp32 summation orders over H*N rows
    assert ea < 1e-5 and eb < 1e-5


@pytest.mark.parametrize("P", [1, 2, 4, 8])
def test_particles_per_cta_variants(P):
    ops, g = gu.load("cartpole_37x2_n7_h12")   # N=7: ragged last CTA for every P > 1
    r = _run(ops, g["x0"], int(g["H"]), env={"PMB_PARTICLES_PER_CTA": P, "PMB_STREAM_MODE": 2})
    assert (r["S"] - g["nomm_states"]).abs().max() < 2e-6
    assert gu.rel_l2(r["grads"], gu.policy_grad_list(g, "nomm", ops)) < 1e-5


@pytest.mark.parametrize("C", [4, 8])
@pytest.mark.parametrize("PG", [1, 2, 3, 5, 8])
def test_particles_per_cluster_variants(PG, C):
    """Cluster-resident sweeps with every tiling of N=7 particles (ragged last cluster, partly filled tiles) and
    both cluster sizes (different column slices, hence a different grouping of the partial sums)."""
    ops, g = gu.load("cartpole_37x2_n7_h12")
    r = _run(ops, g["x0"], int(g["H"]), env={"PMB_STREAM_MODE": 3, "PMB_CLUSTER_PG": PG, "PMB_CLUSTER_C": C})
    assert (r["S"] - g["nomm_states"]).abs().max() < 2e-6
    assert (r["A"] - g["nomm_actions"]).abs().max() < 2e-5
    assert (r["R"] - g["nomm_rewards"]).abs().max() < 2e-6
    assert gu.rel_l2(r["grads"], gu.policy_grad_list(g, "nomm", ops)) < 1e-5
    assert gu.rel_l2(r["dx0"], g["nomm_dx0"]) < 1e-5


def test_planner_runs_the_bench_workload_on_the_cluster_resident_sweeps(monkeypatch):
    """On the device the planner sizes the clusters from the co-residency query: c2 (100 particles, 2x[200]) lands on
    the cluster-resident sweeps with all 100 particles in flight at once (one wave of 8-CTA clusters)."""
    from prob_mbrl_b200 import _lib
    monkeypatch.delenv("PMB_STREAM_MODE", raising=False)
    ops, g = gu.load("cartpole_200x2_n100_h400")
    o = _ops_cuda(ops)
    prob, keep = _lib.make_problem(o, 100, int(g["H"]))
    info = _lib.describe_plan(prob, _lib.make_tuning())
    assert info["variant"] == 1 and info["cluster_size"] == 8
    assert info["ctas"] <= 148 and info["ctas"] // 8 * info["particles_per_group"] >= 100
    mm = _lib.make_problem(o, 100, int(g["H"]), mm_states=True, z_mm=torch.zeros(500, o.D, device="cuda"))[0]
    info = _lib.describe_plan(mm, _lib.make_tuning())      # moment matching (c3): cluster-resident as well, every cluster
    assert info["variant"] == 1 and info["ctas"] <= 148     # co-resident (the per-step exchange is a grid-wide barrier)
    assert info["ctas"] // 8 * info["particles_per_group"] >= 100
    assert _lib.describe_plan(mm, _lib.make_tuning(stream_mode=2))["variant"] == 0      # streaming sweeps on request
    info = _lib.describe_plan(mm, _lib.make_tuning(stream_mode=4))          # ... or, opt-in, the tensor-core cluster sweeps
    assert info["variant"] == 2 and info["cluster_size"] == 16 and info["ctas"] == 16


@pytest.mark.parametrize("sweeps", ["ring", "cluster", "tc", "cw"])
@pytest.mark.parametrize("D,U", [(3, 2), (4, 3)])
def test_multi_dimensional_actions_match_oracle(D, U, sweeps):
    """The reference's environments all have one action dimension; the kernels are written for U >= 1.  Random nets
    with U = 2, 3 (different hidden widths per layer, ragged particle count) against the fp64 oracle: trajectory,
    loss, policy gradient and dL/dx0, with the generic-cotangent path as well."""
    ops, x0 = gu.synthetic_ops(D=D, U=U, hid=(24, 20), N=9)
    ops64, x064 = gu.synthetic_ops(D=D, U=U, hid=(24, 20), N=9, dtype=torch.float64)
    H = 7
    r = _run(ops, x0, H, env=SWEEPS[sweeps])
    ref = orc.loss_and_grads(ops64, x064, H)
    keys = orc.policy_param_keys(ops64)
    assert (r["S"].double() - torch.stack(ref["states"])).abs().max() < 5e-6
    assert (r["A"].double() - torch.stack(ref["actions"])).abs().max() < 2e-5
    assert abs(float(r["obj"]) - float(ref["loss"])) < 2e-6
    assert gu.rel_l2(r["grads"], [ref["grads"][k] for k in keys]) < 2e-5
    assert gu.rel_l2(r["dx0"], ref["dx0"]) < 2e-5
    rc = _run(ops, x0, H, cot="generic", env=SWEEPS[sweeps])
    d = dict(ops64)
    for k in keys:
        d[k] = d[k].clone().requires_grad_(True)
    x = x064.clone().requires_grad_(True)
    S, A, R = orc.rollout(d, x, H)
    gS, gA, gR = rc["cots"]
    obj = (torch.stack(S) * gS).sum() + (torch.stack(A) * gA).sum() + (torch.stack(R).squeeze(-1) * gR).sum()
    auto = torch.autograd.grad(obj, [d[k] for k in keys] + [x])
    assert gu.rel_l2(rc["grads"], list(auto[:-1])) < 2e-5
    assert gu.rel_l2(rc["dx0"], auto[-1]) < 2e-5


@pytest.mark.parametrize("sweeps", ["ring", "cluster", "tc", "cw"])
@pytest.mark.parametrize("pol_density,dyn_density", [(False, True), (True, False), (False, False)])
def test_nets_without_output_density_match_oracle(pol_density, dyn_density, sweeps):
    """Deterministic policy (plain Linear output, models/core.py:243 applies tanh to it) and / or a dynamics model
    without output_density (models/core.py:185: outs * Sy + my): against the fp64 oracle."""
    kw = dict(D=3, U=2, hid=(24, 20), N=9, pol_density=pol_density, dyn_density=dyn_density)
    ops, x0 = gu.synthetic_ops(**kw)
    ops64, x064 = gu.synthetic_ops(dtype=torch.float64, **kw)
    H = 7
    r = _run(ops, x0, H, env=SWEEPS[sweeps])
    ref = orc.loss_and_grads(ops64, x064, H)
    keys = orc.policy_param_keys(ops64)
    assert (r["S"].double() - torch.stack(ref["states"])).abs().max() < 5e-6
    assert (r["A"].double() - torch.stack(ref["actions"])).abs().max() < 2e-5
    assert abs(float(r["obj"]) - float(ref["loss"])) < 2e-6
    assert gu.rel_l2(r["grads"], [ref["grads"][k] for k in keys]) < 2e-5
    assert gu.rel_l2(r["dx0"], ref["dx0"]) < 2e-5


@pytest.mark.parametrize("other", ["cluster", "cw"])
def test_sweep_variants_agree(other):
    """The streaming and the cluster-resident sweeps give the same trajectory and gradient to fp32 rounding."""
    ops, g = gu.load("cartpole_200x2_n25_h40")
    ref = gu.policy_grad_list(g, "nomm", ops)
    a = _run(ops, g["x0"], int(g["H"]), env=SWEEPS["ring"])
    b = _run(ops, g["x0"], int(g["H"]), env=SWEEPS[other])
    assert (a["S"] - b["S"]).abs().max() < 1e-6 and (a["R"] - b["R"]).abs().max() < 1e-6
    assert gu.rel_l2(a["grads"], b["grads"]) < 5e-6
    assert gu.rel_l2(a["grads"], ref) < 1e-5 and gu.rel_l2(b["grads"], ref) < 1e-5


@pytest.mark.parametrize("hid,N,H", [((512, 512), 80, 12), ((300, 404), 41, 9), ((512, 512), 250, 5)])
def test_wide_nets_on_the_wide_cluster_sweeps_match_streaming_and_oracle(hid, N, H):
    """c5-shaped nets (2x[512]), ragged widths and particle counts, 36 particles per cluster: the planner picks the wide
    cluster-resident sweeps on its own; trajectory against the streaming sweeps and the fp64 oracle, gradients against
    the oracle's autograd for random cotangents.  With ~10^6 hidden units per rollout some ReLU inputs are within fp32
    rounding of zero, where the gate (hence that particle's gradient, by O(1/width)) legitimately depends on the
    summation order: those particles get zero cotangents."""
    from prob_mbrl_b200 import _lib
    kw = dict(D=4, U=1, hid=hid, N=N)
    ops, x0 = gu.synthetic_ops(**kw)
    ops64, x064 = gu.synthetic_ops(dtype=torch.float64, **kw)
    o = _ops_cuda(ops)
    prob, _keep = _lib.make_problem(o, N, H)
    info = _lib.describe_plan(prob, _lib.make_tuning())
    assert info["variant"] == 3 and info["cluster_size"] == 16 and info["threads_per_cta"] == 512
    if N == 250:
        assert info["particles_per_group"] == 36
    a = _run(ops, x0, H, env=SWEEPS["ring"])
    b = _run(ops, x0, H)
    assert (a["S"] - b["S"]).abs().max() < 2e-6 and (a["R"] - b["R"]).abs().max() < 2e-6
    ref = orc.loss_and_grads(ops64, x064, H)
    assert (b["S"].double() - torch.stack(ref["states"])).abs().max() < 5e-6
    assert (b["A"].double() - torch.stack(ref["actions"])).abs().max() < 2e-5
    assert abs(float(b["obj"]) - float(ref["loss"])) < 2e-6
    safe = (gu.min_abs_preactivation(ops64, x064, H, per_particle=True) > 5e-6).double()
    assert safe.sum() > 0.6 * N
    keys = orc.policy_param_keys(ops64)
    for sweeps in (None, SWEEPS["ring"]):
        rc = _run(ops, x0, H, cot="generic", env=sweeps, cot_mask=safe)
        d = dict(ops64)
        for k in keys:
            d[k] = d[k].clone().requires_grad_(True)
        x = x064.clone().requires_grad_(True)
        S, A, R = orc.rollout(d, x, H)
        gS, gA, gR = rc["cots"]
        obj = (torch.stack(S) * gS).sum() + (torch.stack(A) * gA).sum() + (torch.stack(R).squeeze(-1) * gR).sum()
        auto = torch.autograd.grad(obj, [d[k] for k in keys] + [x])
        assert gu.rel_l2(rc["grads"], list(auto[:-1])) < 2e-5
        assert gu.rel_l2(rc["dx0"], auto[-1]) < 2e-5


def test_determinism():
    ops, g = gu.load("cartpole_200x2_n25_h40")
    a = _run(ops, g["x0"], int(g["H"]))
    b = _run(ops, g["x0"], int(g["H"]))
    for x, y in zip(a["grads"], b["grads"]):
        assert torch.equal(x, y)


def test_clip_adam_matches_torch():
    import ctypes as C
    from prob_mbrl_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(0)
    shapes = [(200, 5), (200,), (200, 200), (200,), (2, 200), (2,)]
    ps = [torch.randn(s, device="cuda") for s in shapes]
    gs = [torch.randn(s, device="cuda") * 3 for s in shapes]
    ref_p = [torch.nn.Parameter(p.clone()) for p in ps]
    opt = torch.optim.Adam(ref_p, lr=1e-3)
    m = [torch.zeros_like(p) for p in ps]
    v = [torch.zeros_like(p) for p in ps]
    entries = (_lib.PmbAdamTensor * len(ps))()
    for i in range(len(ps)):
        entries[i].param, entries[i].grad = ps[i].data_ptr(), gs[i].data_ptr()
        entries[i].exp_avg, entries[i].exp_avg_sq, entries[i].n = m[i].data_ptr(), v[i].data_ptr(), ps[i].numel()
    table = torch.frombuffer(bytearray(bytes(entries)), dtype=torch.uint8).clone().cuda()
    scratch = torch.zeros(1024, device="cuda")
    for step in range(1, 4):
        for rp, g in zip(ref_p, gs):
            rp.grad = g.clone()
        norm = torch.nn.utils.clip_grad_norm_(ref_p, 1.0)
        opt.step()
        # a set status word predicates the whole update off (reference: a failed rollout skips the iteration)
        before = [p.clone() for p in ps + m + v]
        flag = torch.ones(1, dtype=torch.int32, device="cuda")
        _lib.check(lib.pmb_clip_adam_step(table.data_ptr(), len(ps), 1.0, 1e-3, 0.9, 0.999, 1e-8, step, None,
                                          scratch.data_ptr(), flag.data_ptr(), _lib.current_stream_ptr()))
        torch.cuda.synchronize()
        assert all(torch.equal(a, b) for a, b in zip(before, ps + m + v))
        flag.zero_()
        _lib.check(lib.pmb_clip_adam_step(table.data_ptr(), len(ps), 1.0, 1e-3, 0.9, 0.999, 1e-8, step, None,
                                          scratch.data_ptr(), flag.data_ptr(), _lib.current_stream_ptr()))
        torch.cuda.synchronize()
        assert abs(float(scratch[0]) - float(norm)) < 1e-3 * float(norm)
        for p, rp in zip(ps, ref_p):
            assert (p - rp.detach()).abs().max() < 2e-6
        # clip scales .grad in place, like clip_grad_norm_: re-arm the gradients for the next step
        for g, rp in zip(gs, ref_p):
            g.copy_(torch.randn_like(g) * 3)


@pytest.mark.parametrize("graph", ["0", "1"])
def test_mc_pilco_iterations_match_reference_golden(graph):
    """The on-device iteration (rollout + BPTT + wgrad + clip + Adam) against the reference's own
    algorithms.mc_pilco on identical noise: final parameters after 6 iterations."""
    import prob_mbrl_b200 as pm
    ops, g = gu.load("mcpilco_cartpole_32x2_n16_h10")
    dyn, pol = gu.modules_from_ops(ops, "cuda")
    opt = torch.optim.Adam(pol.parameters(), float(g["lr"]))
    eng_env = {"PMB_CUDA_GRAPH": graph, "PMB_NO_PBAR": "1"}
    old = {k: os.environ.get(k) for k in eng_env}
    os.environ.update(eng_env)
    try:
        H, N = int(g["H"]), int(g["N"])
        g_r = torch.full((H, N), -1.0 / (H * N), device="cuda")
        eng = pm.FusedIteration(dyn, pol, g["x0"].cuda(), H, opt, g_r, 1.0)
        losses = []
        for _ in range(int(g["iters"])):
            losses.append(float(eng.step(g["x0"].cuda())))
    finally:
        for k, v in old.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
    assert torch.allclose(torch.tensor(losses, dtype=torch.float64), g["losses"].double(), rtol=0, atol=5e-7)
    for i, p in enumerate(pol.parameters()):
        assert (p.detach().cpu() - g["final%d" % i]).abs().max() < 2e-6
    assert int(float(opt.state[next(iter(pol.parameters()))]["step"])) == int(g["iters"])


def test_drop_in_rollout_api_lists_and_autograd():
    """rollout() keeps the reference's return convention and is differentiable end to end."""
    import prob_mbrl_b200 as pm
    ops, g = gu.load("cartpole_200x2_n25_h40")
    dyn, pol = gu.modules_from_ops(ops, "cuda")
    H = int(g["H"])
    x0 = g["x0"].cuda().requires_grad_(True)
    S, A, R = pm.rollout(x0, dyn, pol, H, resample_state_noise=False, resample_action_noise=False)
    assert len(S) == H + 1 and len(A) == H and len(R) == H and R[0].shape == (25, 1)
    loss = -(torch.stack(R).sum(0) / H).mean()
    loss.backward()
    grads = [p.grad.cpu() for p in pol.parameters()]
    assert abs(float(loss) - float(g["nomm_loss"])) < 1e-7
    assert gu.rel_l2(grads, gu.policy_grad_list(g, "nomm", ops)) < 1e-5
    assert gu.rel_l2(x0.grad.cpu(), g["nomm_dx0"]) < 1e-5


def test_unfused_configuration_raises_not_silently_falls_back():
    import prob_mbrl_b200 as pm
    ops, g = gu.load("cartpole_37x2_n7_h12")
    dyn, pol = gu.modules_from_ops(ops, "cuda")
    with pytest.raises(pm.NotEligible):
        pm.rollout(g["x0"].cuda(), dyn, pol, 3, resample_model=True)
    with pytest.raises(pm.NotEligible):
        pm.rollout(g["x0"], dyn, pol, 3, resample_state_noise=False, resample_action_noise=False)  # CPU tensor


# ----------------------------------------------------------------------------------------------
# moment matching (reference utils/rollout.py:20-29,121-145): tolerances of SURVEY App. C.3 for mm
# (the matching amplifies rounding: states 2e-3, loss rtol 1e-5, policy-grad rel-L2 2e-3)
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("sweeps", ["ring", "tc", "cluster"])
@pytest.mark.parametrize("name,tag,groups", [("cartpole_200x2_n25_h40", "mm", None),
                                             ("cartpole_37x2_n7_h12", "mm", None),
                                             ("dcartpole_48x3_n24_h30", "mm", None),
                                             ("dcartpole_48x3_n24_h30", "mmg", 2)])
def test_moment_matching_matches_reference_golden(name, tag, groups, sweeps):
    if sweeps == "cluster" and "x3" in name:
        pytest.skip("three hidden layers / matching groups: streaming sweeps only")
    ops, g = gu.load(name)
    H = int(g["H"])
    mm = dict(mm_states=True, mm_rewards=True, mm_groups=groups, z_mm=g["z_mm"], z_rr=g["z_rr"])
    r = _run(ops, g["x0"], H, mm=mm, env=SWEEPS[sweeps])
    assert r["status"] == 0
    # Budgets: the matching is explosive (|s| reaches ~17 on the double-pole fixture, where
    # the reference's own fp32-vs-fp64 state error is 2e-3 .. 4e-3) and the matching is ill-conditioned when a group has few particles (7 particles in 5
    # dims: the reference's own fp32 gradient is 3.4e-3 from its fp64 twin; measured on the B200 ours is
    # 7.3e-3, and 4e-4 vs the reference's 1.6e-3 on the grouped fixture -- rounding noise amplified by the
    # conditioning, either sign; two builds of this kernel that only differ in fp32 summation order gave
    # 0.7e-2 and 1.2e-2 on the 7-particle fixture), so the bar is stated against the fp64 oracle and
    # relative to the reference's own fp32 error (5x), never tighter than 2e-3.
    ops64, g64 = gu.load(name, torch.float64)
    r64 = orc.loss_and_grads(ops64, g64["x0"], H, mm_states=True, mm_rewards=True, z_mm=g64["z_mm"],
                             z_rr=g64["z_rr"], mm_groups=groups)
    keys = orc.policy_param_keys(ops64)
    g64l = [r64["grads"][k] for k in keys]
    gold = gu.policy_grad_list(g, tag, ops)
    ref_err = gu.rel_l2(gold, g64l)
    S64, R64 = torch.stack(r64["states"]), torch.stack(r64["rewards"]).squeeze(-1)
    assert (r["S"].double() - S64).abs().max() < max(2e-3, 3 * float((g[tag + "_states"].double() - S64).abs().max()))
    assert (r["R"].double() - R64).abs().max() < max(2e-4, 3 * float((g[tag + "_rewards"].double() - R64).abs().max()))
    loss_budget = max(1e-5 * abs(float(r64["loss"])), 3 * abs(float(g[tag + "_loss"]) - float(r64["loss"])))
    assert abs(float(r["obj"]) - float(r64["loss"])) <= loss_budget
    assert abs(float(r["obj"]) - float(g[tag + "_loss"])) <= 2 * loss_budget
    assert gu.rel_l2(r["grads"], g64l) < max(2e-3, 5 * ref_err)
    assert gu.rel_l2(r["grads"], gold) < max(2e-3, 5 * ref_err)
    assert gu.rel_l2(r["dx0"], r64["dx0"]) < max(2e-3, 5 * gu.rel_l2(g[tag + "_dx0"], r64["dx0"]))


@pytest.mark.parametrize("sweeps", ["ring", "tc", "cluster"])
def test_c3_full_size_moment_matching_matches_reference_golden(sweeps):
    """BASELINE.json configs[2]: Cartpole 2x[200], 100 particles, H=400, mm_states + mm_rewards on the whitened
    z_mm table of SURVEY.md section 8d.  On this fixture the reference's own fp32-vs-fp64 error is 1.8e-6 on the
    states and 3e-6 on the gradient; the bar is 1e-4 on the gradient (the H=400 budget of App. C.3) against both
    the reference's fp32 result and the fp64 oracle."""
    ops, g = gu.load("cartpole_200x2_n100_h400_mm")
    H, thin = int(g["H"]), int(g["thin"])
    mm = dict(mm_states=True, mm_rewards=True, mm_groups=None, z_mm=g["z_mm"], z_rr=g["z_rr"])
    r = _run(ops, g["x0"], H, mm=mm, env=SWEEPS[sweeps])
    assert r["status"] == 0
    assert (r["S"][::thin] - g["mm_states"]).abs().max() < 1e-4
    assert (r["R"][::thin] - g["mm_rewards"]).abs().max() < 1e-5
    assert abs(float(r["obj"]) - float(g["mm_loss"])) <= 2e-6 * abs(float(g["mm_loss"]))
    gold = gu.policy_grad_list(g, "mm", ops)
    assert gu.rel_l2(r["grads"], gold) < 1e-4
    assert gu.rel_l2(r["dx0"], g["mm_dx0"]) < 1e-4
    ops64, g64 = gu.load("cartpole_200x2_n100_h400_mm", torch.float64)
    r64 = orc.loss_and_grads(ops64, g64["x0"], H, mm_states=True, mm_rewards=True, z_mm=g64["z_mm"], z_rr=g64["z_rr"])
    keys = orc.policy_param_keys(ops64)
    assert gu.rel_l2(r["grads"], [r64["grads"][k] for k in keys]) < 1e-4


@pytest.mark.parametrize("sweeps", ["ring", "tc", "cluster"])
@pytest.mark.parametrize("which", ["states", "rewards"])
def test_moment_matching_single_flag_matches_oracle(which, sweeps):
    """mm_states and mm_rewards alone, against the fp64 oracle with generic cotangents (25 particles in
    5 dims: a well-conditioned matching)."""
    name = "cartpole_200x2_n25_h40"
    ops, g = gu.load(name)
    H = int(g["H"])
    flags = dict(mm_states=which == "states", mm_rewards=which == "rewards")
    r = _run(ops, g["x0"], H, cot="generic", mm=dict(flags, mm_groups=None, z_mm=g["z_mm"], z_rr=g["z_rr"]),
             env=SWEEPS[sweeps])
    gS, gA, gR = r["cots"]

    def oracle(dtype):
        o_, g_ = gu.load(name, dtype)
        keys = orc.policy_param_keys(o_)
        d = dict(o_)
        for k in keys:
            d[k] = d[k].clone().requires_grad_(True)
        x0 = g_["x0"].clone().requires_grad_(True)
        S, A, R = orc.rollout(d, x0, H, z_mm=g_["z_mm"], z_rr=g_["z_rr"], **flags)
        obj = ((torch.stack(S) * gS.to(dtype)).sum() + (torch.stack(A) * gA.to(dtype)).sum()
               + (torch.stack(R).squeeze(-1) * gR.to(dtype)).sum())
        auto = torch.autograd.grad(obj, [d[k] for k in keys] + [x0])
        return torch.stack(S).detach(), list(auto[:-1]), auto[-1]

    S64, G64, X64 = oracle(torch.float64)
    S32, G32, X32 = oracle(torch.float32)      # what a correct fp32 implementation achieves on this objective
    assert (r["S"].double() - S64).abs().max() < max(1e-3, 3 * float((S32.double() - S64).abs().max()))
    assert gu.rel_l2(r["grads"], G64) < max(2e-3, 3 * gu.rel_l2(G32, G64))
    assert gu.rel_l2(r["dx0"], X64) < max(2e-3, 3 * gu.rel_l2(X32, X64))


@pytest.mark.parametrize("D,U,N", [(8, 2, 64), (7, 2, 33)])
def test_moment_matching_many_state_dims_cluster_equals_streaming(D, U, N):
    """State dimensions beyond the fixtures' (up to the cluster sweeps' limit of 2 D <= 16 raw outputs) and ragged
    particle counts: the cluster-resident sweeps with the in-kernel matching against the streaming sweeps and the fp64
    oracle."""
    H = 5
    kw = dict(D=D, U=U, hid=(24, 20), N=N)
    ops, x0 = gu.synthetic_ops(**kw)
    ops64, x064 = gu.synthetic_ops(dtype=torch.float64, **kw)
    g = torch.Generator().manual_seed(11)
    z = torch.randn(H + N, D, generator=g, dtype=torch.float64)
    zz = z[:N] - z[:N].mean(0, keepdim=True)                      # whitened rows: a bounded matched rollout
    L = torch.linalg.cholesky(zz.T @ zz / (N - 1))
    z[:N] = torch.linalg.solve_triangular(L, zz.T, upper=False).T
    z_rr = torch.randn(H + N, 1, generator=g, dtype=torch.float64)
    mm = dict(mm_states=True, mm_rewards=False, mm_groups=None, z_mm=z.float(), z_rr=z_rr.float())
    a = _run(ops, x0, H, mm=mm, env=SWEEPS["ring"])
    b = _run(ops, x0, H, mm=mm, env=SWEEPS["cluster"])
    assert a["status"] == 0 and b["status"] == 0
    ref = orc.loss_and_grads(ops64, x064, H, mm_states=True, mm_rewards=False, z_mm=z, z_rr=z_rr)
    keys = orc.policy_param_keys(ops64)
    S64 = torch.stack(ref["states"])
    for r in (a, b):
        assert (r["S"].double() - S64).abs().max() < 2e-4
        assert gu.rel_l2(r["grads"], [ref["grads"][k] for k in keys]) < 2e-3
        assert gu.rel_l2(r["dx0"], ref["dx0"]) < 2e-3
    assert (a["S"] - b["S"]).abs().max() < 1e-4


def test_moment_matching_rank_deficient_raises_runtime_error():
    """8 particles per group in 8 state dims: the covariance is singular; the reference raises from
    cholesky() (RuntimeError) at step 0 -- the fused path must report the same way."""
    import prob_mbrl_b200 as pm
    ops, g = gu.load("dcartpole_48x3_n24_h30")
    dyn, pol = gu.modules_from_ops(ops, "cuda")
    x0 = g["x0"][:16].cuda()
    with pytest.raises(RuntimeError):
        pm.rollout(x0, dyn, pol, 10, resample_state_noise=False, resample_action_noise=False, mm_states=True,
                   mm_rewards=True, z_mm=g["z_mm"].cuda(), z_rr=g["z_rr"].cuda(), mm_groups=2)


def test_mc_pilco_mm_iterations_match_reference_golden():
    import prob_mbrl_b200 as pm
    ops, g = gu.load("mcpilco_mm_cartpole_32x2_n16_h10")
    dyn, pol = gu.modules_from_ops(ops, "cuda")
    opt = torch.optim.Adam(pol.parameters(), float(g["lr"]))
    os.environ["PMB_NO_PBAR"] = "1"
    H, N = int(g["H"]), int(g["N"])
    g_r = torch.full((H, N), -1.0 / (H * N), device="cuda")
    mm = dict(mm_states=True, mm_rewards=True, mm_groups=None, z_mm=g["z_mm"].cuda(), z_rr=g["z_rr"].cuda())
    eng = pm.FusedIteration(dyn, pol, g["x0"].cuda(), H, opt, g_r, 1.0, mm)
    losses = [float(eng.step(g["x0"].cuda())) for _ in range(int(g["iters"]))]
    assert int(eng.status.item()) == 0
    assert torch.allclose(torch.tensor(losses, dtype=torch.float64), g["losses"].double(), rtol=0, atol=2e-6)
    for i, p in enumerate(pol.parameters()):
        assert (p.detach().cpu() - g["final%d" % i]).abs().max() < 1e-5
