"""The C-ABI shared library: loads on a GPU-less box, exports every symbol include/pmb_b200.h declares,
its structs have the layout the ctypes binding assumes, and descriptor validation works without a device.
No compute calls here (no GPU in the build container)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pmb_b200.h")


@pytest.fixture(scope="module")
def lib():
    from prob_mbrl_b200 import build
    build.build()
    from prob_mbrl_b200 import _lib
    return _lib.load()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pmb_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    names = declared_functions()
    assert "pmb_rollout_forward" in names and "pmb_rollout_backward" in names and "pmb_clip_adam_step" in names
    for n in names:
        assert hasattr(lib, n), "libpmb_b200.so does not export %s" % n
    from prob_mbrl_b200 import _lib
    assert set(_lib.EXPORTS) == set(names)
    assert lib.pmb_abi_version() == _lib.ABI_VERSION


def test_ctypes_struct_layout_matches_the_c_header():
    from prob_mbrl_b200 import _lib
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "pmb_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu\n", sizeof(pmb_net), sizeof(pmb_problem), sizeof(pmb_tuning), sizeof(pmb_adam_tensor));
  printf("%zu %zu %zu %zu %zu\n", offsetof(pmb_net, W), offsetof(pmb_net, keep), offsetof(pmb_net, z),
         offsetof(pmb_net, z_step_stride), offsetof(pmb_net, max_log_std));
  printf("%zu %zu %zu %zu %zu %zu %zu\n", offsetof(pmb_problem, pol), offsetof(pmb_problem, dyn),
         offsetof(pmb_problem, act_scale), offsetof(pmb_problem, rew_rows), offsetof(pmb_problem, z_mm),
         offsetof(pmb_problem, n_global), offsetof(pmb_problem, masks_binary));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "layout.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "layout")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        out = subprocess.check_output([exe]).decode().split("\n")
    sizes = [int(x) for x in out[0].split()]
    assert sizes == [C.sizeof(_lib.PmbNet), C.sizeof(_lib.PmbProblem), C.sizeof(_lib.PmbTuning),
                     C.sizeof(_lib.PmbAdamTensor)]
    net = [int(x) for x in out[1].split()]
    assert net == [getattr(_lib.PmbNet, f).offset for f in ("W", "keep", "z", "z_step_stride", "max_log_std")]
    prob = [int(x) for x in out[2].split()]
    assert prob == [getattr(_lib.PmbProblem, f).offset for f in ("pol", "dyn", "act_scale", "rew_rows", "z_mm",
                                                                 "n_global", "masks_binary")]


def _fake_problem(N=100, H=400, D=5, U=1, hid=(200, 200), mm=False, groups=0, binary=False):
    """Descriptor with non-NULL fake pointers: enough for the host-side planner (never dereferenced)."""
    from prob_mbrl_b200 import _lib
    p = _lib.PmbProblem()
    p.N, p.H, p.D, p.U = N, H, D, U
    for net, nin, nout in ((p.pol, D, 2 * U), (p.dyn, D + U, 2 * D)):
        dims = [nin] + list(hid) + [nout]
        net.n_linear = len(dims) - 1
        for i, d in enumerate(dims):
            net.dims[i] = d
        for i in range(net.n_linear):
            net.W[i] = 0x1000
            net.b[i] = 0x1000
        for i in range(len(hid)):
            net.mask[i] = 0x1000
            net.keep[i] = 0.9
        net.has_density, net.z, net.max_log_std = 1, 0x1000, 1.6
    for f in ("act_scale", "act_bias", "mx", "iSx", "my", "Sy", "rew_C", "rew_c0", "rew_Q", "rew_R"):
        setattr(p, f, 0x1000)
    p.rew_rows, p.rew_scale, p.n_global = 2, 1.0, N
    p.masks_binary = int(binary)
    if mm:
        p.mm_states = p.mm_rewards = 1
        p.mm_groups = groups
        p.z_mm = p.z_rr = 0x1000
    return p


def test_planner_accepts_the_baseline_configs(lib):
    from prob_mbrl_b200 import _lib
    tune = _lib.make_tuning()
    cases = {"c1": dict(N=25, H=40), "c2": dict(N=100, H=400), "c3": dict(N=100, H=400, mm=True),
             "c4": dict(N=125, H=600, D=8, hid=(400, 400, 400)), "c5": dict(N=250, H=1000, hid=(512, 512))}
    for name, kw in cases.items():
        p = _fake_problem(**kw)
        assert lib.pmb_check_problem(C.byref(p), C.byref(tune)) == 0, (name, lib.pmb_last_error())
        nbytes = lib.pmb_workspace_bytes(C.byref(p), C.byref(tune))
        assert nbytes > 0
        # activations kept for the reverse sweep dominate: H*N*(sum of hidden widths) floats, x3 (pol act, pol delta, dyn act)
        hid = kw.get("hid", (200, 200))
        lower = 4 * kw["H"] * kw["N"] * sum(hid) * 3
        assert lower <= nbytes < 3 * lower + (96 << 20), (name, nbytes, lower)   # + split-K partials of the weight gradient
    assert lib.pmb_policy_param_count(C.byref(_fake_problem())) == 200 * 5 + 200 + 200 * 200 + 200 + 2 * 200 + 2


def test_planner_picks_the_sweep_variant(lib, monkeypatch):
    """Two hidden layers of <= 256 units without state moment matching: FFMA2 cluster-resident sweeps (weights in the
    shared memory of an 8-CTA cluster, two 4-slot particle tiles per CTA); everything else: streaming sweeps.
    stream_mode 2 forces the streaming sweeps, 3 requires cluster-resident, 4 requires the tensor-core cluster
    sweeps (16-CTA cluster per 128-particle tile; opt-in: measured slower than the FFMA2 variants)."""
    from prob_mbrl_b200 import _lib
    from prob_mbrl_b200.operands import NotEligible
    monkeypatch.delenv("PMB_STREAM_MODE", raising=False)
    auto = _lib.make_tuning()
    c2 = _lib.describe_plan(_fake_problem(), auto)
    assert c2["variant"] == 1 and c2["cluster_size"] == 8 and c2["threads_per_cta"] == 256
    assert 1 <= c2["particles_per_group"] <= 8 and c2["ctas"] % 8 == 0
    assert c2["ctas"] // 8 * c2["particles_per_group"] >= 100
    assert c2["smem_fwd_bytes"] <= 232448 and c2["smem_bwd_bytes"] <= 232448
    assert c2["launches_fwd"] == 2 and c2["launches_bwd"] == 1 + 1 + 3 + 1
    tc = _lib.make_tuning(stream_mode=4)
    for kw, tiles, tp in ((dict(N=100, H=400, mm=True), 1, 100), (dict(N=125, H=600, D=8, hid=(400, 400, 400)), 1, 125),
                          (dict(N=250, H=1000, hid=(512, 512)), 2, 125), (dict(N=1000, H=40), 8, 125)):
        # auto: 2x[200] nets run on the 8-CTA cluster-resident sweeps, with moment matching as well (one group of
        # <= 128 particles spread over the co-resident clusters); wider / deeper nets stream
        assert _lib.describe_plan(_fake_problem(**kw), auto)["variant"] == (1 if "hid" not in kw else 0), kw
        info = _lib.describe_plan(_fake_problem(**kw), tc)
        assert info["variant"] == 2 and info["cluster_size"] == 16 and info["ctas"] == 16 * tiles, kw
        assert info["particles_per_group"] == tp and info["threads_per_cta"] == 384
        assert info["smem_fwd_bytes"] <= 232448 - 1024 and info["smem_bwd_bytes"] <= 232448 - 1024
        forced = _lib.describe_plan(_fake_problem(**kw), _lib.make_tuning(stream_mode=2))
        assert forced["variant"] == 0 and forced["cluster_size"] == 1, kw
    tc2 = _lib.describe_plan(_fake_problem(), _lib.make_tuning(stream_mode=4))       # c2 on the tensor-core sweeps
    assert tc2["variant"] == 2 and tc2["launches_fwd"] == 3 and tc2["launches_bwd"] == 1 + 1 + 3 + 1
    with pytest.raises(NotEligible):
        _lib.describe_plan(_fake_problem(N=300, H=40, mm=True), _lib.make_tuning(stream_mode=4))   # group > one tile
    ring = _lib.describe_plan(_fake_problem(), _lib.make_tuning(stream_mode=2))
    assert ring["variant"] == 0 and ring["launches_bwd"] == 1 + 3 + 1
    with pytest.raises(NotEligible):
        _lib.describe_plan(_fake_problem(hid=(400, 400, 400)), _lib.make_tuning(stream_mode=3))
    # ragged particle counts: the last cluster is partly filled
    small = _lib.describe_plan(_fake_problem(N=7, H=12, hid=(37, 37)), auto)
    assert small["variant"] == 1 and small["ctas"] // 8 * small["particles_per_group"] >= 7
    # two hidden layers wider than 256 (c5) with masks declared binary: the wide cluster-resident sweeps (16-CTA
    # cluster, <= 36 particles per cluster; 7 clusters are co-resident on a B200); narrower nets stay on the 8-CTA
    # sweeps, three hidden layers / moment matching / undeclared masks on the streaming sweeps
    c5 = _lib.describe_plan(_fake_problem(N=250, H=1000, hid=(512, 512), binary=True), auto)
    assert c5["variant"] == 3 and c5["cluster_size"] == 16 and c5["threads_per_cta"] == 512
    assert c5["particles_per_group"] == 36 and c5["ctas"] == 7 * 16
    assert c5["smem_fwd_bytes"] <= 232448 - 1024 and c5["smem_bwd_bytes"] <= 232448 - 1024
    assert c5["launches_fwd"] == 2 and c5["launches_bwd"] == 1 + 1 + 3 + 1
    assert _lib.describe_plan(_fake_problem(binary=True), auto)["variant"] == 1
    assert _lib.describe_plan(_fake_problem(N=125, H=600, D=8, hid=(400, 400, 400), binary=True), auto)["variant"] == 0
    assert _lib.describe_plan(_fake_problem(N=100, H=100, hid=(300, 300), binary=True, mm=True), auto)["variant"] == 0
    # moment matching on the cluster-resident sweeps: every cluster co-resident (<= 15 x 8 CTAs), one matching group
    c3 = _lib.describe_plan(_fake_problem(N=100, H=400, mm=True), auto)
    assert c3["variant"] == 1 and c3["ctas"] <= 15 * 8 and c3["ctas"] // 8 * c3["particles_per_group"] >= 100
    assert c3["smem_fwd_bytes"] <= 232448 - 1024 and c3["smem_bwd_bytes"] <= 232448 - 1024
    assert _lib.describe_plan(_fake_problem(N=100, H=400, mm=True, groups=2), auto)["variant"] == 0
    assert _lib.describe_plan(_fake_problem(N=200, H=40, mm=True), auto)["variant"] == 0
    wide = _lib.make_tuning(stream_mode=5)
    assert _lib.describe_plan(_fake_problem(binary=True), wide)["variant"] == 3          # c2 on request
    for kw in (dict(hid=(512, 512)), dict(hid=(400, 400, 400), binary=True), dict(hid=(600, 600), binary=True)):
        with pytest.raises(NotEligible):
            _lib.describe_plan(_fake_problem(**kw), wide)


def test_cluster_tunables_from_the_environment(lib, monkeypatch):
    """PMB_STREAM_MODE=3 + PMB_CLUSTER_PG / PMB_CLUSTER_C / PMB_CLUSTER_PINGPONG reach the planner through
    pmb_tuning.reserved[1]; bad values are rejected by the library, not silently clamped."""
    from prob_mbrl_b200 import _lib
    monkeypatch.setenv("PMB_STREAM_MODE", "3")
    monkeypatch.setenv("PMB_CLUSTER_PG", "5")
    monkeypatch.setenv("PMB_CLUSTER_C", "8")
    monkeypatch.setenv("PMB_CLUSTER_PINGPONG", "0")
    t = _lib.make_tuning()
    assert t.stream_mode == 3 and t.reserved[1] == (5 | (8 << 4) | (1 << 8))
    info = _lib.describe_plan(_fake_problem(), t)
    assert info["variant"] == 1 and info["particles_per_group"] == 5 and info["ctas"] == 20 * 8
    monkeypatch.setenv("PMB_CLUSTER_C", "4")         # 200 columns / 4 CTAs = 52 per CTA > 32: outside the kernels
    from prob_mbrl_b200.operands import NotEligible
    with pytest.raises(NotEligible):
        _lib.describe_plan(_fake_problem(), _lib.make_tuning())
    monkeypatch.setenv("PMB_CLUSTER_C", "8")
    monkeypatch.setenv("PMB_CLUSTER_PG", "9")
    with pytest.raises(RuntimeError):
        _lib.describe_plan(_fake_problem(), _lib.make_tuning())
    # narrow nets fit a 4-CTA cluster
    monkeypatch.setenv("PMB_CLUSTER_PG", "2")
    monkeypatch.setenv("PMB_CLUSTER_C", "4")
    info = _lib.describe_plan(_fake_problem(N=7, H=12, hid=(37, 37)), _lib.make_tuning())
    assert info["cluster_size"] == 4 and info["ctas"] == 4 * 4


def test_planner_rejects_what_the_kernels_cannot_run(lib):
    from prob_mbrl_b200 import _lib
    from prob_mbrl_b200.operands import NotEligible
    tune = _lib.make_tuning()
    too_wide = _fake_problem(hid=(2048, 2048))
    with pytest.raises(NotEligible):
        _lib.check_problem(too_wide, tune)                      # PMB_E_UNSUPPORTED -> NotEligible
    bad = _fake_problem()
    bad.mx = None
    with pytest.raises(RuntimeError):
        _lib.check_problem(bad, tune)                           # PMB_E_INVALID -> RuntimeError
    ragged_groups = _fake_problem(N=100, mm=True, groups=3)
    with pytest.raises(RuntimeError):
        _lib.check_problem(ragged_groups, tune)
    huge_mm = _fake_problem(N=5000, mm=True)                    # grid barrier needs one co-resident grid
    with pytest.raises(NotEligible):
        _lib.check_problem(huge_mm, tune)


def test_product_path_fails_loudly_without_cuda():
    """CPU tensors / missing extension never fall back silently (fused is the default backend)."""
    import prob_mbrl_b200 as pm
    import golden_util as gu
    ops, g = gu.load("cartpole_37x2_n7_h12")
    dyn, pol = gu.modules_from_ops(ops)
    os.environ.pop("PROB_MBRL_BACKEND", None)
    with pytest.raises(pm.NotEligible):
        pm.rollout(g["x0"], dyn, pol, 3, resample_state_noise=False, resample_action_noise=False)
    from prob_mbrl_b200 import _lib
    saved, _lib._lib = _lib._lib, None
    real = _lib.LIB_PATH
    try:
        _lib.LIB_PATH = real + ".missing"
        with pytest.raises(_lib.LibraryMissing):
            _lib.load()
    finally:
        _lib.LIB_PATH, _lib._lib = real, saved


def test_sharded_moment_matching_descriptor(lib, monkeypatch):
    """a descriptor with mm_world > 1 (this rank's equal shard of N * world particles that are matched
    together) plans the cluster-resident sweeps, reports the sizes of its three exchange areas, and is rejected when the
    shard arithmetic is inconsistent or the problem is outside those sweeps (host-side planner only, no device work)."""
    from prob_mbrl_b200 import _lib
    from prob_mbrl_b200.operands import NotEligible
    monkeypatch.delenv("PMB_STREAM_MODE", raising=False)
    tune = _lib.make_tuning()
    p = _fake_problem(N=100, H=400, mm=True)
    p.mm_world, p.mm_rank, p.n_global = 2, 1, 200
    info = _lib.describe_plan(p, tune)
    assert info["variant"] == 1 and info["ctas"] <= 15 * 8
    sizes = (C.c_size_t * 3)()
    assert lib.pmb_mm_exchange_bytes(C.byref(p), C.byref(tune), sizes) == 0, lib.pmb_last_error()
    nq = 5 + 15                                              # D sums + D (D + 1) / 2 products
    tiles = 2 * (info["ctas"] // 8) * 2                      # two particle tiles per cluster, two ranks
    assert sizes[0] >= 2 * 2 * tiles * nq * 16               # [2 sweeps][2 parities][tiles][nq] 16-byte tagged entries
    assert sizes[1] == lib.pmb_peer_buffer_bytes(400 * 100, 2) and sizes[2] >= 40
    # the same shard on one rank of four
    p.mm_world, p.mm_rank, p.n_global = 4, 3, 400
    assert lib.pmb_check_problem(C.byref(p), C.byref(tune)) == 0, lib.pmb_last_error()
    # inconsistent shard arithmetic / rank out of range
    p.n_global = 399
    assert lib.pmb_check_problem(C.byref(p), C.byref(tune)) == -1      # PMB_E_INVALID
    p.n_global, p.mm_rank = 400, 4
    assert lib.pmb_check_problem(C.byref(p), C.byref(tune)) == -1      # PMB_E_INVALID
    # matching groups and nets outside the cluster-resident sweeps are not sharded
    q = _fake_problem(N=100, H=40, mm=True, groups=2)
    q.mm_world, q.mm_rank, q.n_global = 2, 0, 200
    with pytest.raises(NotEligible):
        _lib.describe_plan(q, tune)
    q = _fake_problem(N=100, H=40, mm=True, hid=(400, 400, 400))
    q.mm_world, q.mm_rank, q.n_global = 2, 0, 200
    with pytest.raises(NotEligible):
        _lib.describe_plan(q, tune)


def test_peer_exchange_sizes(lib):
    """pmb_peer_buffer_bytes: [2 parities][world][n] floats + [world] flags; 0 for nonsense arguments."""
    assert lib.pmb_peer_buffer_bytes(41803, 8) >= 2 * 8 * 41803 * 4 + 8 * 8
    assert lib.pmb_peer_buffer_bytes(41803, 1) >= 2 * 41803 * 4 + 8
    assert lib.pmb_peer_buffer_bytes(0, 2) == 0 and lib.pmb_peer_buffer_bytes(10, 17) == 0
