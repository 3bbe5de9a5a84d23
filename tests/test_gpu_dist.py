"""Multi-GPU tests of the sharded path (one process per GPU, NCCL): run scripts/dist_check.py (particle-sharded
fused iteration == single-GPU iteration on the same global batch, incl. PEGASUS resamples) and a short bench.py
run under torch.distributed.run for 2 and 4 ranks.  Skipped when the box has fewer GPUs."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(nranks, script_args, timeout, **extra_env):
    port = 29600 + (os.getpid() + 7 * nranks + 13 * len(extra_env)) % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks),
           "--master-addr", "127.0.0.1", "--master-port", str(port)] + script_args
    env = dict(os.environ, PMB_NO_PBAR="1", **extra_env)
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=env)


@pytest.mark.parametrize("nranks,sync", [(2, "peer"), (2, "nccl"), (4, "peer")])
def test_sharded_iteration_equals_single_gpu(nranks, sync):
    """The gradient exchange over NVLink peer memory (default: `pmb_peer_allreduce`, inside the iteration's CUDA graph)
    and the NCCL all-reduce (PMB_GRAD_SYNC=nccl) both reproduce the single-GPU parameters; with the peer exchange the
    parameters are bitwise identical on every rank."""
    if torch.cuda.device_count() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    out = _torchrun(nranks, [os.path.join(ROOT, "scripts", "dist_check.py")], 420, PMB_GRAD_SYNC=sync)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-3000:]
    assert "DIST PASS" in out.stdout, out.stdout[-1500:]
    assert ("sync=%s" % sync) in out.stdout


@pytest.mark.parametrize("nranks", [2, 4])
def test_sharded_moment_matching_equals_single_gpu(nranks):
    """moment matching of states and rewards over the particles of ALL ranks -- per-step records and arrivals
    over NVLink peer memory inside the cluster-resident sweeps, the rewards through a peer all-gather -- reproduces the
    single-GPU mc_pilco on the same global batch (parameters after 4 iterations incl. a PEGASUS resample)."""
    if torch.cuda.device_count() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    out = _torchrun(nranks, [os.path.join(ROOT, "scripts", "dist_check.py")], 420, DIST_MM="1")
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-3000:]
    assert "DIST PASS mm=1" in out.stdout, out.stdout[-1500:]


@pytest.mark.parametrize("nranks", [2, 4])
def test_bench_runs_sharded_with_a_rank_agreed_iteration_count(nranks):
    """r1's SCALE run died at N=4: a wall-clock-bounded warm-up issued a rank-dependent number of all-reduces.
    Three back-to-back short runs must all return one JSON line with rc 0."""
    if torch.cuda.device_count() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    for _ in range(3):
        out = _torchrun(nranks, [os.path.join(ROOT, "bench.py"), "--gpus", str(nranks), "--steps", "10", "--warmup", "3",
                                 "--quick"], 420)
        assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-3000:]
        line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
        assert line["n_gpus"] == nranks and line["value"] > 0
