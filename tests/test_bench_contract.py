"""bench.py on a GPU-less box: the reference arm (the unmodified reference's algorithms.mc_pilco on the host cores
when the reference package is importable, else the CPU oracle port) prints one JSON line with the contract's keys,
and the algorithmic-work model matches SURVEY.md section 8d."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_flop_model_matches_the_survey():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.flop_model("c2")["total"] == 421200 and bench.flop_model("c1")["total"] == 421200
    assert bench.flop_model("c4")["total"] == 3264000
    assert bench.flop_model("c5")["total"] == 2675712
    fm = bench.flop_model("c2")
    assert fm["fwd"] == fm["bwd_data"] == 169200 and fm["wgrad"] == 82800


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, PMB_NO_PBAR="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c1",
                          "--steps", "1", "--warmup", "3"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "rollout-steps/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] >= 1 and line["warmup"] == 3
    assert line["config"]["workload"].startswith("c1:") and "model" not in line["config"]
    cb = line["cpu_baseline"]
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_shim
    if ref_shim.available():
        # the reference itself, at the same particle count, threads = 1 and = nproc reported separately
        assert cb["kind"] == "reference" and "unmodified reference algorithms.mc_pilco" in cb["sample"]
        assert cb["threads_1"] > 0 and cb["threads_all"] > 0 and cb["value"] == max(cb["threads_1"], cb["threads_all"])
        assert "N=25" in cb["sample"] and line["config"]["timed"] == cb["sample"]
        assert line["cpu_port"]["kind"] == "port"
    else:
        assert cb["kind"] == "port"
    assert cb["cores"] >= 1 and cb["value"] == line["value"] and "H=40" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "rollout-steps/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0
