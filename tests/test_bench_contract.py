"""The CPU arm of bench.py (`--impl reference`: the host-C oracle on all host cores) prints one JSON line with the
keys the driver reads; under torchrun only rank 0 prints.  The GPU arm needs a B200 and is exercised by the
driver itself."""
import json
import os
import subprocess
import sys

from common import ROOT

REQUIRED = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"}


def _cuda_device_present():
    import conftest
    return conftest._has_gpu()


def _run(extra_env=None, gpus=1):
    env = dict(os.environ, **(extra_env or {}))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", str(gpus), "--steps", "1",
                          "--warmup", "1", "--workload", "c2_tiny"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                         timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.strip()]


def test_reference_arm_prints_the_contract_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert REQUIRED <= set(d), REQUIRED - set(d)
    assert d["impl"] == "reference" and d["metric"] == "agent-steps/sec" and d["unit"] == "agent-steps/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert d["warmup"] >= 3                                  # the timing rules ask for at least three warm-up steps
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_under_torchrun_only_rank0_prints():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, gpus=2) == []
    lines = _run({"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"}, gpus=2)
    assert len(lines) == 1 and json.loads(lines[0])["n_gpus"] == 2


def test_reference_cuda_leg_never_raises():
    """bench.py's reference_cuda leg (the reference's own sm_100a build, run as a separate process) reports a reason
    instead of a number when the binary is missing or cannot run -- here: no binary for the tiny workload, and no CUDA
    device in this container for the real one -- and never raises into the bench."""
    import bench
    from ev_diffusion_b200 import workloads as W
    missing = bench.reference_cuda(W.BROWNIAN["c2_tiny"])
    assert set(missing) == {"unavailable"} and "not built" in missing["unavailable"]
    exe = os.path.join(ROOT, "oracle", "_ref", "Brownian", "ref_gpu_Brownian")
    if os.path.exists(exe) and not _cuda_device_present():
        failed = bench.reference_cuda(W.BROWNIAN["c2"], iterations=1, timeout_s=120)
        assert set(failed) == {"unavailable"} and "CUDA" in failed["unavailable"]


def test_reference_arm_of_the_shipped_size_configs():
    """`--workload c1 / c5` (the examples' own 0.xml from tests/golden/): the CPU arm prints the same contract line, on one
    thread (an iteration of a thousand agents is microseconds of work), with the population the fixture holds."""
    import pytest
    from oracle import gen_oracle
    for wl, agents in (("c1", 1024), ("c5", 598)):
        name = {"c1": "CirclesPartitioning_float", "c5": "Keratinocyte"}[wl]
        if not os.path.exists(gen_oracle.oracle_path(name)):
            pytest.skip("oracle library of %s is not prebuilt" % name)
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--workload", wl],
                             stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, cwd=ROOT)
        assert out.returncode == 0, out.stderr[-2000:]
        d = json.loads([l for l in out.stdout.splitlines() if l.strip()][-1])
        assert REQUIRED <= set(d), REQUIRED - set(d)
        assert d["impl"] == "reference" and d["config"]["agents"] == agents and d["cpu_baseline"]["cores"] == 1 and d["value"] > 0
