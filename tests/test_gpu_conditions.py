"""Function conditions and moves between state lists on the authored TwoStage model (a13/a14 of SURVEY.md 8a;
simulation.xslt:1439-1528, 1590-1601, 1936-1972): after EVERY agent function both state lists -- counts, order,
every variable -- and the rand48 seeds must equal the oracle's.  The oracle itself is checked against an
independent numpy restatement of the same list discipline in tests/test_oracle_independent.py."""
import os

import numpy as np
import pytest

from common import build, gen_oracle, OracleSimulation
from ev_diffusion_b200.runtime import Simulation

pytestmark = pytest.mark.gpu
XML = os.path.join(build.MODELS, "TwoStage", "src", "model", "XMLModelFile.xml")


def _pair(n):
    p = build.model_path("TwoStage_exact")
    if not os.path.exists(p):
        p = build.build_model(XML, "TwoStage_exact", fmad=False)
    so = gen_oracle.oracle_path("TwoStage")
    if not os.path.exists(so):
        so = gen_oracle.build_oracle(XML, "TwoStage")
    gpu, orc = Simulation(p), OracleSimulation(so)
    cols = {"id": np.arange(n, dtype=np.int32), "energy": (np.arange(n, dtype=np.int32) * 7) % 5,
            "fired": np.zeros(n, np.int32), "heat": np.zeros(n, np.float32)}
    for s in (gpu, orc):
        s.set_agents("Unit", "idle", cols)
    return gpu, orc


def _same(gpu, orc, label):
    for state in ("idle", "active"):
        n = orc.agent_count("Unit", state)
        assert gpu.agent_count("Unit", state) == n, (label, state)
        a, b = gpu.agents("Unit", state), orc.agents("Unit", state)
        for k in a:
            assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), (label, state, k)
    assert np.array_equal(gpu.seeds(), orc.seeds()), label


@pytest.mark.parametrize("n", [1, 33, 1500, 4096])
def test_every_function_both_lists(n):
    gpu, orc = _pair(n)
    for it in range(10):
        gpu.begin_iteration()
        orc.begin_iteration()
        for f in ("charge", "fire", "burn", "cool"):
            gpu.run_function(f)
            orc.run_function(f)
            _same(gpu, orc, (it, f))
    if n >= 33:
        assert gpu.agent_count("Unit", "active") + gpu.agent_count("Unit", "idle") == n
        assert gpu.agents("Unit", "idle")["fired"].max() >= 1
    gpu.close()
    orc.close()


def test_whole_iterations_and_append_overflow():
    gpu, orc = _pair(3000)
    gpu.step(25)
    orc.step(25)
    _same(gpu, orc, "25 iterations")
    gpu.close()
    orc.close()
