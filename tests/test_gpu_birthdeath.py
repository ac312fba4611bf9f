"""Fixture C5' (SURVEY.md 8d): death compaction (gpu:reallocate=true), birth into the
function's own next state (newborns first, then the surviving parents by append),
rand48 and a spatial message, checked after EVERY agent function: counts, id
multiset and ORDER, every variable and the seeds array -- bit-exact."""
import os

import numpy as np
import pytest

from common import ROOT, W, build, gen_oracle, OracleSimulation
from ev_diffusion_b200.runtime import Simulation

pytestmark = pytest.mark.gpu
XML = os.path.join(build.MODELS, "BirthDeath", "src", "model", "XMLModelFile.xml")


def _sims():
    p = build.model_path("BirthDeath_exact")
    if not os.path.exists(p):
        p = build.build_model(XML, "BirthDeath_exact", fmad=False)
    so = gen_oracle.oracle_path("BirthDeath")
    if not os.path.exists(so):
        so = gen_oracle.build_oracle(XML, "BirthDeath")
    return Simulation(p), OracleSimulation(so)


def _state(n, seed=45):
    u = W.uniform24(seed, 3 * n)
    return {"id": np.arange(n, dtype=np.int32), "x": u[0::3] * np.float32(32), "y": u[1::3] * np.float32(32),
            "z": u[2::3] * np.float32(32), "age": np.zeros(n, np.int32), "crowd": np.zeros(n, np.int32)}


def _same(gpu, orc, label):
    n = orc.agent_count("Cell", "default")
    assert gpu.agent_count("Cell", "default") == n, label
    a, b = gpu.agents("Cell"), orc.agents("Cell")
    for k in a:
        assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), (label, k)
    assert np.array_equal(gpu.seeds(), orc.seeds()), label


@pytest.mark.parametrize("n,iters", [(1, 30), (1000, 40), (5000, 40), (1 << 16, 12)])
def test_birth_death_every_function(n, iters):
    gpu, orc = _sims()
    st = _state(n)
    for s in (gpu, orc):
        s.set_agents("Cell", "default", st)
    for it in range(iters):
        gpu.begin_iteration()
        orc.begin_iteration()
        for f in ("output_location", "live"):
            gpu.run_function(f)
            orc.run_function(f)
            _same(gpu, orc, (it, f))
    gpu.close()
    orc.close()


@pytest.mark.parametrize("death,birth", [(1.0, 0.0), (0.0, 1.0), (0.5, 0.5)])
def test_extreme_turnover(death, birth):
    """Everyone dies (the list empties and later functions early-out, sim:1417), and
    everyone gives birth until the buffer-size guard fires (sim:1816)."""
    gpu, orc = _sims()
    st = _state(3000)
    for s in (gpu, orc):
        s.set_agents("Cell", "default", st)
        s.set_constant("DEATH_P", death)
        s.set_constant("BIRTH_P", birth)
    for it in range(6):
        gpu.step(1)
        orc.step(1)
        _same(gpu, orc, it)
    gpu.close()
    orc.close()


def test_birth_overflow_is_reported():
    from ev_diffusion_b200.runtime import FgpuError
    gpu, orc = _sims()
    n = gpu.model.agents[0].buffer_size // 2 + 10
    gpu.set_agents("Cell", "default", _state(n))
    gpu.set_constant("DEATH_P", 0.0)
    gpu.set_constant("BIRTH_P", 1.0)
    with pytest.raises(FgpuError, match="will be exceeded"):
        gpu.step(1)
    gpu.close()
    orc.close()
