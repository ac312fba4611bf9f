"""optional_message outputs (simulation.xslt:1660-1750, FLAMEGPU_kernals.xslt:382-441) on
the authored OptionalSignal model: a spatially partitioned and a non-partitioned optional message.  After
every agent function the message counts, the message lists (cell order / agent order), the partition
boundary matrix, the agent state and the rand48 seeds must equal the host oracle's, which restates the
reference's sparse-list + exclusive-scan + scatter sequence."""
import os

import numpy as np
import pytest

from common import build, gen_oracle, OracleSimulation
from ev_diffusion_b200.runtime import FgpuError, Simulation

pytestmark = pytest.mark.gpu
XML = os.path.join(build.MODELS, "OptionalSignal", "src", "model", "XMLModelFile.xml")


def _pair(n, seed=3):
    p = build.model_path("OptionalSignal_exact")
    if not os.path.exists(p):
        p = build.build_model(XML, "OptionalSignal_exact", fmad=False)
    so = gen_oracle.oracle_path("OptionalSignal")
    if not os.path.exists(so):
        so = gen_oracle.build_oracle(XML, "OptionalSignal")
    gpu, orc = Simulation(p), OracleSimulation(so)
    rng = np.random.default_rng(seed)
    cols = {"id": np.arange(n, dtype=np.int32)}
    for k in "xyz":
        cols[k] = (rng.random(n) * 8).astype(np.float32)
    for k in ("level", "heard", "loudest", "beacons"):
        cols[k] = np.zeros(n, np.int32)
    for s in (gpu, orc):
        s.set_agents("Sensor", "default", cols)
    return gpu, orc


def _same_messages(gpu, orc, name):
    assert gpu.message_count(name) == orc.message_count(name)
    gm, om = gpu.messages(name), orc.messages(name)
    for k in gm:
        assert np.array_equal(gm[k].view(np.uint32), om[k].view(np.uint32)), (name, k)


@pytest.mark.parametrize("n,emit_p", [(3000, 0.3), (1, 1.0), (2500, 0.0), (1024, 1.0), (5000, 0.002)])
def test_every_function_matches_the_oracle(n, emit_p):
    gpu, orc = _pair(n)
    for s in (gpu, orc):
        s.set_constant("EMIT_P", np.float32(emit_p))
    for it in range(4):
        for s in (gpu, orc):
            s.begin_iteration()
        for f in ("emit", "listen", "announce", "tally"):
            gpu.run_function(f)
            orc.run_function(f)
            if f == "emit":
                _same_messages(gpu, orc, "signal")
                if gpu.message_count("signal"):
                    gs, gc = gpu.pbm("signal")
                    os_, oc = orc.pbm("signal")
                    assert np.array_equal(gc, oc) and np.array_equal(gs[oc > 0], os_[oc > 0])
            if f == "announce":
                _same_messages(gpu, orc, "beacon")
            a, b = gpu.agents("Sensor"), orc.agents("Sensor")
            for k in a:
                assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), (it, f, k)
            assert np.array_equal(gpu.seeds(), orc.seeds())
    if emit_p == 0.3:
        assert 0 < gpu.message_count("signal") < n and 0 < gpu.message_count("beacon") < n
    if emit_p == 0.0:
        assert gpu.message_count("signal") == 0 and gpu.message_count("beacon") == 0
    gpu.close()
    orc.close()


def test_whole_iterations_and_message_order():
    gpu, orc = _pair(4000, seed=11)
    gpu.step(12)
    orc.step(12)
    a, b = gpu.agents("Sensor"), orc.agents("Sensor")
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    # the non-partitioned optional list is the emitters in agent order (stable compaction)
    ids = gpu.messages("beacon")["id"]
    assert np.all(np.diff(ids) > 0) and np.array_equal(ids, a["id"][a["loudest"] >= 0])
    gpu.close()
    orc.close()
