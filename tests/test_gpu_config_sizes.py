"""GPU parity at BASELINE-config size: one iteration of the bit-exact build against one oracle iteration on
the same seeded input, for the configs whose oracle iteration takes seconds on the box's host cores.  (C2-16M:
tests/test_gpu_brownian.py::test_north_star_16m_one_iteration_exact.)  The plugins and oracle libraries are built in the
build container (the scripts are the reference's own, compiled where they lie) and travel to the GPU box."""
import os

import numpy as np
import pytest

from common import W, build
from ev_diffusion_b200.runtime import Simulation
from oracle import gen_oracle
from oracle.oracle import OracleSimulation

pytestmark = pytest.mark.gpu


def _pair(key):
    w = W.MODEL_WORKLOADS[key]
    plugin, so = build.model_path(w.name + "_exact"), gen_oracle.oracle_path(w.name)
    if not (os.path.exists(plugin) and os.path.exists(so)):
        pytest.skip(f"{w.name}_exact / its oracle are not prebuilt (they need the reference tree at build time)")
    st = W.model_state(w)
    gpu, orc = Simulation(plugin), OracleSimulation(so, threads=os.cpu_count() or 1)
    for s_ in (gpu, orc):
        s_.set_agents(w.agent, w.state, st)
    return w, gpu, orc


def test_boids_4m_one_iteration_exact():
    """C3 at its BASELINE size: 4 194 304 boids, 128^3 cells (21-bit keys: three 7-bit passes), 32-byte message records.
    Boids_Partitioning only adds, multiplies, divides and takes square roots: with -fmad=false every variable is bit-identical."""
    w, gpu, orc = _pair("c3")
    gpu.step(1)
    orc.step(1)
    assert np.array_equal(gpu.message_order("location"), orc.message_order("location"))
    gs, gc = gpu.pbm("location")
    os_, oc = orc.pbm("location")
    assert np.array_equal(gc, oc) and np.array_equal(gs[oc > 0], os_[oc > 0])
    a, b = gpu.agents(w.agent), orc.agents(w.agent)
    for k in a:
        assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k
    gpu.close()
    orc.close()


def test_sph_1m_one_iteration():
    """C4's scripts on 1 000 000 particles at the shipped partition radius (50^3 cells, 8 particles per cell, two spatial
    messages, two function conditions moving particles between state lists).  Integer state, list order and both message
    orders are exact; floats go through powf (libdevice vs glibc): |a - b| <= 1e-4 * max|b| per variable."""
    w, gpu, orc = _pair("c4_1m")
    gpu.step(1)
    orc.step(1)
    for m in ("location", "density_pressure"):
        assert gpu.message_count(m) == orc.message_count(m)
        assert np.array_equal(gpu.message_order(m), orc.message_order(m)), m
    for st in ("default", "static"):
        n = orc.agent_count(w.agent, st)
        assert gpu.agent_count(w.agent, st) == n
        if n == 0:
            continue
        a, b = gpu.agents(w.agent, st), orc.agents(w.agent, st)
        for k in a:
            if a[k].dtype.kind in "iu":
                assert np.array_equal(a[k], b[k]), (st, k)
            else:
                scale = max(float(np.abs(b[k]).max()), 1e-30)
                assert float(np.abs(a[k].astype(np.float64) - b[k]).max()) <= 1e-4 * scale, (st, k)
    gpu.close()
    orc.close()
