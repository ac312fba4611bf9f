"""Host-side analytics API: reduce_/min_/max_/count_<A>_<S>_<v>_variable()
(header.xslt:729-751, simulation.xslt:1325-1357) as device reductions behind fgpu_sim_reduce_agent_var, and
through the generated names when a script's step function calls them (examples/Analytics)."""
import os

import numpy as np
import pytest

from common import W, brownian_oracle, brownian_plugin, build
from ev_diffusion_b200.runtime import FgpuError, Simulation

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("key,count", [("c2_tiny", 1), ("c2_tiny", 4096), ("c2", None)])
def test_reductions_against_numpy_and_oracle(key, count):
    cfg = W.BROWNIAN[key]
    st = W.brownian_state(cfg, count)
    sim = Simulation(brownian_plugin(key))
    sim.set_agents("Particle", "default", st)
    sim.step(2)
    a = sim.agents("Particle")
    # integers: exact, with the wrap-around of the variable's own type (the reference returns thrust::reduce<int>)
    wrapped = np.array(a["id"].astype(np.int64).sum() & 0xFFFFFFFF, dtype=np.uint32).astype(np.int32)
    assert sim.reduce("Particle", "default", "id", "sum") == wrapped
    for k in ("id", "nbr"):
        assert sim.reduce("Particle", "default", k, "min") == a[k].min()
        assert sim.reduce("Particle", "default", k, "max") == a[k].max()
    for v in (0, int(a["nbr"].max()), 12345678):
        assert sim.reduce("Particle", "default", "nbr", "count", v) == np.count_nonzero(a["nbr"] == v)
    # floats: min/max exact, the sum a fixed-order float32 tree (reference: order unspecified)
    for k in "xyz":
        assert sim.reduce("Particle", "default", k, "min") == a[k].min()
        assert sim.reduce("Particle", "default", k, "max") == a[k].max()
        s1 = sim.reduce("Particle", "default", k, "sum")
        assert s1 == sim.reduce("Particle", "default", k, "sum")          # deterministic
        assert abs(float(s1) - a[k].astype(np.float64).sum()) <= 2e-6 * max(1.0, a[k].astype(np.float64).sum())
    if key == "c2_tiny":
        orc = brownian_oracle(key)
        orc.set_agents("Particle", "default", st)
        orc.step(2)
        for k, op in (("id", "sum"), ("nbr", "max"), ("x", "min"), ("z", "max")):
            assert orc.reduce("Particle", "default", k, op) == sim.reduce("Particle", "default", k, op)
        assert orc.reduce("Particle", "default", "nbr", "count", 3) == sim.reduce("Particle", "default", "nbr", "count", 3)
        orc.close()
    sim.close()


def test_reduction_edge_cases():
    sim = Simulation(brownian_plugin("c2_tiny"))
    assert sim.reduce("Particle", "default", "x", "sum") == 0.0          # empty list
    assert sim.reduce("Particle", "default", "nbr", "count", 0) == 0
    with pytest.raises(FgpuError, match="empty"):
        sim.reduce("Particle", "default", "x", "min")
    sim.set_agents("Particle", "default", W.brownian_state(W.BROWNIAN["c2_tiny"], 10))
    with pytest.raises(FgpuError, match="integer"):
        sim.reduce("Particle", "default", "x", "count", 1.0)
    sim.close()


def test_analytics_step_function_runs_the_generated_names(capfd):
    """examples/Analytics: its step function prints mean/min/max positions through reduce_/min_/max_."""
    p = build.model_path("Analytics_exact")
    if not os.path.exists(p):
        pytest.skip("Analytics plugin not prebuilt (needs /root/reference at build time)")
    sim = Simulation(p)
    rng = np.random.default_rng(1)
    n = 500
    cols = {v.name: np.zeros(n, v.dtype) for v in sim.model.agents[0].vars}
    cols["id"] = np.arange(n, dtype=np.int32)
    cols["x"] = (rng.random(n, dtype=np.float32) * 20 - 10).astype(np.float32)
    cols["y"] = (rng.random(n, dtype=np.float32) * 20 - 10).astype(np.float32)
    sim.set_agents(0, 0, cols)
    sim.run_init_functions()
    sim.step(1)
    a = sim.agents(0)
    out = capfd.readouterr().out
    assert "FLAME GPU Step function. Min circle position is (%f, %f, %f)" % (a["x"].min(), a["y"].min(), a["z"].min()) in out
    assert "FLAME GPU Step function. Max circle position is (%f, %f, %f)" % (a["x"].max(), a["y"].max(), a["z"].max()) in out
    sim.close()


def test_host_agent_creation_from_init_and_step_functions():
    """models/HostCreate: h_allocate_agent_<A>[_array], h_add_agent(s)_<A>_<S>, h_free_agent_<A>[_array]
    (header.xslt:665-706) called from an init and a step function, XML default values, the buffer-size guard of
    the step function, and max_<A>_<S>_<v>_variable() feeding back into the new agents -- bit for bit against
    the oracle, which runs the same functions.c."""
    from common import gen_oracle, OracleSimulation
    xml = os.path.join(build.MODELS, "HostCreate", "src", "model", "XMLModelFile.xml")
    p = build.model_path("HostCreate_exact")
    if not os.path.exists(p):
        p = build.build_model(xml, "HostCreate_exact", fmad=False)
    so = gen_oracle.oracle_path("HostCreate")
    if not os.path.exists(so):
        so = gen_oracle.build_oracle(xml, "HostCreate")
    gpu, orc = Simulation(p), OracleSimulation(so)
    for s in (gpu, orc):
        s.run_init_functions()
    assert gpu.agent_count("Item", "default") == orc.agent_count("Item", "default") == 50
    a = gpu.agents("Item")
    assert np.all(a["y"] == np.float32(2.5)) and np.all(a["age"] == 7) and np.array_equal(a["id"], np.arange(50))
    for it in range(1, 9):
        gpu.step(1)
        orc.step(1)
        n = 50 + 3 * it
        assert gpu.agent_count("Item", "default") == orc.agent_count("Item", "default") == n
        a, b = gpu.agents("Item"), orc.agents("Item")
        for k in a:
            assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), (it, k)
    # the newcomers of the last step: ids continue, x = oldest age + k, crowd = population before them
    assert np.array_equal(a["id"][-3:], [71, 72, 73]) and np.array_equal(a["crowd"][-3:], [71, 71, 71])
    assert np.array_equal(a["x"][-3:], np.float32([15, 16, 17]))
    gpu.close()
    orc.close()


def test_generate_id_on_device_and_host_and_sort_agents(tmp_path):
    """models/IdSort: generate_<A>_id() on the device (atomic counter) and in a step function on the host, the
    counter initialised from the largest id of the 0.xml (io.xslt:599-616) and synchronised around the step
    functions (simulation.xslt:918-960); sort_<A>s_<S> with the model's own key/value kernel, stable among equal
    keys.  Device-side tickets arrive in no particular order, so the checks are on sets and on the sort."""
    xml = os.path.join(build.MODELS, "IdSort", "src", "model", "XMLModelFile.xml")
    p = build.model_path("IdSort_exact")
    if not os.path.exists(p):
        p = build.build_model(xml, "IdSort_exact", fmad=False)
    n = 5000
    st = {"id": np.arange(n, dtype=np.int32), "x": np.arange(n, dtype=np.float32), "ticket": np.zeros(n, dtype=np.int32)}
    inp = W.write_states_xml(str(tmp_path / "0.xml"), "Item", st)
    sim = Simulation(p)
    sim.load_states(inp)
    sim.step(1)
    a = sim.agents("Item")
    assert len(a["id"]) == n + 1
    # every agent drew one ticket, the first being the largest id read + 1
    assert np.array_equal(np.sort(a["ticket"][:n]), np.arange(n, 2 * n, dtype=np.int32))
    # sorted by ticket % 7, and among equal keys in the order the list had before (ids ascending): stable
    key = a["ticket"][:n] % 7
    assert (np.diff(key) >= 0).all()
    for k in range(7):
        assert (np.diff(a["id"][:n][key == k]) > 0).all()
    assert np.array_equal(a["x"][:n], a["id"][:n].astype(np.float32))      # whole agents moved, not single columns
    # the host-created agent took the next value of the same counter
    assert a["id"][n] == 2 * n and a["ticket"][n] == -1
    sim.step(1)
    b = sim.agents("Item")
    assert len(b["id"]) == n + 2 and b["id"][-1] == 3 * n + 2
    assert np.array_equal(np.sort(b["ticket"][:n + 1]), np.arange(2 * n + 1, 3 * n + 2, dtype=np.int32))
    sim.close()
