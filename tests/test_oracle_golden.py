"""The CPU oracle against the committed golden fixtures (tests/golden/, made by make_golden.py from the
reference's own examples/<model>/iterations/0.xml inputs): replaying the fixture's initial state must give
the fixture's final state bit for bit -- integers, floats, doubles, per-state counts, rand48 seeds.
The reference ships no expected outputs, so this pins the oracle over time (a change in oracle_rt.hpp or
gen_oracle.py that alters any result fails here) and checks that single-thread and OpenMP runs agree."""
import os

import numpy as np
import pytest

from common import REFERENCE, ROOT, OracleSimulation, gen_oracle

GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = {"CirclesPartitioning_float": 100, "Boids_Partitioning": 50, "Keratinocyte": 200,
         "CirclesBruteForce_float": 20, "CirclesBruteForce_double": 20, "CirclesPartitioning_double": 50,
         "SocialParticleSimulation": 20, "EmptyExample": 5, "Analytics": 10, "MonteCarlo_MSMPR": 5}


def _oracle(name):
    so = gen_oracle.oracle_path(name)
    if not os.path.exists(so):
        xml = f"{REFERENCE}/examples/{name}/src/model/XMLModelFile.xml"
        if not os.path.exists(xml):
            pytest.skip("oracle library not prebuilt and the reference tree is not mounted")
        so = gen_oracle.build_oracle(xml, name)
    return so


def _replay(name, threads, iters):
    sim = OracleSimulation(_oracle(name), threads=threads)
    init = np.load(os.path.join(GOLDEN, f"{name}_initial.npz"))
    for k in init.files:
        if k.startswith("env/"):
            sim.set_constant(k[4:], init[k])
    for a in sim.model.agents:
        sim.set_agents(a.name, a.states[a.initial_state], {v.name: init[f"{a.name}/{v.name}"] for v in a.vars})
    sim.run_init_functions()
    sim.step(iters)
    return sim


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden_fixture(name):
    iters = CASES[name]
    want = np.load(os.path.join(GOLDEN, f"{name}_iter{iters}.npz"))
    sim = _replay(name, 1, iters)
    for a in sim.model.agents:
        for s in a.states:
            n = int(want[f"count/{a.name}/{s}"])
            assert sim.agent_count(a.name, s) == n, (a.name, s)
            for v in a.vars:
                got = sim.agent_var(a.name, s, v.name, n)
                ref = want[f"agent/{a.name}/{s}/{v.name}"]
                assert got.dtype == ref.dtype and np.array_equal(got.view(np.uint8), ref.view(np.uint8)), (a.name, s, v.name)
    assert np.array_equal(sim.seeds(), want["seeds"])
    sim.close()


def test_oracle_reproduces_the_brownian_fixture():
    """The flagship workload (C2 at parity size): seeded synthetic input from workloads.py, 10 iterations."""
    from common import W, brownian_oracle
    want = np.load(os.path.join(GOLDEN, "Brownian_small_iter10.npz"))
    sim = brownian_oracle("c2_small")
    sim.set_agents("Particle", "default", W.brownian_state(W.BROWNIAN["c2_small"]))
    sim.step(10)
    n = int(want["count/Particle/default"])
    assert sim.agent_count("Particle", "default") == n
    for v in ("id", "x", "y", "z", "nbr"):
        got, ref = sim.agent_var("Particle", "default", v, n), want[f"agent/Particle/default/{v}"]
        assert got.dtype == ref.dtype and np.array_equal(got.view(np.uint8), ref.view(np.uint8)), v
    assert np.array_equal(sim.seeds(), want["seeds"])
    sim.close()


@pytest.mark.parametrize("name", ["Boids_Partitioning", "CirclesBruteForce_float", "SocialParticleSimulation"])
def test_openmp_oracle_equals_sequential_oracle(name):
    """bench.py's CPU legs run the oracle with OpenMP over agents; per-agent work is independent, so the
    thread count must not change a single bit."""
    iters = min(CASES[name], 10)
    a, b = _replay(name, 1, iters), _replay(name, 4, iters)
    for ag in a.model.agents:
        for s in ag.states:
            n = a.agent_count(ag.name, s)
            assert b.agent_count(ag.name, s) == n
            for v in ag.vars:
                assert np.array_equal(a.agent_var(ag.name, s, v.name, n).view(np.uint8), b.agent_var(ag.name, s, v.name, n).view(np.uint8))
    a.close()
    b.close()
