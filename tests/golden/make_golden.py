#!/usr/bin/env python
"""Generates the golden fixtures under tests/golden/ (run in the build container,
where /root/reference is mounted; the GPU box only sees the committed .npz files).

For each reference example on the hot path:
  <model>_initial.npz   the initial state lists parsed from the reference's own
                        examples/<model>/iterations/0.xml by the Python restatement
                        of the reference reader (oracle/xml_states.py, io.xslt:211-621)
  <model>_iter<N>.npz   the state after N iterations of the CPU oracle, which executes
                        the reference's own functions.c (compiled where it lies under
                        /root/reference; no reference source is copied into the repo)

The reference ships no expected outputs, so these fixtures keep the
oracle fixed over time and give the GPU parity tests a reference-input case that
travels.  The independent check is tests/test_oracle_vs_reference.py: the reference's
own generated code, built by oracle/build_ref.py and run on the CPU, reproduces
every one of these files bit for bit.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from ev_diffusion_b200 import xmml  # noqa: E402
from oracle import gen_oracle  # noqa: E402
from oracle.oracle import OracleSimulation  # noqa: E402
from oracle.xml_states import read_states  # noqa: E402

REF = "/root/reference/examples"
CASES = [("CirclesPartitioning_float", 100), ("Boids_Partitioning", 50), ("Keratinocyte", 200),
         # second batch: brute-force (non-partitioned) messages, double-precision variables, a second spatial model
         ("CirclesBruteForce_float", 20), ("CirclesBruteForce_double", 20), ("CirclesPartitioning_double", 50),
         ("SocialParticleSimulation", 20), ("EmptyExample", 5),
         # init/step/exit functions calling the host analytics API (reduce_/min_/max_)
         ("Analytics", 10),
         # brute-force messages + rand48 + death compaction + powf (float state compared with a tolerance on the GPU)
         ("MonteCarlo_MSMPR", 5)]


def snapshot(sim):
    out = {}
    for ai, a in enumerate(sim.model.agents):
        for si, s in enumerate(a.states):
            n = sim.agent_count(ai, si)
            out[f"count/{a.name}/{s}"] = np.int32(n)
            for v in a.vars:
                out[f"agent/{a.name}/{s}/{v.name}"] = sim.agent_var(ai, si, v.name, n)
    out["seeds"] = sim.seeds()
    return out


def brownian_fixture():
    """The flagship workload at parity size (c2_small: 16384 particles, 16^3 cells): state after 10 iterations of
    the oracle from the seeded synthetic input of ev_diffusion_b200.workloads (no 0.xml involved)."""
    from ev_diffusion_b200 import workloads as W
    cfg = W.BROWNIAN["c2_small"]
    so = gen_oracle.build_oracle(W.brownian_xml(cfg), cfg.name, function_files=W.brownian_function_files())
    sim = OracleSimulation(so)
    sim.set_agents("Particle", "default", W.brownian_state(cfg))
    sim.step(10)
    np.savez_compressed(os.path.join(HERE, "Brownian_small_iter10.npz"), **snapshot(sim))
    print("Brownian_small: 10 iterations,", sim.agent_count("Particle", "default"), "particles")
    sim.close()


def main():
    brownian_fixture()
    for name, iters in CASES:
        xml = f"{REF}/{name}/src/model/XMLModelFile.xml"
        model = xmml.parse_model(xml)
        so = gen_oracle.build_oracle(xml, name)
        spec = [(a.name, [(v.name, v.type, float(v.default or 0)) for v in a.variables]) for a in model.agents]
        consts = [(c.name, c.type) for c in model.constants if not c.array_length]
        states, env = read_states(f"{REF}/{name}/iterations/0.xml", spec, consts)
        init = {}
        for a in model.agents:
            for v in a.variables:
                init[f"{a.name}/{v.name}"] = states[a.name][v.name]
        for k, v in env.items():
            init[f"env/{k}"] = np.asarray(v)
        np.savez_compressed(os.path.join(HERE, f"{name}_initial.npz"), **init)
        sim = OracleSimulation(so)
        for k, v in env.items():
            sim.set_constant(k, v)          # io.xslt:523-556: applied while reading
        for a in model.agents:
            sim.set_agents(a.name, a.initial_state, states[a.name])
        sim.run_init_functions()            # simulation.xslt:703-716: after the read
        sim.step(iters)
        snap = snapshot(sim)
        np.savez_compressed(os.path.join(HERE, f"{name}_iter{iters}.npz"), **snap)
        counts = {k: int(v) for k, v in snap.items() if k.startswith("count/")}
        print(name, "agents read:", {a.name: len(states[a.name][a.variables[0].name]) for a in model.agents},
              "after", iters, "iterations:", counts)
        sim.close()


if __name__ == "__main__":
    main()
