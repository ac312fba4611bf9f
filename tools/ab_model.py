"""A/B helper for the model workloads of bench.py: stage times and cold-L2 iteration time of several builds of one model.

usage: python tools/ab_model.py c3 Boids_4M Boids_4M_mb10 Boids_4M_mb12[:option=value]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ev_diffusion_b200 import build as B, workloads as W
from ev_diffusion_b200.runtime import Simulation

w = W.MODEL_WORKLOADS[sys.argv[1]]
st = W.model_state(w)
ref = None
for spec in sys.argv[2:]:
    name, *opts = spec.split(":")
    sim = Simulation(B.model_path(name), device=0)
    for kv in opts:
        k, v = kv.split("=")
        sim.set_option(k, int(v))
    W.start_model(sim, w, st)
    sim.step(3)
    sim.profile_iterations(3)
    stages = sim.profile_iterations(10)
    ms = sim.step_timed(10, 256 << 20)
    got = {a.name + "/" + s: sim.agents(a.name, s) for a in sim.model.agents for s in a.states if sim.agent_count(a.name, s) > 0}
    if ref is None:
        ref = got
    same = set(got) == set(ref) and all(np.array_equal(got[k][v].view(np.uint8), ref[k][v].view(np.uint8)) for k in got for v in got[k])
    print(spec, "cold ms/iter %.4f" % float(np.mean(ms)), {k: round(v * 1e3, 1) for k, v in stages}, "identical to first:", bool(same), flush=True)
    sim.close()
