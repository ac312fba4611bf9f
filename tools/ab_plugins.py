"""A/B helper: steady-state stage times of several builds of the C2 Brownian plugin on one GPU.

usage: python tools/ab_plugins.py Brownian Brownian_generic:sm_local=0 ...   (plugin[:option=value]...)
       AB_WORKLOAD=c2_16m python tools/ab_plugins.py Brownian_16M Brownian_16M_pf
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ev_diffusion_b200 import build as B, workloads as W
from ev_diffusion_b200.runtime import Simulation

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cfg = W.BROWNIAN[os.environ.get("AB_WORKLOAD", "c2")]
state = W.brownian_state(cfg)
ref = None
for spec in sys.argv[1:]:
    name, *opts = spec.split(":")          # Brownian:sm_local=0:order=1
    sim = Simulation(B.model_path(name), device=0)
    for kv in opts:
        k, v = kv.split("=")
        sim.set_option(k, int(v))
    sim.set_agents("Particle", "default", state)
    sim.step(5)
    sim.profile_iterations(5)
    st = sim.profile_iterations(20)
    ms = sim.step_timed(30, 256 << 20)
    nbr = sim.agent_var("Particle", "default", "nbr")
    x = sim.agent_var("Particle", "default", "x")
    if ref is None:
        ref = (nbr.copy(), x.copy())
    same = bool(np.array_equal(ref[0], nbr) and np.array_equal(ref[1], x))
    print(spec, "cold ms/iter %.4f" % float(np.mean(ms)), {k: round(v * 1e3, 1) for k, v in st}, "identical to first:", same,
          flush=True)
