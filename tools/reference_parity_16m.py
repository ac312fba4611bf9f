#!/usr/bin/env python
"""One-off parity check at the north-star size (too slow for the test suite: about 6 minutes, 2 GB of 0.xml in /tmp):
C2-16M -- 16 777 216 Brownian particles, 128 x 128 x 256 cells -- one iteration of the reference's own generated code
(oracle/build_ref.py, run on the CPU, reading the 0.xml with its own reader) against the oracle: every agent variable
and all 16 777 216 rand48 seeds must be byte-identical.  Result of the round-1 run: profiles/r01_notes.md.
Needs /root/reference (or a prebuilt oracle/_ref/Brownian_16M/libref_Brownian_16M.so)."""
import os, sys, time, tempfile, shutil
os.environ["CUDA_ON_HOST_BLOCK"] = "1024"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ev_diffusion_b200 import workloads as W
from oracle import build_ref, gen_oracle
from oracle.ref_sim import RefSimulation
from oracle.oracle import OracleSimulation
cfg = W.BROWNIAN["c2_16m"]
fdir = os.path.dirname(W.brownian_function_files()[0])
t = time.time(); lib = build_ref.build_ref(W.brownian_xml(cfg), cfg.name, function_dir=fdir); print("ref lib", round(time.time()-t), "s", flush=True)
st = W.brownian_state(cfg)
d = tempfile.mkdtemp(dir="/tmp")
t = time.time(); inp = W.write_states_xml(os.path.join(d, "0.xml"), "Particle", st); print("wrote", os.path.getsize(inp)/1e9, "GB in", round(time.time()-t), "s", flush=True)
t = time.time(); ref = RefSimulation(lib, W.brownian_xml(cfg), inp); print("reference read+init", round(time.time()-t), "s; agents", ref.agent_count("Particle", "default"), flush=True)
shutil.rmtree(d)
orc = OracleSimulation(gen_oracle.oracle_path(cfg.name), threads=os.cpu_count())
orc.set_agents("Particle", "default", st)
t = time.time(); ref.step(1); print("reference 1 iteration", round(time.time()-t), "s", flush=True)
t = time.time(); orc.step(1); print("oracle 1 iteration", round(time.time()-t, 1), "s", flush=True)
ok = True
for v in ("id", "x", "y", "z", "nbr"):
    a, b = ref.agent_var("Particle", "default", v), orc.agent_var("Particle", "default", v, cfg.n)
    same = np.array_equal(a.view(np.uint8), b.view(np.uint8)); ok &= same
    print(v, "identical" if same else "DIFFERENT", flush=True)
s = np.array_equal(ref.seeds(), orc.seeds()); ok &= s
print("seeds", "identical" if s else "DIFFERENT", len(ref.seeds()))
print("RESULT", "ALL IDENTICAL" if ok else "MISMATCH", "nbr max", int(orc.agent_var("Particle", "default", "nbr", cfg.n).max()))
