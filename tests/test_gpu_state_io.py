"""State writer and checkpoint/resume.

* fgpu_sim_save_states writes the reference's iteration file format (io.xslt:124-200): compared byte for
  byte with the Python restatement of saveIterationData (oracle/xml_states.py::write_states) and read back
  through both readers.  The format prints floats with %f, so an XML round trip is lossy by design.
* fgpu_sim_save_snapshot / load_snapshot resume a run exactly: state lists, rand48 seeds, environment
  constants and the iteration counter survive, and the continued run is bit-identical to the uninterrupted one."""
import os

import numpy as np
import pytest

from common import ROOT, W, brownian_plugin, build
from ev_diffusion_b200.runtime import FgpuError, Simulation
from oracle.xml_states import read_states, write_states

pytestmark = pytest.mark.gpu

SPEC = [("Particle", [("id", "int", 0), ("x", "float", 0), ("y", "float", 0), ("z", "float", 0), ("nbr", "int", 0)])]
CONSTS = [("SIGMA", "float"), ("DOMAIN_X", "float"), ("DOMAIN_Y", "float"), ("DOMAIN_Z", "float")]


def _same(a, b):
    return all(np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)) for k in a)


def test_written_file_is_the_reference_format_and_reads_back(tmp_path):
    cfg = W.BROWNIAN["c2_tiny"]
    sim = Simulation(brownian_plugin("c2_tiny", exact=True))
    sim.set_agents("Particle", "default", W.brownian_state(cfg, 300))
    sim.set_constant("SIGMA", np.float32(0.3))
    sim.step(3)
    path = sim.save_states(str(tmp_path) + os.sep, 3)
    assert path.endswith(os.sep + "3.xml")
    got = sim.agents("Particle")
    consts = [(n, t, sim.get_constant(n)) for n, t in CONSTS]
    ref = tmp_path / "restated.xml"
    write_states(str(ref), 3, [("Particle", [("default", [(v, t) for v, t, _ in SPEC[0][1]])])], {("Particle", "default"): got}, consts)
    assert open(path, "rb").read() == ref.read_bytes()
    # both readers agree on what the file says; floats come back rounded to 6 decimals (%f)
    want, env = read_states(path, SPEC, CONSTS)
    back = Simulation(brownian_plugin("c2_tiny", exact=True))
    back.load_states(path)
    again = back.agents("Particle")
    assert _same(again, want["Particle"])
    assert np.array_equal(again["id"], got["id"]) and np.array_equal(again["nbr"], got["nbr"])
    assert np.abs(again["x"] - got["x"]).max() <= 1e-6 and not np.array_equal(again["x"], got["x"])
    assert np.float32(back.get_constant("SIGMA")) == np.float32(0.3)
    sim.close()
    back.close()


def test_empty_population_writes_an_empty_states_file(tmp_path):
    sim = Simulation(brownian_plugin("c2_tiny", exact=True))
    path = sim.save_states(str(tmp_path / "it_"), 0)
    text = open(path).read()
    assert text.startswith("<states>\n<itno>0</itno>\n<environment>\n\t<SIGMA>") and text.endswith("</environment>\n</states>\n")
    assert "<xagent>" not in text
    with pytest.raises(FgpuError):
        sim.save_states(str(tmp_path / "no_such_dir" / "x"), 0)
    sim.close()


def _birthdeath(n=3000):
    sim = Simulation(build.model_path("BirthDeath_exact"))
    rng = np.random.default_rng(5)
    cols = {"id": np.arange(n, dtype=np.int32), "x": rng.random(n, dtype=np.float32) * 16, "y": rng.random(n, dtype=np.float32) * 16,
            "z": rng.random(n, dtype=np.float32) * 16, "age": np.zeros(n, np.int32), "crowd": np.zeros(n, np.int32)}
    sim.set_agents("Cell", "default", cols)
    return sim


def test_snapshot_resume_is_bit_exact(tmp_path):
    """BirthDeath: rand48, deaths, births -- the continued run must not differ from the uninterrupted one."""
    a = _birthdeath()
    a.set_constant("BIRTH_P", np.float32(0.07))
    a.step(6)
    snap = str(tmp_path / "run.fgsnap")
    a.save_snapshot(snap)
    a.step(7)
    b = Simulation(build.model_path("BirthDeath_exact"))
    b.load_snapshot(snap)
    assert b.iteration == 6 and np.float32(b.get_constant("BIRTH_P")) == np.float32(0.07)
    b.step(7)
    assert a.iteration == b.iteration == 13
    assert a.agent_count("Cell", "default") == b.agent_count("Cell", "default") > 0
    assert _same(a.agents("Cell"), b.agents("Cell"))
    assert np.array_equal(a.seeds(), b.seeds())
    a.close()
    b.close()


def test_snapshot_of_another_model_is_refused(tmp_path):
    a = _birthdeath(10)
    snap = str(tmp_path / "bd.fgsnap")
    a.save_snapshot(snap)
    other = Simulation(brownian_plugin("c2_tiny"))
    with pytest.raises(FgpuError, match="another model"):
        other.load_snapshot(snap)
    with open(snap, "r+b") as fh:
        fh.truncate(200)
    with pytest.raises(FgpuError, match="truncated|differs"):
        a.load_snapshot(snap)
    with pytest.raises(FgpuError):
        a.load_snapshot(str(tmp_path / "missing.fgsnap"))
    a.close()
    other.close()
