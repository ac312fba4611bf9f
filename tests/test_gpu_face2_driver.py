"""Drop-in boundary, host face: a console driver written only against the reference's
generated host API names (initialise, singleIteration, cleanup, get_agent_<A>_<S>_count,
get_<A>_<S>_variable_<v>; header.xslt:531-645) links against a model plugin +
libfgpu_rt.so and reproduces the run driven through the C ABI."""
import os
import re
import subprocess

import numpy as np
import pytest

from common import W, brownian_plugin, build
from ev_diffusion_b200.runtime import Simulation

pytestmark = pytest.mark.gpu


def test_console_driver_matches_c_abi(tmp_path):
    exe = os.path.join(build.LIB, "drivers", "console_Brownian_tiny")
    if not os.path.exists(exe):
        pytest.skip("driver not prebuilt")
    cfg = W.BROWNIAN["c2_tiny"]
    st = W.brownian_state(cfg, 3000)
    xml = W.write_states_xml(str(tmp_path / "0.xml"), "Particle", st)
    out = subprocess.run([exe, xml, "7"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert out.returncode == 0, out.stdout
    assert "Total Processing time:" in out.stdout
    m = re.search(r"iteration (\d+) count (\d+) max (\d+)", out.stdout)
    assert m and (int(m.group(1)), int(m.group(2)), int(m.group(3))) == (7, 3000, cfg.n)
    sim = Simulation(brownian_plugin("c2_tiny", exact=True))
    sim.load_states(xml)
    sim.step(7)
    a = sim.agents("Particle")
    rows = re.findall(r"agent (\d+) id (-?\d+) x (\S+) y (\S+) z (\S+) nbr (-?\d+)", out.stdout)
    assert len(rows) == 8
    for i, idv, x, y, z, nbr in rows:
        i = int(i)
        assert int(idv) == a["id"][i] and int(nbr) == a["nbr"][i]
        assert np.float32(x) == a["x"][i] and np.float32(y) == a["y"][i] and np.float32(z) == a["z"][i]
    sim.close()
    # error behaviour of the reference: message + exit(EXIT_FAILURE) (simulation.xslt:198-206)
    bad = subprocess.run([exe, str(tmp_path / "missing.xml"), "1"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert bad.returncode != 0 and "Could not open" in bad.stdout


def test_model_agnostic_console_mode(tmp_path):
    """csrc/fgpu_console.cpp: the reference's console mode (main.xslt:60-448) for any plugin -- arguments,
    messages, iteration files every XML_output_frequency iterations plus the last one, in the directory of the
    input file -- against the same run through the Python mirror of the C ABI."""
    exe = os.path.join(build.LIB, "fgpu_console")
    if not os.path.exists(exe):
        pytest.skip("console driver not prebuilt")
    cfg = W.BROWNIAN["c2_tiny"]
    st = W.brownian_state(cfg, 500)
    d = tmp_path / "run"
    d.mkdir()
    xml = W.write_states_xml(str(d / "0.xml"), "Particle", st)
    plugin = brownian_plugin("c2_tiny", exact=True)
    out = subprocess.run([exe, plugin, xml, "7", "0", "3"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert out.returncode == 0, out.stdout
    lines = out.stdout.splitlines()
    assert lines[0] == "FLAMEGPU Console mode"
    assert [l for l in lines if l.startswith("Processing Simulation Step")] == [f"Processing Simulation Step {i}" for i in range(1, 8)]
    assert [l for l in lines if l.endswith("Saved to XML")] == [f"Iteration {i} Saved to XML" for i in (3, 6, 7)]
    assert any(l.startswith("Total Processing time:") for l in lines)
    assert sorted(p.name for p in d.iterdir()) == ["0.xml", "3.xml", "6.xml", "7.xml"]
    sim = Simulation(plugin)
    sim.load_states(xml)
    sim.step(7)
    ref = sim.save_states(str(tmp_path / "ref_"), 7)
    assert open(ref, "rb").read() == (d / "7.xml").read_bytes()
    sim.close()
    # the reference's instrumentation lines (simulation.xslt:913,973), switched on from the environment
    out = subprocess.run([exe, plugin, xml, "2", "0", "0"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120,
                         env=dict(os.environ, FGPU_INSTRUMENT="iterations,agent_functions"))
    assert out.returncode == 0, out.stdout
    import re
    ins = [l for l in out.stdout.splitlines() if l.startswith("Instrumentation: ")]
    names = [re.match(r"Instrumentation: (.*) = [0-9.]+ \(ms\)$", l).group(1) for l in ins]
    assert names == ["Particle_output_location", "Particle_count_neighbours", "Particle_move", "Iteration Time"] * 2, out.stdout
    # no XML output, usage and argument errors
    out = subprocess.run([exe, plugin, xml, "2", "0", "0"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert out.returncode == 0 and "Saved to XML" not in out.stdout
    assert subprocess.run([exe, plugin, xml], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).returncode != 0
    bad = subprocess.run([exe, plugin, xml, "-4"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert bad.returncode != 0 and "positive integer" in bad.stdout
    bad = subprocess.run([exe, str(tmp_path / "nope.so"), xml, "1"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert bad.returncode != 0 and "nope.so" in bad.stdout
