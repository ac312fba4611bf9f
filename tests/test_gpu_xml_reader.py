"""The C++ 0.xml reader of the runtime (fgpu_sim_load_states) against the Python
restatement of the reference reader (oracle/xml_states.py, io.xslt:211-621) on
authored files that exercise its quirks: comments, unknown tags, values with
surrounding whitespace, hex/octal integers (strtol base 0), missing variables
(defaultValue or 0), an <environment> block, a file that is one single line."""
import os

import numpy as np
import pytest

from common import W, brownian_plugin
from ev_diffusion_b200.runtime import FgpuError, Simulation
from oracle.xml_states import read_states

pytestmark = pytest.mark.gpu

SPEC = [("Particle", [("id", "int", 0), ("x", "float", 0), ("y", "float", 0), ("z", "float", 0), ("nbr", "int", 0)])]
CONSTS = [("SIGMA", "float"), ("DOMAIN_X", "float"), ("DOMAIN_Y", "float"), ("DOMAIN_Z", "float")]

QUIRKY = """<states>
<itno>17</itno>
<!-- a comment with <xagent> inside -- and dashes -->
<environment>
  <SIGMA> 0.125 </SIGMA>
  <unknown_constant>9</unknown_constant>
</environment>
<xagent><name>Particle</name><id>0x10</id><x>1.5</x><y>
 2.25e0
</y><z>3</z><range>77</range></xagent>
<xagent>
  <name>Particle</name>
  <id>010</id>
  <x>-0.0</x>
  <nbr>5</nbr>
</xagent>
<xagent><name>Ghost</name><id>3</id></xagent>
<xagent><name>Particle</name><id>7</id><x>1e-3</x><y>4.000000001</y><z>15.9999999</z></xagent>
</states>
<xagent><name>Particle</name><id>99</id></xagent>
"""


def _check(path):
    want, env = read_states(path, SPEC, CONSTS)
    sim = Simulation(brownian_plugin("c2_tiny"))
    sim.load_states(path)
    n = len(want["Particle"]["id"])
    assert sim.agent_count("Particle", "default") == n
    got = sim.agents("Particle")
    for k in got:
        assert np.array_equal(got[k].view(np.uint32), want["Particle"][k].view(np.uint32)), k
    for k, v in env.items():
        assert np.float32(sim.get_constant(k)) == np.float32(v)
    sim.close()
    return want, env


def test_quirks(tmp_path):
    p = tmp_path / "0.xml"
    p.write_text(QUIRKY)
    want, env = _check(str(p))
    assert want["Particle"]["id"].tolist() == [16, 8, 7]           # 0x10, 010 (octal), 7; Ghost ignored; after </states> ignored
    assert want["Particle"]["nbr"].tolist() == [0, 5, 0]
    assert env == {"SIGMA": np.float32(0.125)}


def test_generated_file_and_single_line_file(tmp_path):
    cfg = W.BROWNIAN["c2_tiny"]
    st = W.brownian_state(cfg, 500)
    p = W.write_states_xml(str(tmp_path / "gen.xml"), "Particle", st, env={"SIGMA": 0.5})
    want, _ = _check(p)
    for k in st:   # %.9g round-trips binary32
        assert np.array_equal(want["Particle"][k].view(np.uint32), st[k].view(np.uint32))
    one_line = tmp_path / "one_line.xml"
    one_line.write_text(open(p).read().replace("\n", ""))   # like Boids_Partitioning/iterations/0.xml
    _check(str(one_line))


def test_overflow_and_missing_file(tmp_path):
    cfg = W.BROWNIAN["c2_tiny"]
    sim = Simulation(brownian_plugin("c2_tiny"))
    with pytest.raises(FgpuError, match="Could not open"):
        sim.load_states(str(tmp_path / "nope.xml"))
    st = W.brownian_state(cfg, cfg.n + 1)
    p = W.write_states_xml(str(tmp_path / "big.xml"), "Particle", st)
    with pytest.raises(FgpuError, match="MAX Buffer size"):   # io.xslt:406 (without its off-by-one)
        sim.load_states(p)
    sim.close()
