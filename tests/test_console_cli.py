"""The console program (csrc/fgpu_console.cpp) beside the reference's own (its main.cu, built for the host by
oracle/build_ref.build_ref_console): what can be compared without a device -- the usage text and exit code for missing
arguments and -h, and the path warnings printed before any device work.  Everything after device creation is covered on
the GPU by tests/test_gpu_face2_driver.py."""
import os
import re
import subprocess

import pytest

from common import REFERENCE, build

NAME = "CirclesPartitioning_float"
pytestmark = pytest.mark.skipif(not os.path.isdir(f"{REFERENCE}/FLAMEGPU/templates"), reason="the reference tree is not mounted")


def _programs():
    from oracle import build_ref
    ours, plugin = os.path.join(build.LIB, "fgpu_console"), build.model_path(NAME + "_exact")
    if not (os.path.exists(ours) and os.path.exists(plugin)):
        pytest.skip("fgpu_console / plugin not prebuilt")
    theirs = build_ref.build_ref_console(f"{REFERENCE}/examples/{NAME}/src/model/XMLModelFile.xml", NAME)
    return ours, plugin, theirs


def _run(cmd):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=60)
    return p.returncode, p.stdout


def _normalised_usage(text):
    """The reference prints the basename of its executable, this program argv[0] as given; this program takes the model
    plugin as an extra first positional argument."""
    text = re.sub(r"usage: \S+", "usage: PROGRAM", text).replace(" model_plugin input_path", " input_path")
    return "\n".join(line for line in text.splitlines() if not line.startswith("  model_plugin "))


@pytest.mark.parametrize("args", ["none", "-h", "--help", "one positional short"])
def test_usage_text_and_exit_code(args, tmp_path):
    ours, plugin, theirs = _programs()
    inp = str(tmp_path / "0.xml")
    mine = {"none": [], "one positional short": [plugin, inp]}.get(args, [args])
    ref = {"none": [], "one positional short": [inp]}.get(args, [args])
    rc_o, out_o = _run([ours] + mine)
    rc_t, out_t = _run([theirs] + ref)
    assert rc_o == rc_t == 1
    assert _normalised_usage(out_o) == _normalised_usage(out_t)
    assert "usage:" in out_o and "XML_output_frequency" in out_o


@pytest.mark.parametrize("case", ["missing_parent", "missing_file"])
def test_path_warnings(case, tmp_path):
    ours, plugin, theirs = _programs()
    (tmp_path / "sub").mkdir()
    inp = str(tmp_path / "nosuch" / "0.xml") if case == "missing_parent" else str(tmp_path / "sub" / "missing.xml")
    _, out_o = _run([ours, plugin, inp, "1", "0", "0"])
    _, out_t = _run([theirs, inp, "1", "0", "0"])
    warn = lambda text: [line for line in text.splitlines() if line.startswith("Warning:")]     # noqa: E731
    assert warn(out_o) == warn(out_t) and len(warn(out_t)) == 1


def test_reference_console_run_is_what_the_gpu_test_expects(tmp_path):
    """tests/test_gpu_face2_driver.py::test_model_agnostic_console_mode states what fgpu_console must print and write
    for `0.xml 7 0 3` (and for bad arguments).  The same command through the reference's own console program gives
    exactly those lines, files and exit codes, and its last iteration file is the one the oracle's state produces
    through the writer restatement -- so the expectations of the GPU test are the reference's behaviour, not a reading
    of main.xslt."""
    import numpy as np
    from common import W, brownian_oracle
    from oracle import build_ref
    from oracle.xml_states import write_states
    cfg = W.BROWNIAN["c2_tiny"]
    fdir = os.path.dirname(W.brownian_function_files()[0])
    exe = build_ref.build_ref_console(W.brownian_xml(cfg), cfg.name, function_dir=fdir)
    st = W.brownian_state(cfg, 500)
    d = tmp_path / "run"
    d.mkdir()
    xml = W.write_states_xml(str(d / "0.xml"), "Particle", st)
    rc, out = _run([exe, xml, "7", "0", "3"])
    assert rc == 0, out
    lines = out.splitlines()
    assert lines[0] == "FLAMEGPU Console mode"
    assert [l for l in lines if l.startswith("Processing Simulation Step")] == [f"Processing Simulation Step {i}" for i in range(1, 8)]
    assert [l for l in lines if l.endswith("Saved to XML")] == [f"Iteration {i} Saved to XML" for i in (3, 6, 7)]
    assert any(l.startswith("Total Processing time:") for l in lines)
    assert sorted(p.name for p in d.iterdir()) == ["0.xml", "3.xml", "6.xml", "7.xml"]
    orc = brownian_oracle("c2_tiny")
    orc.set_agents("Particle", "default", st)
    orc.step(7)
    cols = {v: orc.agent_var("Particle", "default", v, 500) for v in ("id", "x", "y", "z", "nbr")}
    consts = [(name, ctype, orc.get_constant(name)) for name, ctype, _ in orc.model.constants]
    mine = tmp_path / "mine.xml"
    write_states(str(mine), 7, [("Particle", [("default", [("id", "int"), ("x", "float"), ("y", "float"), ("z", "float"), ("nbr", "int")])])],
                 {("Particle", "default"): cols}, consts)
    assert mine.read_bytes() == (d / "7.xml").read_bytes()
    orc.close()
    rc, out = _run([exe, xml, "2", "0", "0"])
    assert rc == 0 and "Saved to XML" not in out
    assert _run([exe, xml])[0] != 0
    rc, out = _run([exe, xml, "-4"])
    assert rc != 0 and "positive integer" in out
