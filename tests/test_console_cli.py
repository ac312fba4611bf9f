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
