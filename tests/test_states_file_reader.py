"""The CUDA path's own 0.xml reader (csrc/fgpu_rt_reader.inl, the parser behind fgpu_sim_load_states) run WITHOUT a device
through fgpu_states_file_count / fgpu_states_file_read_var: against the reference's reader (the reference build of
oracle/build_ref.py) on the shipped inputs and on an awkward file, against the golden initial states, and on a
generated file.  The GPU tests (tests/test_gpu_xml_reader.py) cover the upload that follows the parse."""
import os
import subprocess
import sys

import numpy as np
import pytest

from common import REFERENCE, ROOT, W, build
from ev_diffusion_b200.runtime import FgpuError, read_states_file
from test_oracle_golden import CASES, GOLDEN

HAVE_REFERENCE = os.path.isdir(f"{REFERENCE}/examples")


def _plugin(name):
    p = build.model_path(name + "_exact")
    if not os.path.exists(p):
        pytest.skip(f"plugin {name}_exact is not prebuilt")
    return p


@pytest.mark.skipif(not HAVE_REFERENCE, reason="the reference tree is not mounted (shipped 0.xml inputs)")
@pytest.mark.parametrize("name", sorted(CASES))
def test_shipped_inputs_equal_the_golden_initial_states(name):
    got = read_states_file(_plugin(name), f"{REFERENCE}/examples/{name}/iterations/0.xml")
    want = np.load(os.path.join(GOLDEN, f"{name}_initial.npz"))
    keys = [k for k in want.files if not k.startswith("env/")]
    assert keys
    for k in keys:
        agent, var = k.split("/")
        assert got[agent][var].dtype == want[k].dtype and np.array_equal(got[agent][var].view(np.uint8), want[k].view(np.uint8)), k


@pytest.mark.skipif(not HAVE_REFERENCE, reason="the reference tree is not mounted")
def test_awkward_input_equals_the_reference_reader(tmp_path):
    """Same file as tests/test_oracle_vs_reference.py::test_reader_restatement_on_awkward_input, this time through the
    product's C++ parser: strtol base 0, atof corner cases, unknown tags, a comment holding tags, an agent of an unknown
    type (warning, skipped), defaults for variables left out, agents after </states>."""
    from oracle import build_ref
    from test_oracle_vs_reference import _NASTY_INPUT
    name = "CirclesPartitioning_float"
    xml = f"{REFERENCE}/examples/{name}/src/model/XMLModelFile.xml"
    inp = tmp_path / "0.xml"
    inp.write_text(_NASTY_INPUT)
    worker = ("import sys, numpy as np\nsys.path.insert(0, sys.argv[1])\nfrom oracle.ref_sim import RefSimulation\n"
              "sim = RefSimulation(sys.argv[2], sys.argv[3], sys.argv[4])\n"
              "np.savez(sys.argv[5], **{v.name: sim.agent_var('Circle', 'default', v.name) for v in sim.model.agents[0].variables})\n")
    out = tmp_path / "read.npz"
    subprocess.run([sys.executable, "-c", worker, ROOT, build_ref.build_ref(xml, name), xml, str(inp), str(out)], check=True,
                   stdout=subprocess.DEVNULL)
    theirs = np.load(out)
    mine = read_states_file(_plugin(name), str(inp))["Circle"]
    assert mine["id"].tolist() == [31, -12, 15, 2147483647]
    for k in theirs.files:
        assert np.array_equal(theirs[k].view(np.uint8), mine[k].view(np.uint8)), (k, theirs[k], mine[k])


def test_generated_file_round_trip_and_errors(tmp_path):
    cfg = W.BROWNIAN["c2_tiny"]
    plugin = build.model_path(cfg.name + "_exact")
    if not os.path.exists(plugin):
        pytest.skip("Brownian_tiny_exact is not prebuilt")
    st = W.brownian_state(cfg, 777)
    path = W.write_states_xml(str(tmp_path / "0.xml"), "Particle", st)
    got = read_states_file(plugin, path)["Particle"]
    for k in st:
        assert np.array_equal(got[k].view(np.uint8), st[k].view(np.uint8)), k          # %.9g floats come back exactly
    with pytest.raises(FgpuError, match="Could not open"):
        read_states_file(plugin, str(tmp_path / "missing.xml"))
    # more agents than the model's bufferSize: the reader's guard (io.xslt:406) fires without a device too
    big = W.write_states_xml(str(tmp_path / "big.xml"), "Particle", W.brownian_state(cfg, cfg.buffer_size + 1))
    with pytest.raises(FgpuError, match="MAX Buffer size"):
        read_states_file(plugin, big)


def test_fast_decimal_path_equals_atof(tmp_path):
    """The parser's decimal fast path (std::from_chars) against libc's atof (what the reference calls, io.xslt:55-64) on values
    chosen to be awkward: long mantissas, halfway cases of binary32, exponents at both ends of the double range, subnormals,
    overflow to infinity, forms the fast path must leave to strtod (leading blanks, '+', hex floats, inf / nan, trailing text)."""
    import ctypes
    libc = ctypes.CDLL(None)
    libc.atof.restype = ctypes.c_double
    libc.atof.argtypes = [ctypes.c_char_p]
    rng = np.random.default_rng(5)
    texts = []
    for _ in range(16000):          # Brownian_small holds 16 384 agents
        k = rng.integers(0, 6)
        if k == 0:
            texts.append("%.9g" % np.float32(rng.standard_normal() * 10.0 ** rng.integers(-30, 30)))
        elif k == 1:
            texts.append("%d.%s" % (rng.integers(-10 ** 6, 10 ** 6), "".join(str(d) for d in rng.integers(0, 10, rng.integers(1, 40)))))
        elif k == 2:
            texts.append("%.17ge%d" % (rng.standard_normal(), rng.integers(-340, 320)))
        elif k == 3:      # exactly between two binary32 neighbours, as a long decimal
            f = np.float32(rng.random() * 100)
            mid = (np.float64(f) + np.float64(np.nextafter(f, np.float32(np.inf)))) / 2
            texts.append("%.40f" % mid)
        elif k == 4:
            texts.append("%se%d" % (rng.integers(1, 10 ** 9), rng.integers(-60, -30)))      # binary32 subnormal range
        else:
            texts.append(["  1.5", "+2.25", "0x1.8p3", "inf", "-inf", "nan", "1.5abc", "1e", "1e+", ".", "-", "", "1e400", "-1e400", "1e-400",
                          "00012.5000", "-.5", "5.", "1E5", "1.e-2"][rng.integers(0, 20)])
    path = tmp_path / "0.xml"
    with open(path, "w") as fh:
        fh.write("<states><itno>0</itno>\n")
        for t in texts:
            fh.write("<xagent><name>Particle</name><id>1</id><x>%s</x></xagent>\n" % t)
        fh.write("</states>\n")
    got = read_states_file(_plugin("Brownian_small"), str(path))["Particle"]["x"]
    want = np.array([libc.atof(t.encode()) for t in texts], dtype=np.float64)
    with np.errstate(over="ignore"):
        want = want.astype(np.float32)
    assert len(got) == len(texts)
    same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
    assert same.all(), [(texts[i], got[i], want[i]) for i in np.flatnonzero(~same)[:10]]
