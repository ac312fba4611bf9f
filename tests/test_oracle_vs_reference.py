"""The CPU oracle against THE REFERENCE ITSELF.

oracle/build_ref.py expands the reference's own code templates (FLAMEGPU/templates/*.xslt, with oracle/xslt_subset.py in
place of the missing xsltproc) over a model, rewrites only the <<<...>>> launch syntax, and compiles the result -- the
reference's host orchestration, its kernels and the model's functions.c -- against oracle/cuda_on_host/, which runs every
CUDA thread on the CPU.  These tests run that build and the oracle side by side from the same input and require the same
bytes in every agent variable of every state, the same per-state counts and the same rand48 seed array:

  * the reference's examples from their shipped 0.xml (read by the reference's OWN reader, io.xslt:211-621, on one side and
    by oracle/xml_states.py on the other), also against the committed golden fixtures;
  * the authored fixtures that reach what no shipped input reaches: death compaction with births
    (BirthDeath), optional messages, function conditions with state change, host-side agent creation, globalCondition;
  * the flagship workload (Brownian, C2 at parity size) against tests/golden/Brownian_small_iter10.npz.

Block size: the reference's non-partitioned message walk starts at the calling block's own tile
(FLAMEGPU_kernals.xslt:445-520), so the ORDER in which an agent meets brute-force messages depends on the block size the
occupancy API picks on the device at hand.  The oracle and the CUDA path define the order as message 0..N-1 for every
agent, which is the reference's order whenever one block holds the whole list; the runs below use CUDA_ON_HOST_BLOCK=1024
for that reason and test_brute_force_order_depends_on_block_size shows the effect at 128.  Spatially partitioned walks do
not depend on it.

Needs the reference tree (templates + vendored glm + the examples' 0.xml) to build and to read inputs: skipped where
/root/reference is not mounted, except the authored models whose prebuilt libraries travel in oracle/_ref/.
"""
import os
import subprocess
import sys

os.environ.setdefault("CUDA_ON_HOST_BLOCK", "1024")

import numpy as np
import pytest

from common import REFERENCE, ROOT, W, OracleSimulation, gen_oracle
from ev_diffusion_b200 import build
from oracle import build_ref
from oracle.ref_sim import RefSimulation
from test_oracle_golden import CASES as GOLDEN_CASES, GOLDEN, _replay

HAVE_REFERENCE = build_ref.reference_available()


def _ref_library(name, xml, function_dir=None):
    lib = build_ref.ref_path(name)
    if HAVE_REFERENCE:
        return build_ref.build_ref(xml, name, function_dir=function_dir)
    if os.path.exists(lib):
        return lib
    pytest.skip("oracle/_ref/%s is not prebuilt and the reference tree is not mounted" % name)


def _assert_same(ref, orc, label):
    for a in orc.model.agents:
        for s in a.states:
            n = orc.agent_count(a.name, s)
            assert ref.agent_count(a.name, s) == n, (label, a.name, s)
            for v in a.vars:
                got, want = ref.agent_var(a.name, s, v.name), orc.agent_var(a.name, s, v.name, n)
                assert got.dtype == want.dtype and np.array_equal(got.view(np.uint8), want.view(np.uint8)), (label, a.name, s, v.name)
    assert np.array_equal(ref.seeds(), orc.seeds()), (label, "rand48 seeds")


# name -> (iterations run side by side, compare with the golden fixture at its iteration count too)
EXAMPLES = {"CirclesPartitioning_float": (100, True), "Boids_Partitioning": (50, True), "Keratinocyte": (200, True),
            "CirclesBruteForce_float": (20, True), "CirclesBruteForce_double": (20, True), "CirclesPartitioning_double": (50, True),
            "EmptyExample": (5, True), "Analytics": (10, True), "SocialParticleSimulation": (3, False), "MonteCarlo_MSMPR": (1, False)}


@pytest.mark.skipif(not HAVE_REFERENCE, reason="the reference tree is not mounted (inputs are its examples' 0.xml)")
@pytest.mark.parametrize("name", sorted(EXAMPLES))
def test_reference_example_matches_oracle(name):
    iters, with_golden = EXAMPLES[name]
    xml = f"{REFERENCE}/examples/{name}/src/model/XMLModelFile.xml"
    ref = RefSimulation(_ref_library(name, xml), xml, f"{REFERENCE}/examples/{name}/iterations/0.xml")
    orc = _replay(name, 1, 0)
    _assert_same(ref, orc, "after reading 0.xml")
    checkpoints = sorted(k for k in {1, 2, iters // 2, iters} if 0 < k <= iters)
    done = 0
    for stop in checkpoints:
        ref.step(stop - done)
        orc.step(stop - done)
        done = stop
        _assert_same(ref, orc, f"iteration {stop}")
    assert ref.iteration == iters
    if with_golden:
        assert iters == GOLDEN_CASES[name]
        want = np.load(os.path.join(GOLDEN, f"{name}_iter{iters}.npz"))
        for a in orc.model.agents:
            for s in a.states:
                assert ref.agent_count(a.name, s) == int(want[f"count/{a.name}/{s}"])
                for v in a.vars:
                    assert np.array_equal(ref.agent_var(a.name, s, v.name).view(np.uint8), want[f"agent/{a.name}/{s}/{v.name}"].view(np.uint8))
        assert np.array_equal(ref.seeds(), want["seeds"])
    ref.close()
    orc.close()


@pytest.mark.skipif(not HAVE_REFERENCE, reason="the reference tree is not mounted")
@pytest.mark.parametrize("name,iters", [("Boids_BruteForce", 10)])
def test_further_reference_examples_without_fixtures(name, iters):
    """A shipped example outside the GPU suite and without a golden fixture: oracle and reference build are both made on
    demand and both start from the example's own 0.xml (the oracle through oracle/xml_states.py).  2048 boids walk a
    non-partitioned list in two 1024-thread blocks, so the oracle is told the block size (module docstring)."""
    from ev_diffusion_b200 import xmml
    from oracle.xml_states import read_states
    xml = f"{REFERENCE}/examples/{name}/src/model/XMLModelFile.xml"
    inp = f"{REFERENCE}/examples/{name}/iterations/0.xml"
    model = xmml.parse_model(xml)
    ref = RefSimulation(_ref_library(name, xml), xml, inp)
    orc = OracleSimulation(gen_oracle.build_oracle(xml, name))
    spec = [(a.name, [(v.name, v.type, float(v.default or 0)) for v in a.variables]) for a in model.agents]
    states, env = read_states(inp, spec, [(c.name, c.type) for c in model.constants if not c.array_length])
    for k, v in env.items():
        orc.set_constant(k, v)
    for a in model.agents:
        orc.set_agents(a.name, a.initial_state, states[a.name])
    orc.run_init_functions()
    orc.set_brute_force_block(int(os.environ["CUDA_ON_HOST_BLOCK"]))
    _assert_same(ref, orc, "after reading 0.xml")
    for it in range(iters):
        ref.step(1)
        orc.step(1)
        _assert_same(ref, orc, f"iteration {it + 1}")
    orc.set_brute_force_block(0)
    ref.close()
    orc.close()


@pytest.mark.skipif(not HAVE_REFERENCE, reason="the reference tree is not mounted")
@pytest.mark.parametrize("name,iters", [("CirclesPartitioning_float", 100), ("CirclesPartitioning_double", 10), ("Keratinocyte", 50)])
def test_device_patched_sources_compute_the_same(name, iters):
    """The sm_100a build of the reference (build_ref_device, bench.py's reference_cuda leg) differs from the reference's
    text in one thing: texture references become `__device__ const T*` read through __ldg, bound by cudaMemcpyToSymbol.
    The SAME patched sources run on the CPU must still reproduce the golden fixtures (float, int2-as-double and PBM
    textures; several messages), so the GPU baseline times the reference's computation and not a broken variant."""
    xml = f"{REFERENCE}/examples/{name}/src/model/XMLModelFile.xml"
    lib = build_ref.build_ref(xml, name, device_patch=True)
    ref = RefSimulation(lib, xml, f"{REFERENCE}/examples/{name}/iterations/0.xml")
    orc = _replay(name, 1, iters)
    ref.step(iters)
    _assert_same(ref, orc, f"iteration {iters}")
    ref.close()
    orc.close()


@pytest.mark.skipif(not HAVE_REFERENCE, reason="the reference tree is not mounted")
@pytest.mark.parametrize("fast_atomic", [True, False])
@pytest.mark.parametrize("name,iters", [("CirclesPartitioning_float", 7), ("Boids_Partitioning", 5), ("Keratinocyte", 9)])
def test_message_build_order_and_partition_matrix(name, iters, fast_atomic):
    """SURVEY.md 8(a) rows a5/a6, 8(c) "comparable bit-exactly": after the last message build of an iteration the sorted
    message list (every variable, in order) and the partition matrix equal the oracle's, for BOTH builds of the
    reference -- its default counting sort with atomics (start[] = exclusive scan, end_or_count[] = count; one host thread
    makes the atomics resolve in thread order) and the radix-sort variant with FAST_ATOMIC_SORTING removed (start[] =
    first index or 0xffffffff for an empty cell, end_or_count[] = one past the last) -- compared in normalised form:
    count per cell, start of every non-empty cell."""
    xml = f"{REFERENCE}/examples/{name}/src/model/XMLModelFile.xml"
    lib = build_ref.build_ref(xml, name, fast_atomic=fast_atomic)
    ref = RefSimulation(lib, xml, f"{REFERENCE}/examples/{name}/iterations/0.xml")
    orc = _replay(name, 1, iters)
    ref.step(iters)
    _assert_same(ref, orc, f"iteration {iters}")
    checked = 0
    for m in ref.model.messages:
        n = orc.message_count(m.name)
        assert ref.message_count(m.name) == n, m.name
        if n == 0 or m.partitioning != "spatial":
            continue
        om = orc.messages(m.name)
        for v in m.variables:
            assert np.array_equal(ref.message_var(m.name, v.name).view(np.uint8), om[v.name][:n].view(np.uint8)), (m.name, v.name)
        start, second = ref.pbm(m.name)
        count = second if fast_atomic else np.where(start.view(np.uint32) == 0xFFFFFFFF, 0, second - start)
        o_start, o_count = orc.pbm(m.name)
        assert np.array_equal(count, o_count), m.name
        assert np.array_equal(start[count > 0], o_start[o_count > 0]), m.name
        assert int(count.sum()) == n
        checked += 1
    assert checked > 0
    ref.close()
    orc.close()


def test_full_size_c2_against_the_reference_build(tmp_path):
    """BASELINE config C2 at its full size -- 1 048 576 particles, 64^3 cells -- for two iterations: the reference's own
    generated code (reading the 130 MB 0.xml with its own reader) and the oracle end in the same bytes and the same
    1 048 576 rand48 seeds.  (The CUDA path is compared with the oracle at this size in
    tests/test_gpu_brownian.py::test_full_size_c2_properties.)"""
    cfg = W.BROWNIAN["c2"]
    fdir = os.path.dirname(W.brownian_function_files()[0])
    st = W.brownian_state(cfg)
    ref = RefSimulation(_ref_library(cfg.name, W.brownian_xml(cfg), fdir), W.brownian_xml(cfg),
                        W.write_states_xml(str(tmp_path / "0.xml"), "Particle", st))
    orc = OracleSimulation(gen_oracle.oracle_path(cfg.name), threads=os.cpu_count() or 1)
    orc.set_agents("Particle", "default", st)
    _assert_same(ref, orc, "after reading 0.xml")
    ref.step(2)
    orc.step(2)
    _assert_same(ref, orc, "iteration 2")
    assert int(orc.agent_var("Particle", "default", "nbr", cfg.n).max()) > 0
    ref.close()
    orc.close()


def _write_input(path, states):
    """0.xml in the reference's layout with round-trippable %.9g floats (atof gives the same float back)."""
    with open(path, "w") as fh:
        fh.write("<states>\n<itno>0</itno>\n<environment>\n</environment>\n")
        for agent, cols in states.items():
            names = list(cols)
            for i in range(len(cols[names[0]])):
                fh.write(f"<xagent>\n<name>{agent}</name>\n")
                for k in names:
                    isf = np.issubdtype(cols[k].dtype, np.floating)
                    vals = np.atleast_1d(cols[k][i])          # an array variable: comma-separated (io.xslt:176-186)
                    fh.write(f"<{k}>" + ",".join(f"{float(v):.9g}" if isf else f"{int(v)}" for v in vals) + f"</{k}>\n")
                fh.write("</xagent>\n")
        fh.write("</states>\n")
    return str(path)


def _authored(key):
    m = build.MODELS
    if key == "Brownian_small":
        cfg = W.BROWNIAN["c2_small"]
        return cfg.name, W.brownian_xml(cfg), os.path.join(m, "Brownian", "src", "model"), {"Particle": W.brownian_state(cfg)}, 10
    if key == "BirthDeath":
        n = 5000
        u = W.uniform24(45, 3 * n)
        st = {"id": np.arange(n, dtype=np.int32), "x": u[0::3] * np.float32(32), "y": u[1::3] * np.float32(32),
              "z": u[2::3] * np.float32(32), "age": np.zeros(n, np.int32), "crowd": np.zeros(n, np.int32)}
        return key, os.path.join(m, key, "src", "model", "XMLModelFile.xml"), os.path.join(m, key, "src", "model"), {"Cell": st}, 40
    if key == "OptionalSignal":
        n, rng = 1000, np.random.default_rng(3)     # one 1024-thread block: the brute-force order is 0..N-1 (module docstring)
        st = {"id": np.arange(n, dtype=np.int32)}
        for k in "xyz":
            st[k] = (rng.random(n) * 8).astype(np.float32)
        for k in ("level", "heard", "loudest", "beacons"):
            st[k] = np.zeros(n, np.int32)
        return key, os.path.join(m, key, "src", "model", "XMLModelFile.xml"), os.path.join(m, key, "src", "model"), {"Sensor": st}, 12
    if key == "TwoStage":
        n = 4000
        st = {"id": np.arange(n, dtype=np.int32), "energy": ((np.arange(n, dtype=np.int32) * 7) % 5).astype(np.int32),
              "fired": np.zeros(n, np.int32), "heat": np.zeros(n, np.float32)}
        return key, os.path.join(m, key, "src", "model", "XMLModelFile.xml"), os.path.join(m, key, "src", "model"), {"Unit": st}, 25
    if key == "GlobalGate":
        n = 300
        ident = np.arange(n, dtype=np.int32)
        level = (ident * 5 % 4).astype(np.int32)
        level[0] = -20
        d = os.path.join(ROOT, "tests", "models", "GlobalGate")
        return key, os.path.join(d, "XMLModelFile.xml"), d, {"Node": {"id": ident, "level": level, "rounds": np.zeros(n, np.int32)}}, 40
    if key == "SmoothedParticleHydrodynamics":     # C4: no 0.xml ships for it; the seeded lattice of tests/test_gpu_sph.py
        if not HAVE_REFERENCE:
            pytest.skip("the reference tree is not mounted (the model's scripts live there)")
        from test_gpu_sph import lattice
        d = f"{REFERENCE}/examples/{key}/src/model"
        return key, f"{d}/XMLModelFile.xml", d, {"FluidParticle": lattice(12, 12, 12)}, 5
    if key == "ArrayLedger":     # agent array variables through death compaction and both appends of a state change
        n = 3000
        st = {"id": np.arange(n, dtype=np.int32), "x": W.uniform24(5, n) * np.float32(10),
              "hist": (np.arange(n * 8, dtype=np.int32).reshape(n, 8) * 3) % 11, "steps": (np.arange(n, dtype=np.int32) % 5),
              "trail": W.uniform24(6, n * 4).reshape(n, 4), "sum": np.zeros(n, np.int32)}
        return key, os.path.join(m, key, "src", "model", "XMLModelFile.xml"), os.path.join(m, key, "src", "model"), {"Walker": st}, 30
    if key == "HostCreate":      # population comes from the init function (h_add_agents_*), then 3 more per step function
        return key, os.path.join(m, key, "src", "model", "XMLModelFile.xml"), os.path.join(m, key, "src", "model"), {}, 8
    raise KeyError(key)


@pytest.mark.parametrize("key", ["Brownian_small", "BirthDeath", "OptionalSignal", "TwoStage", "GlobalGate", "HostCreate",
                                 "SmoothedParticleHydrodynamics", "ArrayLedger"])
def test_authored_model_matches_the_reference_build(key, tmp_path):
    name, xml, fdir, states, iters = _authored(key)
    model_files = [os.path.join(fdir, f) for f in __import__("ev_diffusion_b200.xmml", fromlist=["x"]).parse_model(xml).function_files]
    ref = RefSimulation(_ref_library(name, xml, fdir), xml, _write_input(tmp_path / "0.xml", states))
    orc = OracleSimulation(gen_oracle.build_oracle(xml, name, function_files=model_files))
    for agent, cols in states.items():
        a = orc.model.agents[[x.name for x in orc.model.agents].index(agent)]
        orc.set_agents(agent, a.states[a.initial_state], cols)
    orc.run_init_functions()
    _assert_same(ref, orc, "after reading 0.xml")
    for it in range(iters):
        ref.step(1)
        orc.step(1)
        _assert_same(ref, orc, f"iteration {it + 1}")
    if key == "Brownian_small":
        want = np.load(os.path.join(GOLDEN, "Brownian_small_iter10.npz"))
        for v in ("id", "x", "y", "z", "nbr"):
            assert np.array_equal(ref.agent_var("Particle", "default", v).view(np.uint8), want[f"agent/Particle/default/{v}"].view(np.uint8))
        assert np.array_equal(ref.seeds(), want["seeds"])
    if key == "HostCreate":
        assert ref.agent_count("Item", "default") == 50 + 3 * iters
    if key == "BirthDeath":
        ids = ref.agent_var("Cell", "default", "id")                 # both happened: newborns carry ids >= 2^30,
        assert (ids >= (1 << 30)).any() and len(np.setdiff1d(states["Cell"]["id"], ids)) > 0     # some founders are gone
    ref.close()
    orc.close()


_WORKER = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
from oracle.ref_sim import RefSimulation
from oracle import build_ref
name, ref_root, iters = sys.argv[2], sys.argv[3], int(sys.argv[4])
xml = f"{ref_root}/examples/{name}/src/model/XMLModelFile.xml"
sim = RefSimulation(build_ref.ref_path(name), xml, f"{ref_root}/examples/{name}/iterations/0.xml")
sim.step(iters)
out = {"seeds": sim.seeds()}
for a in sim.model.agents:
    for s in a.states:
        out[f"count/{a.name}/{s}"] = np.int32(sim.agent_count(a.name, s))
        for v in a.variables:
            out[f"agent/{a.name}/{s}/{v.name}"] = sim.agent_var(a.name, s, v.name)
np.savez(sys.argv[5], **out)
"""


def _reference_run_with_block(name, iters, block, out):
    _ref_library(name, f"{REFERENCE}/examples/{name}/src/model/XMLModelFile.xml")
    env = dict(os.environ, CUDA_ON_HOST_BLOCK=str(block))
    subprocess.run([sys.executable, "-c", _WORKER, ROOT, name, REFERENCE, str(iters), str(out)], check=True, env=env,
                   stdout=subprocess.DEVNULL)
    return np.load(out)


@pytest.mark.skipif(not HAVE_REFERENCE, reason="the reference tree is not mounted")
def test_brute_force_order_depends_on_block_size(tmp_path):
    """With 128-thread blocks the reference sums the same brute-force messages in a rotated order per block: integer
    state is unchanged, float state moves by rounding only.  (One block of 1024 reproduces the golden fixture bit for
    bit -- test_reference_example_matches_oracle.)  The block size on a real device is the occupancy API's choice, so
    bit-level agreement of float sums over non-partitioned messages is not defined by the reference; the parity tests
    of the CUDA path state a tolerance for them (tests/test_gpu_reference_models.py)."""
    name = "CirclesBruteForce_float"
    got, want = _reference_run_with_block(name, 20, 128, tmp_path / "b128.npz"), np.load(os.path.join(GOLDEN, f"{name}_iter20.npz"))
    assert np.array_equal(got["agent/Circle/default/id"], want["agent/Circle/default/id"])
    moved = 0
    for v in ("x", "y", "fx", "fy"):
        a, b = got[f"agent/Circle/default/{v}"], want[f"agent/Circle/default/{v}"]
        assert np.allclose(a, b, rtol=0, atol=2e-3), v
        moved += int((a != b).sum())
    assert moved > 0


@pytest.mark.skipif(not HAVE_REFERENCE, reason="the reference tree is not mounted")
@pytest.mark.parametrize("name,iters,block", [("CirclesBruteForce_float", 20, 128), ("CirclesBruteForce_double", 5, 96),
                                             ("Analytics", 10, 256), ("MonteCarlo_MSMPR", 1, 512)])
def test_oracle_restates_the_block_rotated_walk(name, iters, block, tmp_path):
    """The rotation itself is restated too (oracle option brute_force_block: block b of B threads starts at message
    (b*B) mod wrap and walks round, krn:445-520): with the same block size on both sides the reference build and the
    oracle agree bit for bit again -- including block sizes that do not divide the population (96) and populations
    of several blocks with deaths in the list (MonteCarlo_MSMPR)."""
    want = _reference_run_with_block(name, iters, block, tmp_path / "ref.npz")
    orc = _replay(name, 1, 0)
    orc.set_brute_force_block(block)
    orc.step(iters)
    for a in orc.model.agents:
        for s in a.states:
            n = int(want[f"count/{a.name}/{s}"])
            assert orc.agent_count(a.name, s) == n
            for v in a.vars:
                assert np.array_equal(orc.agent_var(a.name, s, v.name, n).view(np.uint8), want[f"agent/{a.name}/{s}/{v.name}"].view(np.uint8)), (a.name, s, v.name)
    assert np.array_equal(orc.seeds(), want["seeds"])
    orc.set_brute_force_block(0)
    orc.close()


@pytest.mark.skipif(not HAVE_REFERENCE, reason="the reference tree is not mounted")
@pytest.mark.parametrize("name,iters", [("Keratinocyte", 3), ("CirclesPartitioning_double", 2), ("Boids_Partitioning", 2)])
def test_iteration_file_writer_and_reader_against_the_reference(name, iters, tmp_path):
    """SURVEY.md 8(f) row 1.  The reference's OWN saveIterationData (io.xslt:124-200, compiled into the reference
    build) writes an iteration file; oracle/xml_states.py::write_states -- the restatement the CUDA path's writer is
    compared with byte for byte in tests/test_gpu_state_io.py -- must produce the same bytes from the same state
    (environment constants, several states, %f floats and doubles, ints).  Then the file goes back through the
    reference's OWN reader in a second process and through xml_states.read_states: same lists."""
    from oracle.xml_states import read_states, write_states
    xml = f"{REFERENCE}/examples/{name}/src/model/XMLModelFile.xml"
    ref = RefSimulation(_ref_library(name, xml), xml, f"{REFERENCE}/examples/{name}/iterations/0.xml")
    orc = _replay(name, 1, iters)
    ref.step(iters)
    theirs = ref.save_iteration(str(tmp_path) + os.sep, iters)
    model = ref.model                                       # the parsed XMML (names, C types, defaults)
    agents = [(a.name, [(s, [(v.name, v.type) for v in a.variables]) for s in a.states]) for a in model.agents]
    states = {(a.name, s): {v.name: orc.agent_var(a.name, s, v.name, orc.agent_count(a.name, s)) for v in a.variables}
              for a in model.agents for s in a.states}
    consts = [(c.name, c.type, orc.get_constant(c.name)) for c in model.constants]
    mine = tmp_path / "mine.xml"
    write_states(str(mine), iters, agents, states, consts)
    assert open(theirs, "rb").read() == mine.read_bytes()
    # read back: every agent returns in its initial state (the file does not record states), %f precision
    spec = [(a.name, [(v.name, v.type, float(v.default or 0)) for v in a.variables]) for a in model.agents]
    back, env = read_states(theirs, spec, [(c.name, c.type) for c in model.constants if not c.array_length])
    worker = ("import sys, numpy as np\nsys.path.insert(0, sys.argv[1])\nfrom oracle.ref_sim import RefSimulation\n"
              "sim = RefSimulation(sys.argv[2], sys.argv[3], sys.argv[4])\n"
              "np.savez(sys.argv[5], **{f'{a.name}/{v.name}': sim.agent_var(a.name, a.states.index(a.initial_state), v.name)"
              " for a in sim.model.agents for v in a.variables if not v.array_length})\n")
    out = tmp_path / "reread.npz"
    subprocess.run([sys.executable, "-c", worker, ROOT, build_ref.ref_path(name), xml, theirs, str(out)], check=True, stdout=subprocess.DEVNULL)
    reread = np.load(out)
    for a in model.agents:
        for v in a.variables:
            assert np.array_equal(reread[f"{a.name}/{v.name}"].view(np.uint8), back[a.name][v.name].view(np.uint8)), (a.name, v.name)
    ref.close()
    orc.close()


_NASTY_INPUT = """<states>
<itno>7</itno>
<!-- a comment with <xagent> inside -->
<environment>
</environment>
<xagent>
<name>Circle</name>
<id>0x1F</id>
<x>1.5e1</x>
<y> -2.25 </y>
<z>.5</z>
<fx>1e-3</fx>
<fy>7</fy>
</xagent>
<xagent>
<name>Circle</name>
<id>-12</id>
<unknown_tag>99</unknown_tag>
<x>3.999999999999</x>
<y>abc</y>
</xagent>
<xagent>
<name>NoSuchAgent</name>
<id>5</id>
<x>1</x>
</xagent>
<xagent>
<name>Circle</name>
<id>017</id>
<x>0.1</x><y>0.2</y><z>0.30000001192092896</z>
<fx>-0.0</fx>
<fy>123456789.123</fy>
</xagent>
<xagent><name>Circle</name><id>2147483647</id><x>1e39</x><y>-1e-46</y></xagent>
</states>
<xagent>
<name>Circle</name>
<id>1</id>
</xagent>
"""


@pytest.mark.skipif(not HAVE_REFERENCE, reason="the reference tree is not mounted")
def test_reader_restatement_on_awkward_input(tmp_path):
    """the reference's readInitialStates (io.xslt:211-621) and oracle/xml_states.py -- the
    restatement the CUDA path's reader is compared with in tests/test_gpu_state_io.py -- on a file that exercises
    strtol base 0 (hex, octal, negative), atof (exponents, surrounding blanks, garbage = 0, overflow to inf, underflow),
    unknown tags, a comment containing tags, an agent of an unknown type, variables left out (defaults), several tags
    on a line, and agents after </states> (ignored)."""
    from oracle.xml_states import read_states
    name = "CirclesPartitioning_float"
    xml = f"{REFERENCE}/examples/{name}/src/model/XMLModelFile.xml"
    inp = tmp_path / "0.xml"
    inp.write_text(_NASTY_INPUT)
    worker = ("import sys, numpy as np\nsys.path.insert(0, sys.argv[1])\nfrom oracle.ref_sim import RefSimulation\n"
              "sim = RefSimulation(sys.argv[2], sys.argv[3], sys.argv[4])\n"
              "np.savez(sys.argv[5], **{v.name: sim.agent_var('Circle', 'default', v.name) for v in sim.model.agents[0].variables})\n")
    out = tmp_path / "read.npz"
    subprocess.run([sys.executable, "-c", worker, ROOT, _ref_library(name, xml), xml, str(inp), str(out)], check=True, stdout=subprocess.DEVNULL)
    theirs = np.load(out)
    model = __import__("ev_diffusion_b200.xmml", fromlist=["x"]).parse_model(xml)
    spec = [(a.name, [(v.name, v.type, float(v.default or 0)) for v in a.variables]) for a in model.agents]
    mine, _ = read_states(str(inp), spec, [])
    assert len(theirs["id"]) == 4
    assert theirs["id"].tolist() == [31, -12, 15, 2147483647]
    for v in model.agents[0].variables:
        assert np.array_equal(theirs[v.name].view(np.uint8), mine["Circle"][v.name].view(np.uint8)), (v.name, theirs[v.name], mine["Circle"][v.name])


def test_array_variables_in_a_states_file_against_the_reference_reader_and_writer(tmp_path):
    """Agent array variables in 0.xml: the reference's own reader (readArrayInput, io.xslt:69-93 -- full arrays, a short
    one that keeps the defaultValue of the rest, a variable left out, empty items, hex and octal items) against the
    numpy restatement and the CUDA path's C++ parser; then the reference's own writer (io.xslt:176-186, comma-separated)
    against the numpy restatement, byte for byte."""
    from oracle.xml_states import read_states, write_states
    from ev_diffusion_b200.runtime import read_states_file
    xml = os.path.join(build.MODELS, "ArrayLedger", "src", "model", "XMLModelFile.xml")
    fdir = os.path.dirname(xml)
    text = ("<states><itno>0</itno><environment></environment>\n"
            "<xagent><name>Walker</name><id>7</id><x>1.5</x><hist>1,2,3,4,5,6,7,8</hist><steps>2</steps><trail>0.25,0.5,0.75,1</trail></xagent>\n"
            "<xagent><name>Walker</name><id>8</id><hist>9,0x10,,011</hist><trail>1e-3</trail></xagent>\n"
            "<xagent><name>Walker</name><id>9</id></xagent>\n</states>\n")
    inp = tmp_path / "0.xml"
    inp.write_text(text)
    lib = _ref_library("ArrayLedger", xml, fdir)
    worker = ("import sys, os, numpy as np\nsys.path.insert(0, sys.argv[1])\nfrom oracle.ref_sim import RefSimulation\n"
              "sim = RefSimulation(sys.argv[2], sys.argv[3], sys.argv[4])\n"
              "np.savez(sys.argv[5], **{v.name: sim.agent_var('Walker', 'active', v.name) for v in sim.model.agents[0].variables})\n"
              "sim.save_iteration(sys.argv[6], 3)\n")
    out = tmp_path / "read.npz"
    subprocess.run([sys.executable, "-c", worker, ROOT, lib, xml, str(inp), str(out), str(tmp_path) + "/"], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    theirs = np.load(out)
    spec = [("Walker", [("id", "int", 0), ("x", "float", 0), ("hist", "int", -1, 8), ("steps", "int", 0), ("trail", "float", 0, 4), ("sum", "int", 0)])]
    mine, _ = read_states(str(inp), spec, [])
    assert theirs["hist"].tolist() == [[1, 2, 3, 4, 5, 6, 7, 8], [9, 16, 9, -1, -1, -1, -1, -1], [-1] * 8]
    plugin = build.model_path("ArrayLedger_exact")
    cuda_side = read_states_file(plugin, str(inp))["Walker"] if os.path.exists(plugin) else None
    for k in theirs.files:
        assert np.array_equal(theirs[k].view(np.uint8), mine["Walker"][k].view(np.uint8)), k
        if cuda_side is not None:
            assert np.array_equal(theirs[k].view(np.uint8), cuda_side[k].view(np.uint8)), k
    variables = [("id", "int"), ("x", "float"), ("hist", "int"), ("steps", "int"), ("trail", "float"), ("sum", "int")]
    empty = {k: v[:0] for k, v in mine["Walker"].items()}
    write_states(str(tmp_path / "mine.xml"), 3, [("Walker", [("active", variables), ("resting", variables)])],
                 {("Walker", "active"): mine["Walker"], ("Walker", "resting"): empty}, [("DEATH_P", "float", np.float32(0.05))])
    assert open(tmp_path / "3.xml").read() == open(tmp_path / "mine.xml").read()
