"""Agent array variables (header.xslt:187-194, 476-492) on the CUDA path: element e of agent i lives at var[e * MAX + i] and
every mover of the runtime has to carry all elements -- death compaction (k_compact_scatter), the append of a state change
(k_append, both directions), uploads and read-backs, the XML writer / reader (comma-separated, io.xslt:69-93,176-186), the
binary snapshot and the host getter get_<A>_<S>_variable_<v>(index, element).

ArrayLedger (authored, no function condition) is pinned against the reference's own generated code on the CPU
(tests/test_oracle_vs_reference.py); here the CUDA path is compared with the oracle after EVERY agent function, bit for bit.
StableMarriage is the reference's example with int[1024] preference arrays behind function conditions: CUDA = oracle after
every iteration, and the run must end in a stable matching (a property of the algorithm, independent of either
implementation)."""
import os

import numpy as np
import pytest

from common import REFERENCE, W, build, gen_oracle, OracleSimulation
from ev_diffusion_b200.runtime import FgpuError, Simulation, read_states_file
from oracle.xml_states import read_states, write_states

XML = os.path.join(build.MODELS, "ArrayLedger", "src", "model", "XMLModelFile.xml")
SM_XML = f"{REFERENCE}/examples/StableMarriage/src/model/XMLModelFile.xml"     # compiled where it lies, prebuilt by build()


def ledger_state(n, seed=5):
    return {"id": np.arange(n, dtype=np.int32), "x": W.uniform24(seed, n) * np.float32(10),
            "hist": (np.arange(n * 8, dtype=np.int32).reshape(n, 8) * 3) % 11, "steps": (np.arange(n, dtype=np.int32) % 5),
            "trail": W.uniform24(seed + 1, n * 4).reshape(n, 4), "sum": np.zeros(n, np.int32)}


def _plugin(name, xml, **kw):
    p = build.model_path(name)
    return p if os.path.exists(p) else build.build_model(xml, name, fmad=False, **kw)


def _oracle(name, xml, **kw):
    so = gen_oracle.oracle_path(name)
    return so if os.path.exists(so) else gen_oracle.build_oracle(xml, name, **kw)


def _same(gpu, orc, label):
    for a in orc.model.agents:
        for s in a.states:
            n = orc.agent_count(a.name, s)
            assert gpu.agent_count(a.name, s) == n, (label, a.name, s)
            for v in a.vars:
                g, o = gpu.agent_var(a.name, s, v.name), orc.agent_var(a.name, s, v.name)
                assert g.shape == o.shape and np.array_equal(g.view(np.uint8), o.view(np.uint8)), (label, a.name, s, v.name)
    assert np.array_equal(gpu.seeds(), orc.seeds()), label


def test_descriptor_carries_the_array_lengths():
    """no device needed: the plugin's descriptor and the states-file reader"""
    from ev_diffusion_b200.runtime import ModelInfo, load_model
    info = ModelInfo(load_model(_plugin("ArrayLedger_exact", XML)).contents)
    assert [(v.name, v.length) for v in info.agents[0].vars] == [("id", 1), ("x", 1), ("hist", 8), ("steps", 1), ("trail", 4), ("sum", 1)]


def test_states_file_reader_parses_arrays(tmp_path):
    """fgpu_states_file_read_var (host code only) against the numpy restatement of readArrayInput: full arrays, a short
    array (the rest keeps the defaultValue, -1 for hist, with the reference's warning), a variable left out, empty items."""
    text = ("<states><itno>0</itno><environment></environment>\n"
            "<xagent><name>Walker</name><id>7</id><x>1.5</x><hist>1,2,3,4,5,6,7,8</hist><steps>2</steps><trail>0.25,0.5,0.75,1</trail></xagent>\n"
            "<xagent><name>Walker</name><id>8</id><hist>9,0x10,,011</hist><trail>1e-3</trail></xagent>\n"
            "<xagent><name>Walker</name><id>9</id></xagent>\n</states>\n")
    path = tmp_path / "0.xml"
    path.write_text(text)
    got = read_states_file(_plugin("ArrayLedger_exact", XML), str(path))["Walker"]
    spec = [("Walker", [("id", "int", 0), ("x", "float", 0), ("hist", "int", -1, 8), ("steps", "int", 0), ("trail", "float", 0, 4), ("sum", "int", 0)])]
    want, _ = read_states(str(path), spec, [])
    assert got["hist"].tolist() == [[1, 2, 3, 4, 5, 6, 7, 8], [9, 16, 9, -1, -1, -1, -1, -1], [-1] * 8]
    for k in want["Walker"]:
        assert got[k].shape == want["Walker"][k].shape and np.array_equal(got[k].view(np.uint8), want["Walker"][k].view(np.uint8)), k
    path.write_text(text.replace("<hist>1,2,3,4,5,6,7,8</hist>", "<hist>1,2,3,4,5,6,7,8,9</hist>"))
    with pytest.raises(FgpuError, match="too many items"):
        read_states_file(_plugin("ArrayLedger_exact", XML), str(path))


@pytest.mark.gpu
@pytest.mark.parametrize("n,iters", [(1, 20), (777, 40), (16000, 25)])
def test_array_ledger_every_function(n, iters):
    gpu, orc = Simulation(_plugin("ArrayLedger_exact", XML)), OracleSimulation(_oracle("ArrayLedger", XML))
    st = ledger_state(n)
    for s in (gpu, orc):
        s.set_agents("Walker", "active", st)
    _same(gpu, orc, "upload")
    for it in range(iters):
        gpu.begin_iteration()
        orc.begin_iteration()
        for f in ("record", "rest", "wake"):
            gpu.run_function(f)
            orc.run_function(f)
            _same(gpu, orc, (it, f))
    assert 0 < gpu.agent_count("Walker", "active") < n or n == 1      # some walkers died, the arrays of the others moved up
    gpu.close()
    orc.close()


@pytest.mark.gpu
def test_array_ledger_step_and_host_getter():
    gpu, orc = Simulation(_plugin("ArrayLedger_exact", XML)), OracleSimulation(_oracle("ArrayLedger", XML))
    st = ledger_state(5000, seed=9)
    for s in (gpu, orc):
        s.set_agents("Walker", "active", st)
        s.step(12)
    _same(gpu, orc, "12 iterations")
    import ctypes as C
    lib, hist = gpu._lib, gpu.agent_var("Walker", "active", "hist")
    lib.fgpu_sim_read_agent_element.restype = C.c_int
    lib.fgpu_sim_read_agent_element.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_void_p]
    for index, element in ((0, 0), (17, 3), (len(hist) - 1, 7)):
        v = C.c_int(0)
        assert lib.fgpu_sim_read_agent_element(gpu._h, 0, 0, 2, index, element, C.byref(v)) == 0 and v.value == hist[index, element]
    assert lib.fgpu_sim_read_agent_element(gpu._h, 0, 0, 2, 0, 8, C.byref(v)) != 0        # element out of range
    with pytest.raises(FgpuError, match="array"):
        gpu.reduce("Walker", "active", "hist", "sum")                                        # header.xslt:724: no reductions over arrays
    gpu.close()
    orc.close()


@pytest.mark.gpu
def test_xml_and_snapshot_round_trip_with_arrays(tmp_path):
    gpu = Simulation(_plugin("ArrayLedger_exact", XML))
    gpu.set_agents("Walker", "active", ledger_state(900, seed=21))
    gpu.step(7)
    before = {s: gpu.agents("Walker", s) for s in ("active", "resting")}
    seeds = gpu.seeds()
    # the writer against the numpy restatement of saveIterationData: byte for byte
    path = gpu.save_states(str(tmp_path) + "/", 7)
    variables = [("id", "int"), ("x", "float"), ("hist", "int"), ("steps", "int"), ("trail", "float"), ("sum", "int")]
    write_states(str(tmp_path / "want.xml"), 7, [("Walker", [("active", variables), ("resting", variables)])],
                 {("Walker", s): before[s] for s in before}, [("DEATH_P", "float", gpu.get_constant("DEATH_P"))])
    assert open(path).read() == open(tmp_path / "want.xml").read()
    # %f loses float bits; the integer arrays come back exactly, the float ones to the printed precision
    other = Simulation(_plugin("ArrayLedger_exact", XML))
    other.load_states(path)
    back = other.agents("Walker", "active")
    assert np.array_equal(back["hist"], before["active"]["hist"]) and np.array_equal(back["id"], before["active"]["id"])
    assert np.allclose(back["trail"], before["active"]["trail"], atol=1e-6)
    # the snapshot is exact, arrays included, and the run continues identically
    gpu.save_snapshot(str(tmp_path / "s.bin"))
    other.load_snapshot(str(tmp_path / "s.bin"))
    for s in before:
        got = other.agents("Walker", s)
        for k in before[s]:
            assert np.array_equal(got[k].view(np.uint8), before[s][k].view(np.uint8)), (s, k)
    assert np.array_equal(other.seeds(), seeds)
    gpu.step(5)
    other.step(5)
    for k, v in gpu.agents("Walker", "active").items():
        assert np.array_equal(other.agents("Walker", "active")[k].view(np.uint8), v.view(np.uint8)), k
    gpu.close()
    other.close()


def marriage_state(n, seed=3):
    rng = np.random.default_rng(seed)
    men = {"id": np.arange(n, dtype=np.int32), "round": np.zeros(n, np.int32), "engaged_to": np.full(n, -1, np.int32),
           "preferred_woman": np.zeros((n, 1024), np.int32)}
    women = {"id": np.arange(n, dtype=np.int32), "current_suitor": np.full(n, -1, np.int32), "current_suitor_rank": np.full(n, -1, np.int32),
             "preferred_man": np.zeros((n, 1024), np.int32)}
    for i in range(n):
        men["preferred_woman"][i, :n] = rng.permutation(n)              # [round] -> woman proposed to in that round
        women["preferred_man"][i, :n] = np.argsort(rng.permutation(n))  # [man id] -> rank (lower is better)
    return men, women


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(SM_XML) and not os.path.exists(build.model_path("StableMarriage_exact")),
                    reason="StableMarriage plugin is not prebuilt and the reference tree (its model) is not mounted")
@pytest.mark.parametrize("n", [64, 400])
def test_stable_marriage_reference_example(n):
    """examples/StableMarriage (scripts compiled where they lie): int[1024] arrays moved by two function conditions per
    iteration.  The function filter copies whole arrays here (the reference copies element 0 only, krn:172-173, and leaves
    the rest of its working list uninitialised), so the oracle of this repository is the checker and the stable-matching
    property is the independent one."""
    gpu, orc = Simulation(_plugin("StableMarriage_exact", SM_XML)), OracleSimulation(_oracle("StableMarriage", SM_XML))
    men, women = marriage_state(n)
    for s in (gpu, orc):
        s.set_agents("Man", "unengaged", men)
        s.set_agents("Woman", "default", women)
    for it in range(4 * n):
        gpu.step(1)
        orc.step(1)
        if it < 6 or it % 97 == 0:
            _same(gpu, orc, it)
        if gpu.agent_count("Man", "engaged") == n:
            break
    _same(gpu, orc, "final")
    assert gpu.agent_count("Man", "engaged") == n and gpu.agent_count("Man", "unengaged") == 0
    m, w = gpu.agents("Man", "engaged"), gpu.agents("Woman", "default")
    wife = np.empty(n, np.int64)
    wife[m["id"]] = m["engaged_to"]
    husband = np.empty(n, np.int64)
    husband[w["id"]] = w["current_suitor"]
    assert sorted(wife.tolist()) == list(range(n)) and np.array_equal(husband[wife], np.arange(n))     # a perfect matching
    rank_w = women["preferred_man"][:, :n]                       # rank_w[woman, man]
    pref_m = men["preferred_woman"][:, :n]                       # pref_m[man, round] -> woman
    rank_m = np.empty((n, n), np.int64)
    rank_m[np.arange(n)[:, None], pref_m] = np.arange(n)[None, :]   # rank_m[man, woman]
    for man in range(n):                                         # no blocking pair
        better = pref_m[man, :rank_m[man, wife[man]]]            # women he prefers to his wife
        assert (rank_w[better, husband[better]] < rank_w[better, man]).all(), man
    gpu.close()
    orc.close()
