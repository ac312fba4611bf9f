"""Parity of the CUDA path against the CPU oracle on the Brownian model (C2):
bit-exact integers (cell keys, sorted order, PBM, ids, neighbour counts, rand48
seeds) after every agent function; floats bit-exact for the -fmad=false build
(only + - * / there) and within a stated ULP bound for the default
contracted build."""
import numpy as np
import pytest

from common import W, brownian_oracle, brownian_plugin, ulp_diff
from ev_diffusion_b200.runtime import Simulation

pytestmark = pytest.mark.gpu


def _pair(key, exact, count=None, **opts):
    cfg = W.BROWNIAN[key]
    gpu = Simulation(brownian_plugin(key, exact=exact))
    for k, v in opts.items():
        gpu.set_option(k, v)
    orc = brownian_oracle(key)
    st = W.brownian_state(cfg, count)
    gpu.set_agents("Particle", "default", st)
    orc.set_agents("Particle", "default", st)
    return cfg, gpu, orc


def _assert_same_state(gpu, orc, exact, max_ulp=0):
    assert gpu.agent_count("Particle", "default") == orc.agent_count("Particle", "default")
    a, b = gpu.agents("Particle"), orc.agents("Particle")
    for k in ("id", "nbr"):
        assert np.array_equal(a[k], b[k]), k
    for k in ("x", "y", "z"):
        if exact:
            assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k
        else:
            assert ulp_diff(a[k], b[k]).max() <= max_ulp, (k, ulp_diff(a[k], b[k]).max())


def _assert_same_messages(gpu, orc):
    n = orc.message_count("location")
    assert gpu.message_count("location") == n
    assert np.array_equal(gpu.message_keys("location"), orc.message_keys("location"))
    assert np.array_equal(gpu.message_order("location"), orc.message_order("location"))
    gs, gc = gpu.pbm("location")
    os_, oc = orc.pbm("location")
    assert np.array_equal(gc, oc)
    assert np.array_equal(gs[oc > 0], os_[oc > 0])
    gm, om = gpu.messages("location"), orc.messages("location")
    assert np.array_equal(gm["id"], om["id"])
    for k in ("x", "y", "z"):
        assert np.array_equal(gm[k].view(np.uint32), om[k].view(np.uint32)), k


@pytest.mark.parametrize("msd", [0, 1])
@pytest.mark.parametrize("key,count", [("c2_tiny", 1000), ("c2_tiny", 4096), ("c2_small", None)])
def test_function_by_function_exact(key, count, msd):
    """Both message builds -- LSD radix sort + gather, and the MSD pass + bucket finish -- must produce the oracle's
    sorted order, records and partition boundary matrix."""
    cfg, gpu, orc = _pair(key, exact=True, count=count, msd=msd)
    for it in range(5):
        gpu.begin_iteration()
        orc.begin_iteration()
        for f in ("output_location", "count_neighbours", "move"):
            gpu.run_function(f)
            orc.run_function(f)
            if f == "output_location":
                _assert_same_messages(gpu, orc)
            _assert_same_state(gpu, orc, exact=True)
            assert np.array_equal(gpu.seeds(), orc.seeds())
    gpu.close()
    orc.close()


@pytest.mark.parametrize("order,carry", [(0, 0), (1, 0), (1, 1)])
@pytest.mark.parametrize("graph", [0, 1])
def test_many_iterations_exact(order, carry, graph):
    """Processing order (list order / cell order) and the source of a consumer's agent variables (the SoA list at
    order[t] / the producer's carried records moved into cell order by the build) never change a result."""
    cfg, gpu, orc = _pair("c2_small", exact=True, order=order, graph=graph, carry=carry)
    gpu.step(50)
    orc.step(50)
    _assert_same_state(gpu, orc, exact=True)
    assert np.array_equal(gpu.seeds(), orc.seeds())
    assert gpu.iteration == orc.iteration == 50
    gpu.close()
    orc.close()


def test_exact_build_reproduces_the_committed_fixture():
    """tests/golden/Brownian_small_iter10.npz (written by the oracle, tests/golden/make_golden.py): the CUDA path must
    land on the same bytes without the oracle in the loop."""
    import os
    from common import ROOT
    want = np.load(os.path.join(ROOT, "tests", "golden", "Brownian_small_iter10.npz"))
    gpu = Simulation(brownian_plugin("c2_small", exact=True))
    gpu.set_agents("Particle", "default", W.brownian_state(W.BROWNIAN["c2_small"]))
    gpu.step(10)
    assert gpu.agent_count("Particle", "default") == int(want["count/Particle/default"])
    got = gpu.agents("Particle")
    for v in ("id", "x", "y", "z", "nbr"):
        ref = want[f"agent/Particle/default/{v}"]
        assert got[v].dtype == ref.dtype and np.array_equal(got[v].view(np.uint8), ref.view(np.uint8)), v
    assert np.array_equal(gpu.seeds(), want["seeds"])
    gpu.close()


def test_generic_get_next_build_is_bit_identical():
    """The same functions.c compiled byte for byte (flat get_first/get_next cursor, looprewrite off): the
    loop-restructured build every other test uses and this one must both match the oracle exactly."""
    import os
    from common import build
    cfg = W.BROWNIAN["c2_tiny"]
    path = build.model_path(cfg.name + "_generic")
    if not os.path.exists(path):
        path = build.build_model(W.brownian_xml(cfg), cfg.name + "_generic", fmad=False,
                                 function_files=W.brownian_function_files(), rewrite_loops=False)
    assert "rewritten_" not in open(os.path.join(build.GEN, cfg.name + "_generic", "model_glue.cu")).read()
    assert "rewritten_" in open(os.path.join(build.GEN, cfg.name + "_exact", "model_glue.cu")).read()
    gpu = Simulation(path)
    orc = brownian_oracle("c2_tiny")
    st = W.brownian_state(cfg)
    for s_ in (gpu, orc):
        s_.set_agents("Particle", "default", st)
        s_.step(25)
    _assert_same_state(gpu, orc, exact=True)
    assert np.array_equal(gpu.seeds(), orc.seeds())
    gpu.close()
    orc.close()


def test_default_build_within_ulp_bound():
    """nvcc contracts a*b+c into FMA (-fmad=true, the reference's build default,
    tools/common.mk:132-153); the oracle is built -ffp-contract=off.  Each `move`
    evaluates 2 contracted expressions per axis, each within 1 ULP of the
    uncontracted value at the magnitude of the position; over 20 iterations the
    stated bound is 2 ULP per iteration at the position's magnitude, and integer
    state stays identical except for particles that close to a cell face."""
    cfg, gpu, orc = _pair("c2_small", exact=False)
    iters = 20
    gpu.step(iters)
    orc.step(iters)
    a, b = gpu.agents("Particle"), orc.agents("Particle")
    assert np.array_equal(a["id"], b["id"])
    for k in ("x", "y", "z"):
        assert np.abs(a[k] - b[k]).max() <= 2 * iters * 2.0 ** -20   # 2 ULP/iteration at magnitude < 16
    assert np.array_equal(gpu.seeds(), orc.seeds())
    assert (a["nbr"] != b["nbr"]).mean() < 1e-3
    gpu.close()
    orc.close()


def test_reupload_same_size_keeps_graph_and_read_agents_fills_caller_buffers():
    """A new population of the same size reuses the captured iteration (only values changed); the
    multi-column read-back fills caller-owned arrays."""
    cfg, gpu, orc = _pair("c2_tiny", exact=True)
    gpu.step(3)
    orc.step(3)
    launches_per_step = gpu.launch_count
    st = W.brownian_state(cfg)
    st["x"] = st["x"][::-1].copy()
    st["z"] = (cfg.dims[2] - st["z"]).astype(np.float32)
    for s_ in (gpu, orc):
        s_.set_agents("Particle", "default", st)
        s_.step(4)
    _assert_same_state(gpu, orc, exact=True)
    assert gpu.launch_count - launches_per_step == 4 * (launches_per_step // 3)
    n = gpu.agent_count("Particle", "default")
    out = {"x": np.full(n + 7, -1.0, np.float32), "nbr": np.full(n, -1, np.int32)}
    assert gpu.read_agents("Particle", "default", out) == n
    ref = orc.agents("Particle")
    assert np.array_equal(out["x"][:n].view(np.uint32), ref["x"].view(np.uint32)) and np.all(out["x"][n:] == -1.0)
    assert np.array_equal(out["nbr"], ref["nbr"])
    with pytest.raises(ValueError):
        gpu.read_agents("Particle", "default", {"x": np.zeros(n, np.float64)})
    gpu.close()
    orc.close()


def test_empty_population_and_zero_messages():
    cfg, gpu, orc = _pair("c2_tiny", exact=True, count=0)
    gpu.step(3)
    orc.step(3)
    assert gpu.agent_count("Particle", "default") == 0
    assert gpu.message_count("location") == 0
    assert np.array_equal(gpu.seeds(), orc.seeds())
    gpu.close()
    orc.close()


def test_clustered_and_out_of_bounds_agents():
    """All particles in one cell (maximum collisions) and far out-of-bounds positions
    (the hash jumps to the opposite face, krn:880-885)."""
    cfg = W.BROWNIAN["c2_tiny"]
    gpu = Simulation(brownian_plugin("c2_tiny", exact=True))
    orc = brownian_oracle("c2_tiny")
    n = 777
    st = W.brownian_state(cfg, n)
    st["x"][:] = 3.25
    st["y"][:] = 7.5
    st["z"][:] = 0.125
    st["x"][:5] = [-100.0, 1e6, 16.0, -0.0, 15.999999]
    st["z"][5:9] = [-3.0, 16.0, 47.9, 1e9]
    for s in (gpu, orc):
        s.set_agents("Particle", "default", st)
        s.set_constant("SIGMA", 0.0)
        s.begin_iteration()
        s.run_function("output_location")
        s.run_function("count_neighbours")
    _assert_same_messages(gpu, orc)
    _assert_same_state(gpu, orc, exact=True)
    gpu.close()
    orc.close()


def test_full_size_c2_properties():
    """BASELINE size (1M agents): properties that do not need the oracle at full
    size -- the sorted order is a permutation, the PBM sums to N and is the
    exclusive scan of its counts, ids are untouched, positions stay in the box --
    plus one oracle iteration on the same 1M input (about a second of CPU)."""
    cfg = W.BROWNIAN["c2"]
    gpu = Simulation(brownian_plugin("c2"))
    st = W.brownian_state(cfg)
    gpu.set_agents("Particle", "default", st)
    gpu.step(3)
    n = cfg.n
    order = gpu.message_order("location")
    assert np.array_equal(np.sort(order), np.arange(n, dtype=np.uint32))
    start, count = gpu.pbm("location")
    assert count.sum() == n and count.min() >= 0
    assert np.array_equal(start, np.concatenate([[0], np.cumsum(count)[:-1]]).astype(np.int32))
    a = gpu.agents("Particle")
    assert np.array_equal(a["id"], st["id"])
    for k, hi in zip("xyz", cfg.dims):
        assert a[k].min() >= 0.0 and a[k].max() <= hi
    assert 10 < a["nbr"].mean() < 22
    orc = brownian_oracle("c2")
    orc.set_agents("Particle", "default", st)
    orc.step(1)
    gpu2 = Simulation(brownian_plugin("c2"))
    gpu2.set_agents("Particle", "default", st)
    gpu2.step(1)
    assert np.array_equal(gpu2.message_order("location"), orc.message_order("location"))
    assert np.array_equal(gpu2.agent_var("Particle", "default", "nbr"), orc.agent_var("Particle", "default", "nbr"))
    assert np.array_equal(gpu2.seeds(), orc.seeds())
    for s in (gpu, gpu2, orc):
        s.close()


def test_north_star_16m_one_iteration_exact():
    """The north-star configuration itself (16 777 216 agents, 128 x 128 x 256 cells), one iteration of the bit-exact
    build against one oracle iteration on the same input: what no smaller case reaches -- 22-bit keys (an 11 + 11 bit
    MSD split; three 8-bit passes in the LSD build), a 4M-cell partition boundary matrix, 4096 sort tiles.  Message
    order, partition boundary matrix (normalised), every agent variable and every rand48 seed must be equal."""
    import os
    from common import build
    cfg = W.BROWNIAN["c2_16m"]
    plugin = build.model_path(cfg.name + "_exact")
    if not os.path.exists(plugin):
        pytest.skip("Brownian_16M_exact is not built")
    orc = brownian_oracle("c2_16m", threads=os.cpu_count() or 1)
    st = W.brownian_state(cfg)
    orc.set_agents("Particle", "default", st)
    orc.step(1)
    want_order, (os_, oc), b, want_seeds = orc.message_order("location"), orc.pbm("location"), orc.agents("Particle"), orc.seeds()
    orc.close()
    for msd in (1, 0):
        gpu = Simulation(plugin)
        gpu.set_option("msd", msd)
        gpu.set_agents("Particle", "default", st)
        gpu.step(1)
        assert np.array_equal(gpu.message_order("location"), want_order), msd
        gs, gc = gpu.pbm("location")
        assert np.array_equal(gc, oc) and np.array_equal(gs[oc > 0], os_[oc > 0]), msd
        a = gpu.agents("Particle")
        for k in a:
            assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), (msd, k)
        assert np.array_equal(gpu.seeds(), want_seeds), msd
        gpu.close()


def test_async_upload_step_readback_two_simulations_in_alternation():
    """fgpu_sim_set_agents_async / step_async / read_agents_async: two simulations driven in alternation (what
    bench.py's pipelined end-to-end leg does) end in the bytes the synchronous calls give."""
    cfg = W.BROWNIAN["c2_small"]
    st = W.brownian_state(cfg)
    ref = Simulation(brownian_plugin("c2_small", exact=True))
    for _ in range(3):                       # the rand48 streams run on across uploads: same number of rounds on every side
        ref.set_agents("Particle", "default", st)
        ref.step(1)
    want = ref.agents("Particle")
    ref.close()
    sims = [Simulation(brownian_plugin("c2_small", exact=True)) for _ in range(2)]
    outs = [{k: np.zeros(cfg.n, dtype=st[k].dtype) for k in ("x", "y", "z", "nbr")} for _ in range(2)]
    for k in range(6):
        s_, o_ = sims[k & 1], outs[k & 1]
        s_.sync()
        s_.set_agents_async("Particle", "default", st)
        s_.step_async(1)
        s_.read_agents_async("Particle", "default", o_, cfg.n)
    for s_, o_ in zip(sims, outs):
        s_.sync()
        for k in o_:
            assert np.array_equal(o_[k].view(np.uint32), want[k].view(np.uint32)), k
        s_.close()


def test_nvtx_ranges_do_not_change_results():
    """Option "nvtx": every stage of an iteration becomes an NVTX range on the enqueueing thread (the reference's
    PROFILE_SCOPED_RANGE); without a tool attached the calls are no-ops and the run is unchanged."""
    cfg, gpu, orc = _pair("c2_tiny", exact=True, nvtx=1, graph=0)
    gpu.step(6)
    orc.step(6)
    _assert_same_state(gpu, orc, exact=True)
    gpu.close()
    orc.close()
