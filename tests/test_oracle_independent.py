"""A second, independent statement of the Brownian iteration -- vectorised numpy, written from the reference's
definitions (cell = floor((p-min)*dim/(max-min)) in float32, krn:860-871; jump-to-opposite-face hash,
krn:877-889; stable sort by cell with ties by message index, sim:1860-1885; 27-cell neighbourhood; rand48 with
one stream per list slot, sim:661-688 / krn:1488-1547) -- against the C++ oracle, which executes the same
functions.c through a restatement of the reference's runtime.  Two implementations that share no code must
agree bit for bit on cell keys, sorted order, partition matrix, neighbour counts, positions and rand48 seeds."""
import numpy as np
import pytest

from common import W, brownian_oracle

A48, C48, M24 = 0x5DEECE66D, 0xB, 0xFFFFFF


def leapfrog_constants(N):
    A, C = 1, 0
    for _ in range(N):
        C = (C + A * C48) & ((1 << 64) - 1)
        A = (A * A48) & ((1 << 64) - 1)
    return A & ((1 << 48) - 1), C & ((1 << 48) - 1)


def initial_states(N):
    x, out = (123 << 16) | 0x330E, np.zeros(N, dtype=np.uint64)
    for i in range(N):
        x = (A48 * x + C48) & ((1 << 48) - 1)
        out[i] = x
    return out


def rand48_draw(state, A, C):
    """rnd<>: top 31 bits of the 48-bit state as float32(r) / float32(2^31 - 1 rounded), then advance by the stride."""
    lo, hi = state & np.uint64(M24), (state >> np.uint64(24)) & np.uint64(M24)
    r = ((lo >> np.uint64(17)) | (hi << np.uint64(7))).astype(np.int64)
    u = (r.astype(np.float32) / np.float32(2147483647)).astype(np.float32)
    with np.errstate(over="ignore"):
        nxt = (state * np.uint64(A) + np.uint64(C)) & np.uint64((1 << 48) - 1)
    return u, nxt


def cell_keys(x, y, z, dims):
    g = []
    for p, d in zip((x, y, z), dims):
        num = (p.astype(np.float32) - np.float32(0.0)) * np.float32(d)
        c = np.floor(num / np.float32(d - 0.0)).astype(np.int64)          # bounds are [0, dim) with radius 1
        g.append(c)                                                        # unwrapped: the walk starts from this
    w = [np.where(c < 0, d - 1, np.where(c >= d, 0, c)) for c, d in zip(g, dims)]
    return (w[2] * dims[1] + w[1]) * dims[0] + w[0], g


def reflect(p, hi):
    p = np.where(p < 0, -p, p).astype(np.float32)
    return np.where(p > hi, (hi - (p - hi)).astype(np.float32), p).astype(np.float32)


def numpy_iteration(st, seeds, dims, sigma, A, C):
    x, y, z, ident = st["x"], st["y"], st["z"], st["id"]
    n = len(x)
    keys, (gx, gy, gz) = cell_keys(x, y, z, dims)
    order = np.argsort(keys, kind="stable")
    cells = dims[0] * dims[1] * dims[2]
    count = np.bincount(keys, minlength=cells)
    start = np.concatenate([[0], np.cumsum(count)[:-1]])
    # neighbour count: every message of the 27 cells around the UNWRAPPED cell, each neighbour wrapped on its own
    # (opposite face, krn:1037-1160), radius test in float32, script order
    nbr = np.zeros(n, dtype=np.int32)
    sx, sy, sz, sid = x[order], y[order], z[order], ident[order]
    for i in range(n):
        c = 0
        for dz in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    cx, cy, cz = gx[i] + dx, gy[i] + dy, gz[i] + dz
                    cx = dims[0] - 1 if cx < 0 else (0 if cx >= dims[0] else cx)
                    cy = dims[1] - 1 if cy < 0 else (0 if cy >= dims[1] else cy)
                    cz = dims[2] - 1 if cz < 0 else (0 if cz >= dims[2] else cz)
                    h = (cz * dims[1] + cy) * dims[0] + cx
                    lo, hi = start[h], start[h] + count[h]
                    if hi > lo:
                        ddx = x[i] - sx[lo:hi]
                        ddy = y[i] - sy[lo:hi]
                        ddz = z[i] - sz[lo:hi]
                        d2 = ((ddx * ddx).astype(np.float32) + (ddy * ddy).astype(np.float32)).astype(np.float32)
                        d2 = (d2 + (ddz * ddz).astype(np.float32)).astype(np.float32)
                        c += int(np.count_nonzero((d2 < np.float32(1.0)) & (sid[lo:hi] != ident[i])))
        nbr[i] = c
    # move: three draws per agent from its own slot's stream, reflection at the walls
    s = seeds.copy()
    ux, s = rand48_draw(s, A, C)
    uy, s = rand48_draw(s, A, C)
    uz, s = rand48_draw(s, A, C)
    sig = np.float32(sigma)
    step = lambda u: (sig * (np.float32(2.0) * u - np.float32(1.0)).astype(np.float32)).astype(np.float32)
    nx = reflect((x + step(ux)).astype(np.float32), np.float32(dims[0]))
    ny = reflect((y + step(uy)).astype(np.float32), np.float32(dims[1]))
    nz = reflect((z + step(uz)).astype(np.float32), np.float32(dims[2]))
    return {"id": ident, "x": nx, "y": ny, "z": nz, "nbr": nbr}, s, keys, order, start, count


@pytest.mark.parametrize("count", [700, 2048])
def test_cpp_oracle_equals_numpy_restatement(count):
    cfg = W.BROWNIAN["c2_tiny"]
    orc = brownian_oracle("c2_tiny")
    st = W.brownian_state(cfg, count)
    # a few agents on faces and outside the box: wrap rule and reflection
    st["x"][:4] = [0.0, 15.999999, -0.5, 16.25]
    st["z"][4:7] = [16.0, -3.0, 47.0]
    orc.set_agents("Particle", "default", st)
    N = cfg.buffer_size                      # buffer_size_MAX of the model = leapfrog stride
    A, C = leapfrog_constants(N)
    seeds = initial_states(N)[:count]
    sigma = float(orc.get_constant("SIGMA"))
    cur = {k: v.copy() for k, v in st.items()}
    for it in range(3):
        want, seeds, keys, order, start, cnt = numpy_iteration(cur, seeds, cfg.dims, sigma, A, C)
        orc.begin_iteration()
        orc.run_function("output_location")
        assert np.array_equal(orc.message_keys("location"), keys.astype(np.uint32)), it
        assert np.array_equal(orc.message_order("location"), order.astype(np.uint32)), it
        os_, oc = orc.pbm("location")
        assert np.array_equal(oc, cnt) and np.array_equal(os_[cnt > 0], start[cnt > 0]), it
        orc.run_function("count_neighbours")
        orc.run_function("move")
        got = orc.agents("Particle")
        assert np.array_equal(got["nbr"], want["nbr"]), it
        for k in "xyz":
            assert np.array_equal(got[k].view(np.uint32), want[k].view(np.uint32)), (it, k)
        so = orc.seeds()[:count]
        assert np.array_equal(so[:, 0].astype(np.uint64) | (so[:, 1].astype(np.uint64) << np.uint64(24)), seeds), it
        cur = want
    orc.close()


def test_birth_death_list_discipline_against_numpy():
    """BirthDeath (death compaction, birth into the function's own state, rand48): the list discipline restated in
    numpy from simulation.xslt:1763-1829,1955 -- survivors keep their order (stable compaction), newborns take the
    order of their parents, the new state list is [newborns..., survivors...], and a rand48 stream belongs to the
    list SLOT, not to the agent that sits in it."""
    import os
    from common import build, gen_oracle, OracleSimulation
    xml = os.path.join(build.MODELS, "BirthDeath", "src", "model", "XMLModelFile.xml")
    so = gen_oracle.oracle_path("BirthDeath")
    if not os.path.exists(so):
        so = gen_oracle.build_oracle(xml, "BirthDeath")
    orc = OracleSimulation(so)
    n0, dims = 1500, (32, 32, 32)
    u = W.uniform24(45, 3 * n0)
    st = {"id": np.arange(n0, dtype=np.int32), "x": u[0::3] * np.float32(32), "y": u[1::3] * np.float32(32),
          "z": u[2::3] * np.float32(32), "age": np.zeros(n0, np.int32), "crowd": np.zeros(n0, np.int32)}
    orc.set_agents("Cell", "default", st)
    for k, v in (("DEATH_P", 0.1), ("BIRTH_P", 0.15)):
        orc.set_constant(k, np.float32(v))
    death_p, birth_p = np.float32(0.1), np.float32(0.15)
    N = orc.model.agents[0].buffer_size      # one agent type: buffer_size_MAX
    A, C = leapfrog_constants(N)
    seeds = initial_states(N)
    cur = {k: v.copy() for k, v in st.items()}
    for it in range(6):
        n = len(cur["id"])
        keys, _ = cell_keys(cur["x"], cur["y"], cur["z"], dims)
        count = np.bincount(keys, minlength=dims[0] * dims[1] * dims[2]).reshape(dims[2], dims[1], dims[0])
        # crowd = messages in the 27 cells around the agent's UNWRAPPED grid position (krn:1037-1160: the walk adds
        # the relative cell first and wraps each neighbour on its own, so an agent outside the box visits the
        # face cell twice -- a newborn can be jittered out of bounds)
        raw = [np.floor((cur[a] * np.float32(d)) / np.float32(d)).astype(np.int64) for a, d in zip("xyz", dims)]
        wrap = lambda c, d: np.where(c < 0, d - 1, np.where(c >= d, 0, c))
        crowd_of = np.zeros(n, dtype=np.int32)
        for dz in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    crowd_of += count[wrap(raw[2] + dz, dims[2]), wrap(raw[1] + dy, dims[1]), wrap(raw[0] + dx, dims[0])].astype(np.int32)
        s = seeds[:n].copy()
        uu, s = rand48_draw(s, A, C)
        dies = uu < death_p
        births = (~dies) & (uu > np.float32(1.0) - birth_p)
        jit = []
        for _ in range(3):                     # three more draws, only in the slots that give birth
            d, s2 = rand48_draw(s, A, C)
            s = np.where(births, s2, s)
            jit.append((d - np.float32(0.5)).astype(np.float32))
        seeds[:n] = s
        age = cur["age"] + 1
        newborn = {"id": cur["id"][births] ^ np.int32(0x40000000),
                   "x": (cur["x"][births] + jit[0][births]).astype(np.float32),
                   "y": (cur["y"][births] + jit[1][births]).astype(np.float32),
                   "z": (cur["z"][births] + jit[2][births]).astype(np.float32),
                   "age": np.zeros(int(births.sum()), np.int32), "crowd": np.zeros(int(births.sum()), np.int32)}
        keep = ~dies
        survivor = {"id": cur["id"][keep], "x": cur["x"][keep], "y": cur["y"][keep], "z": cur["z"][keep],
                    "age": age[keep], "crowd": crowd_of[keep]}
        cur = {k: np.concatenate([newborn[k], survivor[k]]) for k in cur}
        orc.step(1)
        got = orc.agents("Cell")
        assert orc.agent_count("Cell", "default") == len(cur["id"]), it
        for k in cur:
            assert np.array_equal(got[k].view(np.uint32), cur[k].view(np.uint32)), (it, k)
        so_ = orc.seeds()
        assert np.array_equal(so_[:, 0].astype(np.uint64) | (so_[:, 1].astype(np.uint64) << np.uint64(24)), seeds), it
    assert len(cur["id"]) != n0
    orc.close()


def test_function_conditions_and_state_lists_against_numpy():
    """models/TwoStage: function conditions and moves between state lists, restated from simulation.xslt:1439-1528
    (the agents that meet the condition leave the current list in order, the others close ranks),
    :1590-1601/:1936-1972 (an in-place function swaps its list back; a function into another state appends its
    agents behind those already there) and krn:1320-1387 (a rand48 stream belongs to the slot of the WORKING list)."""
    import os
    from common import build, gen_oracle, OracleSimulation
    xml = os.path.join(build.MODELS, "TwoStage", "src", "model", "XMLModelFile.xml")
    so = gen_oracle.oracle_path("TwoStage")
    if not os.path.exists(so):
        so = gen_oracle.build_oracle(xml, "TwoStage")
    orc = OracleSimulation(so)
    n = 1500
    cols = {"id": np.arange(n, dtype=np.int32), "energy": (np.arange(n, dtype=np.int32) * 7) % 5,
            "fired": np.zeros(n, np.int32), "heat": np.zeros(n, np.float32)}
    orc.set_agents("Unit", "idle", cols)
    N = orc.model.agents[0].buffer_size
    A, C = leapfrog_constants(N)
    seeds = initial_states(N)
    idle = {k: v.copy() for k, v in cols.items()}
    active = {k: v[:0].copy() for k, v in cols.items()}
    take = lambda d, m: {k: v[m] for k, v in d.items()}
    cat = lambda a, b: {k: np.concatenate([a[k], b[k]]) for k in a}
    for it in range(12):
        # charge: idle -> idle, every idle unit, one draw from the stream of its slot
        k = len(idle["id"])
        u, nxt = rand48_draw(seeds[:k].copy(), A, C)
        seeds[:k] = nxt
        idle["energy"] = idle["energy"] + np.where(u < np.float32(0.5), 2, 1).astype(np.int32)
        # fire: idle units with energy >= 5 move behind the units already active
        m = idle["energy"] >= 5
        firing = take(idle, m)
        firing["fired"] = firing["fired"] + 1
        firing["heat"] = firing["energy"].astype(np.float32)
        firing["energy"] = np.zeros_like(firing["energy"])
        idle, active = take(idle, ~m), cat(active, firing)
        # burn: active -> active, in place
        active["heat"] = (active["heat"] * np.float32(0.5)).astype(np.float32)
        # cool: active units with heat < 1.5 move behind the idle ones
        m = active["heat"] < np.float32(1.5)
        cooled = take(active, m)
        cooled["heat"] = np.zeros_like(cooled["heat"])
        active, idle = take(active, ~m), cat(idle, cooled)
        orc.step(1)
        for state, want in (("idle", idle), ("active", active)):
            assert orc.agent_count("Unit", state) == len(want["id"]), (it, state)
            got = orc.agents("Unit", state)
            for key in want:
                assert np.array_equal(got[key].view(np.uint32), want[key].view(np.uint32)), (it, state, key)
        so_ = orc.seeds()
        assert np.array_equal(so_[:, 0].astype(np.uint64) | (so_[:, 1].astype(np.uint64) << np.uint64(24)), seeds), it
    assert len(active["id"]) > 0 and len(idle["id"]) > 0 and not np.array_equal(idle["id"], np.sort(idle["id"]))
    orc.close()


def test_global_condition_against_numpy():
    """tests/models/GlobalGate: a function with a globalCondition (simulation.xslt:1530-1588) runs for EVERY agent of
    its state, and only in an iteration where all of them meet the condition (mustEvaluateTo = true) -- or once the
    condition has failed maxItterations times in a row, after which the counter starts again."""
    import os
    from common import ROOT, gen_oracle, OracleSimulation
    xml = os.path.join(ROOT, "tests", "models", "GlobalGate", "XMLModelFile.xml")
    orc = OracleSimulation(gen_oracle.build_oracle(xml, "GlobalGate"))
    n = 300
    ident = np.arange(n, dtype=np.int32)
    level, rounds = (ident * 5 % 4).astype(np.int32), np.zeros(n, np.int32)
    level[0] = -20                                        # one straggler: the first release comes from the limit
    orc.set_agents("Node", "default", {"id": ident, "level": level.copy(), "rounds": rounds.copy()})
    failed, ran = 0, []
    for it in range(40):
        level = level + 1 + ident % 3
        all_true = bool((level >= 3).all())
        if not all_true and failed < 4:
            failed += 1                                   # skipped this iteration
        else:
            failed = 0
            level, rounds = (ident % 2).astype(np.int32), rounds + 1
            ran.append((it, all_true))
        orc.step(1)
        got = orc.agents("Node")
        assert np.array_equal(got["level"], level) and np.array_equal(got["rounds"], rounds), it
        assert np.array_equal(got["id"], ident)
    # both ways of getting through the gate occurred: everybody ready, and the iteration limit
    assert any(t for _, t in ran) and any(not t for _, t in ran), ran
    orc.close()
