"""Host-side logic of the N>1 path on CPU: slab plan, ownership rule and the
fixed-capacity exchange protocol, run as 2 and as 4 ranks over gloo.  Each rank holds the
particles it owns, exchanges its two boundary planes with its neighbours (for
world_size 2 both neighbours are the same peer, and plane 0 <-> plane nplanes-1 is
the wrap pair) and counts neighbours from local + halo particles; the result must
equal the global brute-force count for every owned particle that does not sit next
to the wrap boundary (wrapped neighbours are geometrically far: they are visited by
the 27-cell walk but never within the radius)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from common import ROOT, W
from ev_diffusion_b200 import slabs

DIMS = (8, 8, 16)
N = 3000
CAP = 1024


def _global_state():
    u = W.uniform24(5, 3 * N)
    return {"id": np.arange(N, dtype=np.int32), "x": u[0::3] * np.float32(DIMS[0]), "y": u[1::3] * np.float32(DIMS[1]),
            "z": u[2::3] * np.float32(DIMS[2])}


def _count(px, py, pz, qx, qy, qz):
    """for each p: number of q within distance < 1 (float32, script evaluation order)."""
    out = np.zeros(len(px), dtype=np.int64)
    for i in range(len(px)):
        dx, dy, dz = px[i] - qx, py[i] - qy, pz[i] - qz
        out[i] = int(((dx * dx + dy * dy) + dz * dz < np.float32(1.0)).sum())
    return out


def _worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = slabs.make_plan(DIMS, (0, 0, 0), DIMS, False, rank, world)
        if world == 2:
            assert plan.lo_rank == plan.hi_rank == 1 - rank
        else:
            assert plan.lo_rank == (rank - 1) % world and plan.hi_rank == (rank + 1) % world
        st = slabs.partition_state(_global_state(), plan)
        planes = slabs.plane_of(st["z"], plan)
        assert ((planes >= plan.p0) & (planes < plan.p1)).all()
        # boundary planes -> fixed-capacity buffers (count in the header), same order as the device path:
        # sends [low plane -> lower neighbour, high plane -> upper neighbour]; receives [from upper, from lower]
        rows = lambda sel: np.stack([st["x"][sel], st["y"][sel], st["z"][sel]], 1)  # noqa: E731
        sends = [torch.from_numpy(slabs.pack_fixed(rows(planes == p), CAP)) for p, _ in slabs.halo_planes_to_send(plan)]
        recv_hi, recv_lo = torch.zeros_like(sends[0]), torch.zeros_like(sends[0])
        # tag 0: a low plane travelling to the lower neighbour (its HIGH halo); tag 1: a high plane to the upper one
        ops = [dist.isend(sends[0], plan.lo_rank, tag=0), dist.isend(sends[1], plan.hi_rank, tag=1),
               dist.irecv(recv_hi, plan.hi_rank, tag=0), dist.irecv(recv_lo, plan.lo_rank, tag=1)]
        for op in ops:
            op.wait()
        halo = np.concatenate([slabs.unpack_fixed(recv_lo.numpy()), slabs.unpack_fixed(recv_hi.numpy())])
        # the received planes are exactly my two halo planes
        hp = slabs.plane_of(halo[:, 2], plan)
        assert set(np.unique(hp)) <= {plan.lo_halo_plane, plan.hi_halo_plane}
        qx = np.concatenate([st["x"], halo[:, 0]])
        qy = np.concatenate([st["y"], halo[:, 1]])
        qz = np.concatenate([st["z"], halo[:, 2]])
        got = _count(st["x"], st["y"], st["z"], qx, qy, qz) - 1          # minus self
        g = _global_state()
        want = _count(st["x"], st["y"], st["z"], g["x"], g["y"], g["z"]) - 1
        results[rank] = (int((got != want).sum()), len(st["id"]), len(halo))
        # migration ownership: after a shift of 0.6 along z every particle's new owner is me or a neighbour
        moved = st["z"] + np.float32(0.6)
        new_owner = slabs.owner_of(moved, plan)
        assert set(np.unique(new_owner)) <= {rank, plan.hi_rank, plan.lo_rank}
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_halo_exchange_over_gloo(world):
    """world 2: both neighbours are the same peer; world 4: distinct lower / upper neighbours, wrap pair at the ends."""
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, 29633 + world, results), nprocs=world, join=True)
    assert all(results[r][0] == 0 for r in range(world)), dict(results)
    assert sum(results[r][1] for r in range(world)) == N
    assert all(results[r][2] > 0 for r in range(world))


def _mig_worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = slabs.make_plan(DIMS, (0, 0, 0), DIMS, False, rank, world)
        st = slabs.partition_state(_global_state(), plan)
        rng = np.random.default_rng(100 + rank)
        for it in range(6):
            # a random step along z, shorter than a slab; the hash wraps out-of-range planes to the opposite face
            st["z"] = (st["z"] + rng.uniform(-1.5, 1.5, len(st["z"])).astype(np.float32)).astype(np.float32)
            cls = slabs.leaver_class(st["z"], plan)
            assert (cls != 4).all()
            names = list(st.keys())
            def pack(sel):
                rows = np.stack([st[k][sel].astype(np.float64) for k in names], 1)
                buf = np.zeros((CAP + 1, len(names)))
                buf[0, 0] = len(rows)
                buf[1:1 + len(rows)] = rows
                return torch.from_numpy(buf)
            down, up = pack(cls == 2), pack(cls == 3)
            r_lo, r_hi = torch.zeros_like(down), torch.zeros_like(down)
            # my "down" migrants arrive in the lower neighbour's from-upper buffer, my "up" ones in the upper's from-lower
            ops = [dist.isend(down, plan.lo_rank, tag=0), dist.isend(up, plan.hi_rank, tag=1),
                   dist.irecv(r_hi, plan.hi_rank, tag=0), dist.irecv(r_lo, plan.lo_rank, tag=1)]
            for op in ops:
                op.wait()
            def unpack(buf):
                n = int(buf[0, 0])
                return {k: buf[1:1 + n, i].numpy().astype(st[k].dtype) for i, k in enumerate(names)}
            before = len(st["id"])
            st = slabs.migrate_in_place(st, cls, unpack(r_lo), unpack(r_hi))
            assert (slabs.owner_of(st["z"], plan) == rank).all(), f"rank {rank} it {it}: a foreign agent stayed"
            results[(rank, it)] = (sorted(st["id"].tolist()), before)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_in_place_migration_over_gloo(world):
    """The hole-filling migration (slabs.migrate_in_place, the numpy statement of k_mig_pack / k_mig_fill): over six
    random moves every agent ends on its owner, ids are conserved globally, and no slot is lost or duplicated."""
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_mig_worker, args=(world, 29653 + world, results), nprocs=world, join=True)
    for it in range(6):
        ids = sum((results[(r, it)][0] for r in range(world)), [])
        assert sorted(ids) == list(range(N)), f"iteration {it}: id multiset not conserved"


def test_in_place_migration_cases():
    """arrivals > holes (append), arrivals < holes (tail agents fill the remaining holes, holes inside the tail are
    dropped), no arrivals at all, and a leaver in the last slot."""
    ids = np.arange(10, dtype=np.int32)
    cols = {"id": ids.copy()}
    cls = np.ones(10, dtype=np.int32)
    cls[[2, 5, 9]] = [2, 3, 2]
    A = lambda *v: {"id": np.array(v, dtype=np.int32)}  # noqa: E731
    assert slabs.migrate_in_place(cols, cls, A(100, 101), A(200, 201))["id"].tolist() == [0, 1, 100, 3, 4, 101, 6, 7, 8, 200, 201]
    # one arrival for three holes: hole 2 <- 100; holes 5 and 9 remain, the list shrinks to 8: slot 9 is a hole in the
    # tail (dropped), slot 8 is live and moves into hole 5
    assert slabs.migrate_in_place(cols, cls, A(100), A())["id"].tolist() == [0, 1, 100, 3, 4, 8, 6, 7]
    # no arrival: the list shrinks to 7; tail slots 7, 8 are live (9 is a hole) and fill holes 2, 5 in ascending order
    assert slabs.migrate_in_place(cols, cls, A(), A())["id"].tolist() == [0, 1, 7, 3, 4, 8, 6]
    none = np.ones(10, dtype=np.int32)
    assert slabs.migrate_in_place(cols, none, A(), A(7))["id"].tolist() == list(range(10)) + [7]


def test_plan_and_ownership_rules():
    plan = slabs.make_plan((64, 64, 512), (0, 0, 0), (64, 64, 512), False, 3, 8)
    assert (plan.p0, plan.p1, plan.lo_rank, plan.hi_rank, plan.plane_cells) == (192, 256, 2, 4, 4096)
    first = slabs.make_plan((64, 64, 512), (0, 0, 0), (64, 64, 512), False, 0, 8)
    assert first.lo_halo_plane == 511 and first.lo_rank == 7           # the wrap pair
    last = slabs.make_plan((64, 64, 512), (0, 0, 0), (64, 64, 512), False, 7, 8)
    assert last.hi_halo_plane == 0 and last.hi_rank == 0
    z = np.array([0.0, 63.999996, 64.0, 511.99997, 512.0, -0.5, 1e9], dtype=np.float32)
    # z == zmax hashes to plane nplanes -> wraps to plane 0; negative -> last plane (krn:880-885)
    assert slabs.plane_of(z, first).tolist() == [0, 63, 64, 511, 0, 511, 0]
    assert slabs.owner_of(z, first).tolist() == [0, 0, 1, 7, 0, 7, 0]
    two_d = slabs.make_plan((8, 8, 1), (0, 0, 0), (8, 8, 1), True, 1, 2)
    assert two_d.axis == 1 and two_d.plane_cells == 8 and (two_d.p0, two_d.p1) == (4, 8)
    with pytest.raises(ValueError):
        slabs.make_plan((64, 64, 64), (0, 0, 0), (64, 64, 64), False, 0, 64)   # 1 plane per slab
    with pytest.raises(OverflowError):
        slabs.unpack_fixed(slabs.pack_fixed(np.zeros((10, 3)), 4))
