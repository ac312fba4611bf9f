"""One rank of the slab parity check (launched by tests/test_gpu_slabs.py through
torch.distributed.run on a box with 2, 4 or 8 GPUs; the workload is c2_s<world>).

Property checked every iteration, with no oracle at this size: the neighbour count of
every agent (keyed by id) equals a brute-force count over the GLOBAL positions at the
start of that iteration -- which needs the halo planes (including the wrap pair),
the migration and the ownership rule to all be right -- and the id multiset is conserved."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from ev_diffusion_b200 import build, slabs, workloads as W  # noqa: E402
from ev_diffusion_b200.runtime import Simulation, nccl_unique_id  # noqa: E402


def gather_cols(sim, names, device):
    """All ranks' state lists on every rank (padded all_gather)."""
    a = sim.agents("Particle")
    n = len(a["id"])
    sizes = torch.zeros(dist.get_world_size(), dtype=torch.int64, device=device)
    sizes[dist.get_rank()] = n
    dist.all_reduce(sizes)
    cap = int(sizes.max().item())
    out = {}
    for k in names:
        buf = torch.zeros(cap, dtype=torch.float64, device=device)
        buf[:n] = torch.from_numpy(a[k].astype(np.float64)).to(device)
        parts = [torch.zeros_like(buf) for _ in range(dist.get_world_size())]
        dist.all_gather(parts, buf)
        out[k] = np.concatenate([p[:int(sizes[r])].cpu().numpy() for r, p in enumerate(parts)])
    owner = np.concatenate([np.full(int(sizes[r]), r) for r in range(dist.get_world_size())])
    return out, owner


def brute_counts(x, y, z):
    from scipy.spatial import cKDTree
    pts = np.stack([x, y, z], 1).astype(np.float64)
    pairs = cKDTree(pts).query_pairs(1.001, output_type="ndarray")
    xf, yf, zf = x.astype(np.float32), y.astype(np.float32), z.astype(np.float32)
    i, j = pairs[:, 0], pairs[:, 1]
    dx, dy, dz = xf[i] - xf[j], yf[i] - yf[j], zf[i] - zf[j]
    d2 = (dx * dx + dy * dy) + dz * dz            # float32, the script's evaluation order, no contraction
    ok = d2 < np.float32(1.0)
    cnt = np.zeros(len(x), dtype=np.int64)
    np.add.at(cnt, i[ok], 1)
    np.add.at(cnt, j[ok], 1)
    return cnt


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    cfg = W.BROWNIAN[f"c2_s{world}"]
    assert cfg.world == world
    sim = Simulation(build.model_path(cfg.name + "_exact"), device=local)
    for kv in filter(None, os.environ.get("SLAB_OPTIONS", "").split(",")):   # e.g. "graph=0,slab_p2p=0"
        k, v = kv.split("=")
        sim.set_option(k, int(v))
    uid = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        uid = torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8).to(device)
    dist.broadcast(uid, 0)
    sim.slab_init(rank, world, bytes(uid.cpu().numpy().tobytes()))
    m = sim.model.messages[0]
    plan = slabs.make_plan(m.dims, m.min_bounds, m.max_bounds, m.is_2d, rank, world)
    info = sim.slab_info()
    assert (info["p0"], info["p1"], info["nplanes"]) == (plan.p0, plan.p1, plan.nplanes)
    state = W.brownian_state(cfg)
    # a few agents exactly on faces / slab boundaries / the wrap boundary
    state["z"][:6] = [0.0, 15.999999, 16.0, 31.999998, 32.0, 16.000002]
    state["z"][6:8] = [np.float32(cfg.dims[2]) - np.float32(1e-5), np.float32(cfg.dims[2])]   # the wrap face itself
    sim.set_agents("Particle", "default", slabs.partition_state(state, plan))
    iters = int(os.environ.get("SLAB_ITERS", "12"))
    bad = 0
    for it in range(iters):
        before, _ = gather_cols(sim, ("id", "x", "y", "z"), device)
        sim.step(1)
        after, owner = gather_cols(sim, ("id", "x", "y", "z", "nbr"), device)
        ids_b, ids_a = before["id"].astype(np.int64), after["id"].astype(np.int64)
        assert np.array_equal(np.sort(ids_a), np.arange(cfg.n)), f"it {it}: id multiset not conserved"
        want = brute_counts(before["x"], before["y"], before["z"])
        want_by_id = np.zeros(cfg.n, dtype=np.int64)
        want_by_id[ids_b] = want
        got_by_id = np.zeros(cfg.n, dtype=np.int64)
        got_by_id[ids_a] = after["nbr"].astype(np.int64)
        bad += int((want_by_id != got_by_id).sum())
        # every agent sits on the rank that owns its plane
        for r in range(world):
            pl = slabs.make_plan(m.dims, m.min_bounds, m.max_bounds, m.is_2d, r, world)
            assert (slabs.owner_of(after["z"][owner == r], pl) == r).all(), f"it {it}: misplaced agent on rank {r}"
    moved = int((after["z"] != state["z"][ids_a]).sum())
    # List order after a migration, rank by rank, against the numpy statement of the in-place scheme
    # (slabs.migrate_in_place): function-by-function stepping so that the lists can be read between `move` and the
    # migration, which runs at the head of the next output function.
    order_bad = 0
    if "slab_inplace=0" not in os.environ.get("SLAB_OPTIONS", ""):
        names = ("id", "x", "y", "z", "nbr")
        sim.begin_iteration()
        for f in ("output_location", "count_neighbours", "move"):
            sim.run_function(f)
        for it in range(4):
            pre, own_pre = gather_cols(sim, names, device)
            sim.begin_iteration()
            sim.run_function("output_location")               # migrates first
            post, own_post = gather_cols(sim, names, device)
            lists, leave = [], []
            for r in range(world):
                pl = slabs.make_plan(m.dims, m.min_bounds, m.max_bounds, m.is_2d, r, world)
                cols = {k: pre[k][own_pre == r] for k in names}
                cls = slabs.leaver_class(cols["z"].astype(np.float32), pl)
                lists.append((pl, cols, cls))
            for r, (pl, cols, cls) in enumerate(lists):
                lo_cols, lo_cls = lists[pl.lo_rank][1], lists[pl.lo_rank][2]
                hi_cols, hi_cls = lists[pl.hi_rank][1], lists[pl.hi_rank][2]
                from_lower = {k: lo_cols[k][lo_cls == 3] for k in names}      # the lower neighbour's "up" migrants
                from_upper = {k: hi_cols[k][hi_cls == 2] for k in names}      # the upper neighbour's "down" migrants
                want = slabs.migrate_in_place(cols, cls, from_lower, from_upper)
                got = {k: post[k][own_post == r] for k in names}
                for k in ("id", "x", "y", "z"):
                    if len(want[k]) != len(got[k]) or not np.array_equal(want[k], got[k]):
                        order_bad += 1
            sim.run_function("count_neighbours")
            sim.run_function("move")
        if rank == 0:
            print(f"SLAB_ORDER mismatching lists={order_bad}")
    assert order_bad == 0, f"{order_bad} per-rank lists differ from slabs.migrate_in_place"
    # a few more iterations in one call (replayed graphs in the default mode), then a digest of the global state in
    # rank order: every transport / scheduling mode must produce the same bits
    sim.step(7)
    final, owner = gather_cols(sim, ("id", "x", "y", "z", "nbr"), device)
    import hashlib
    digest = hashlib.sha1(b"".join(final[k].astype(np.float64).tobytes() for k in ("id", "x", "y", "z", "nbr"))).hexdigest()
    if rank == 0:
        print(f"SLAB_DIGEST {digest} transport={'p2p' if sim.slab_info().get('p2p') else 'nccl'}")
    if rank == 0:
        print(f"SLAB_OK world={world} iters={iters} agents={cfg.n} mismatches={bad} moved={moved} "
              f"per_rank={[int((owner == r).sum()) for r in range(world)]}")
    assert bad == 0, f"{bad} neighbour counts differ from the brute-force count"
    sim.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
