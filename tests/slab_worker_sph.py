"""One rank of the SPH slab check (launched by tests/test_gpu_slabs.py on a box with >= 2 GPUs).

BASELINE config C4 is SmoothedParticleHydrodynamics under slab decomposition: two spatially partitioned messages on
one grid (a halo exchange after each build), function conditions that move particles between state lists (host-side
counts: the synchronous slab mode), positions written by a conditional function (migration).  The model draws no
random numbers, so a P-rank run and a 1-rank run differ only by the order of float sums inside a cell: every particle,
keyed by id, must agree with a single-GPU run of the same global state within a relative tolerance."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from ev_diffusion_b200 import build, slabs  # noqa: E402
from ev_diffusion_b200.runtime import Simulation, nccl_unique_id  # noqa: E402

NAME = "SmoothedParticleHydrodynamics_exact"
AGENT = "FluidParticle"


def lattice(side=24, seed=9):
    """side^3 particles in the lower half of the box, jittered; one particle in four is static."""
    rng = np.random.default_rng(seed)
    i = np.arange(side ** 3)
    h = np.float32(0.9 / side)
    cols = {"id": i.astype(np.int32)}
    for name, idx in (("x", i % side), ("y", (i // side) % side), ("z", i // (side * side))):
        cols[name] = (idx.astype(np.float32) * h - np.float32(0.45) + rng.uniform(0, 0.3, len(i)).astype(np.float32) * h).astype(np.float32)
    cols["z"] = (cols["z"] * np.float32(0.5)).astype(np.float32)          # the script's box is half as tall in z
    for name in ("dx", "dy", "dz", "fx", "fy", "fz", "density", "pressure"):
        cols[name] = np.zeros(len(i), dtype=np.float32)
    cols["isStatic"] = (i % 4 == 0).astype(np.int32)
    return cols


def gather(sim, device):
    """{variable: global array} over both state lists of every rank, on every rank."""
    out = {}
    world = dist.get_world_size()
    for st in ("default", "static"):
        a = sim.agents(AGENT, st)
        n = len(a["id"])
        sizes = torch.zeros(world, dtype=torch.int64, device=device)
        sizes[dist.get_rank()] = n
        dist.all_reduce(sizes)
        cap = max(1, int(sizes.max().item()))
        for k, v in a.items():
            buf = torch.zeros(cap, dtype=torch.float64, device=device)
            buf[:n] = torch.from_numpy(v.astype(np.float64)).to(device)
            parts = [torch.zeros_like(buf) for _ in range(world)]
            dist.all_gather(parts, buf)
            out.setdefault(k, []).extend(p[:int(sizes[r])].cpu().numpy() for r, p in enumerate(parts))
    return {k: np.concatenate(v) for k, v in out.items()}


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    plugin = build.model_path(NAME)
    sim = Simulation(plugin, device=local)
    uid = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        uid = torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8).to(device)
    dist.broadcast(uid, 0)
    sim.slab_init(rank, world, bytes(uid.cpu().numpy().tobytes()), 65536, 65536)
    m = sim.model.messages[0]
    plan = slabs.make_plan(m.dims, m.min_bounds, m.max_bounds, m.is_2d, rank, world)
    st = lattice()
    mine = slabs.partition_state(st, plan)
    # static particles start in the `static` state list? no: the model's initial state is `default`; the function
    # conditions sort them out (computeForce / integrate run for isStatic == 0 only)
    sim.set_agents(AGENT, "default", mine)
    iters = int(os.environ.get("SLAB_ITERS", "6"))
    sim.step(iters)
    got = gather(sim, device)
    ok = True
    if rank == 0:
        ref = Simulation(plugin, device=local)
        ref.set_agents(AGENT, "default", st)
        ref.step(iters)
        want = {}
        for s_ in ("default", "static"):
            for k, v in ref.agents(AGENT, s_).items():
                want.setdefault(k, []).append(v.astype(np.float64))
        want = {k: np.concatenate(v) for k, v in want.items()}
        ref.close()
        n = len(st["id"])
        ids_g, ids_w = got["id"].astype(np.int64), want["id"].astype(np.int64)
        assert np.array_equal(np.sort(ids_g), np.arange(n)) and np.array_equal(np.sort(ids_w), np.arange(n)), "ids not conserved"
        worst = 0.0
        for k in want:
            a = np.zeros(n)
            b = np.zeros(n)
            a[ids_g] = got[k]
            b[ids_w] = want[k]
            if k in ("id", "isStatic"):
                assert np.array_equal(a, b), k
            else:
                scale = max(float(np.abs(b).max()), 1e-30)
                worst = max(worst, float(np.abs(a - b).max()) / scale)
        moved = float(np.abs(want["z"][np.argsort(ids_w)] - st["z"]).max())
        print(f"SPH_SLAB_OK world={world} iters={iters} particles={n} worst_rel={worst:.3g} max_dz={moved:.3g}")
        ok = worst <= 1e-4
    flag = torch.tensor([1 if ok else 0], device=device)
    dist.broadcast(flag, 0)
    assert int(flag.item()) == 1, "slab run differs from the single-GPU run beyond 1e-4 of the column maximum"
    sim.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
