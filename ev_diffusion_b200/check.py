"""Host-side result checks for the Brownian workloads (numpy only; no oracle, no GPU).

Used by ``bench.py`` (untimed, after the measurement) and by the slab tests: properties that hold at any size
and any number of ranks, so that a throughput figure is never the throughput of an unverified state.

  * ``neighbour_sample_check``: for a sample of agents, the neighbour count the simulation wrote (``nbr``)
    equals a brute-force float32 count over the GLOBAL positions at the start of the iteration.  It needs
    the cell hash, the sorted order, the partition boundary matrix, the 27-cell walk and -- in slab mode --
    the halo planes (including the wrap pair), the migration and the ownership rule to all be right.
  * ``owners_ok``: every agent sits on the rank that owns its cell plane.
  * ``ids_conserved``: the id multiset is {0 .. n-1}.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np


def ids_conserved(ids: np.ndarray, n: int) -> bool:
    ids = np.asarray(ids, dtype=np.int64)
    if len(ids) != n:
        return False
    seen = np.zeros(n, dtype=np.uint8)
    ok = (ids >= 0) & (ids < n)
    if not ok.all():
        return False
    seen[ids] = 1
    return bool(seen.all())


def _cells(p: np.ndarray, dim: int) -> np.ndarray:
    # a spatial index for the check only (cell size 1.0 in the Brownian workloads): any binning works as long as
    # agents closer than the interaction radius are at most one bin apart
    return np.clip(np.floor(p).astype(np.int64), 0, dim - 1)


def neighbour_sample_check(before: Dict[str, np.ndarray], after: Dict[str, np.ndarray], dims: Tuple[int, int, int],
                           sample: int = 65536, seed: int = 1, radius: float = 1.0) -> Dict[str, int]:
    """`before`: id, x, y, z of the whole population at the start of an iteration; `after`: id, nbr at its end
    (any order).  Returns the number of sampled agents compared, mismatching, and skipped because one of their
    candidate distances is within rounding of the radius (an FMA-contracted build may legitimately decide those
    the other way) or because an agent sits exactly on an upper face (the reference's hash wraps that cell)."""
    n = len(before["id"])
    x = np.asarray(before["x"], dtype=np.float32)
    y = np.asarray(before["y"], dtype=np.float32)
    z = np.asarray(before["z"], dtype=np.float32)
    ids_b = np.asarray(before["id"], dtype=np.int64)
    dx_, dy_, dz_ = dims
    cx, cy, cz = _cells(x, dx_), _cells(y, dy_), _cells(z, dz_)
    key = (cz * dy_ + cy) * dx_ + cx
    order = np.argsort(key, kind="stable")
    counts = np.bincount(key, minlength=dx_ * dy_ * dz_)
    start = np.zeros(dx_ * dy_ * dz_ + 1, dtype=np.int64)
    np.cumsum(counts, out=start[1:])
    rng = np.random.default_rng(seed)
    pick = rng.choice(n, size=min(sample, n), replace=False)
    sx, sy, sz = x[pick], y[pick], z[pick]
    scx, scy, scz = cx[pick], cy[pick], cz[pick]
    want = np.zeros(len(pick), dtype=np.int64)
    borderline = np.zeros(len(pick), dtype=bool)
    r2 = np.float32(radius * radius)
    for oz in (-1, 0, 1):
        for oy in (-1, 0, 1):
            for ox in (-1, 0, 1):
                nx, ny, nz = scx + ox, scy + oy, scz + oz
                inside = (nx >= 0) & (nx < dx_) & (ny >= 0) & (ny < dy_) & (nz >= 0) & (nz < dz_)
                c = np.where(inside, (nz * dy_ + ny) * dx_ + nx, 0)
                lo = np.where(inside, start[c], 0)
                hi = np.where(inside, start[c + 1], 0)
                m = int((hi - lo).max()) if len(lo) else 0
                for k in range(m):
                    valid = lo + k < hi
                    j = order[np.where(valid, lo + k, 0)]
                    ddx, ddy, ddz = sx - x[j], sy - y[j], sz - z[j]
                    d2 = (ddx * ddx + ddy * ddy) + ddz * ddz          # float32, the script's evaluation order
                    notself = valid & (j != pick)
                    want += (notself & (d2 < r2)).astype(np.int64)
                    borderline |= notself & (np.abs(d2 - r2) < np.float32(2e-6))
    # an agent exactly on an upper face hashes to the opposite face (krn:880-885): rare, semantics differ from
    # plain geometry, so samples near any such agent are skipped rather than modelled
    on_face = (x >= np.float32(dx_)) | (y >= np.float32(dy_)) | (z >= np.float32(dz_))
    if on_face.any():
        edge = (scx == 0) | (scx == dx_ - 1) | (scy == 0) | (scy == dy_ - 1) | (scz == 0) | (scz == dz_ - 1)
        borderline |= edge
    got_by_id = np.zeros(n, dtype=np.int64)
    got_by_id[np.asarray(after["id"], dtype=np.int64)] = np.asarray(after["nbr"], dtype=np.int64)
    got = got_by_id[ids_b[pick]]
    use = ~borderline
    return {"sampled": int(use.sum()), "mismatches": int((got[use] != want[use]).sum()), "skipped_borderline": int(borderline.sum()),
            "max_neighbours": int(want.max()) if len(want) else 0}


def owners_ok(z: np.ndarray, owner_rank: np.ndarray, dims: Tuple[int, int, int], world: int) -> bool:
    """Every agent on the rank owning its cell plane (z-slabs of dims[2] / world planes; float arithmetic of
    message_<M>_grid_position + the wrap rule of message_<M>_hash, krn:860-889)."""
    nplanes = dims[2]
    zf = np.asarray(z, dtype=np.float32)
    p = np.floor((zf - np.float32(0.0)) * np.float32(nplanes) / np.float32(nplanes)).astype(np.int64)
    p = np.where(p < 0, nplanes - 1, p)
    p = np.where(p >= nplanes, 0, p)
    return bool((np.minimum(p // (nplanes // world), world - 1) == np.asarray(owner_rank)).all())
