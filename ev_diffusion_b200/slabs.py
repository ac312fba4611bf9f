"""Host-side logic of the 1-D slab decomposition.

The partition grid is cut into ``world`` slabs of whole cell planes along the
slowest hash axis (z; y for 2-D grids -- ``hash = (z*dimY + y)*dimX + x``,
``FLAMEGPU_kernals.xslt:888``).  Everything here is plain numpy so that it can be
exercised on CPU (``gloo``, world_size 2) -- the device side lives in
``csrc/fgpu_rt.cu`` (``slab_exchange_halo`` / ``slab_migrate``) and uses NCCL.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np


@dataclass
class SlabPlan:
    rank: int
    world: int
    axis: int            # 2 = z, 1 = y
    nplanes: int
    planes_per_rank: int
    plane_cells: int
    axis_min: float
    axis_max: float

    @property
    def p0(self) -> int:
        return self.rank * self.planes_per_rank

    @property
    def p1(self) -> int:
        return self.p0 + self.planes_per_rank

    @property
    def lo_rank(self) -> int:
        return (self.rank - 1) % self.world

    @property
    def hi_rank(self) -> int:
        return (self.rank + 1) % self.world

    @property
    def lo_halo_plane(self) -> int:
        """Plane received from the lower neighbour (wraps: the hash sends out-of-range
        neighbour cells to the opposite face, krn:880-885)."""
        return (self.p0 - 1) % self.nplanes

    @property
    def hi_halo_plane(self) -> int:
        return self.p1 % self.nplanes


def make_plan(dims: Tuple[int, int, int], min_bounds, max_bounds, is_2d: bool, rank: int, world: int) -> SlabPlan:
    axis = 1 if is_2d else 2
    nplanes = dims[axis]
    if nplanes % world or nplanes // world < 2:
        raise ValueError(f"{nplanes} cell planes cannot be cut into {world} slabs of >= 2 whole planes")
    plane_cells = dims[0] * dims[1] if axis == 2 else dims[0]
    return SlabPlan(rank, world, axis, nplanes, nplanes // world, plane_cells, float(min_bounds[axis]), float(max_bounds[axis]))


def plane_of(coord: np.ndarray, plan: SlabPlan) -> np.ndarray:
    """Cell plane of a coordinate: the float arithmetic of message_<M>_grid_position
    (krn:860-871) followed by the wrap rule of message_<M>_hash (krn:877-889)."""
    c = np.asarray(coord, dtype=np.float32)
    lo, hi = np.float32(plan.axis_min), np.float32(plan.axis_max)
    p = np.floor(((c - lo) * np.float32(plan.nplanes)) / (hi - lo)).astype(np.int64)
    p = np.where(p < 0, plan.nplanes - 1, p)
    p = np.where(p >= plan.nplanes, 0, p)
    return p.astype(np.int32)


def owner_of(coord: np.ndarray, plan: SlabPlan) -> np.ndarray:
    return np.minimum(plane_of(coord, plan) // plan.planes_per_rank, plan.world - 1).astype(np.int32)


def partition_state(columns: Dict[str, np.ndarray], plan: SlabPlan) -> Dict[str, np.ndarray]:
    """The agents of a global state list that this rank owns, in global list order."""
    axis_name = "z" if plan.axis == 2 else "y"
    mine = owner_of(columns[axis_name], plan) == plan.rank
    return {k: np.ascontiguousarray(v[mine]) for k, v in columns.items()}


def halo_planes_to_send(plan: SlabPlan) -> List[Tuple[int, int]]:
    """(plane, destination rank) for the two boundary planes of this slab."""
    return [(plan.p0, plan.lo_rank), (plan.p1 - 1, plan.hi_rank)]


# ---- numpy reference of the exchange protocol (used by the gloo tests) -------------------------
def pack_fixed(rows: np.ndarray, cap: int) -> np.ndarray:
    """Fixed-capacity buffer with the count in a header row, as the device path sends it."""
    rows = np.asarray(rows, dtype=np.float32).reshape(len(rows), -1)
    out = np.zeros((cap + 1, rows.shape[1]), dtype=np.float32)
    n = min(len(rows), cap)
    out[0, 0] = n
    out[0, 1] = 1.0 if len(rows) > cap else 0.0
    out[1:1 + n] = rows[:n]
    return out


def unpack_fixed(buf: np.ndarray) -> np.ndarray:
    n = int(buf[0, 0])
    if buf[0, 1] != 0:
        raise OverflowError("exchange buffer overflow on the sending rank")
    return buf[1:1 + n]


# ---- numpy reference of the in-place migration (csrc/fgpu_kernels.cuh: k_mig_count / k_mig_pack / k_mig_fill) ---------
def leaver_class(coord: np.ndarray, plan: SlabPlan) -> np.ndarray:
    """1 = stays, 2 = goes to the lower neighbour, 3 = to the upper neighbour, 4 = further than one slab away."""
    owner = owner_of(coord, plan)
    plane = plane_of(coord, plan)
    cls = np.full(len(owner), 4, dtype=np.int32)
    cls[owner == plan.rank] = 1
    if plan.lo_rank == plan.hi_rank:                       # two slabs: the same peer on both sides
        other = owner == plan.lo_rank
        cls[other & (plane < plan.p0)] = 2
        cls[other & (plane >= plan.p0)] = 3
        cls[owner == plan.rank] = 1
    else:
        cls[owner == plan.lo_rank] = 2
        cls[owner == plan.hi_rank] = 3
    return cls


def migrate_in_place(cols: Dict[str, np.ndarray], cls: np.ndarray, from_lower: Dict[str, np.ndarray],
                     from_upper: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """The list after a migration, exactly as the device path leaves it: the leavers' slots are holes (ascending);
    arrival j -- the lower neighbour's migrants first, each group in its sender's list order -- takes hole j; arrivals
    beyond the holes are appended; holes beyond the arrivals are filled, ascending, with the live agents of the tail
    that is cut off, ascending."""
    names = list(cols.keys())
    n_old = len(cols[names[0]])
    holes = np.flatnonzero((cls == 2) | (cls == 3))
    arrivals = {k: np.concatenate([from_lower[k], from_upper[k]]) for k in names}
    na, nh = len(arrivals[names[0]]), len(holes)
    n_new = n_old + na - nh
    out = {k: np.concatenate([cols[k], np.zeros(max(0, na - nh), dtype=cols[k].dtype)]) for k in names}
    m = min(na, nh)
    for k in names:
        out[k][holes[:m]] = arrivals[k][:m]
    if na >= nh:
        for k in names:
            out[k][n_old:n_old + (na - nh)] = arrivals[k][nh:]
    else:
        rest = holes[na:]                                   # the largest holes
        tail = np.arange(n_new, n_old)
        live = tail[~np.isin(tail, rest)]
        targets = rest[rest < n_new]
        assert len(live) == len(targets)
        for k in names:
            out[k][targets] = cols[k][live]
    return {k: out[k][:n_new].copy() for k in names}
