"""Pins the oracle against the known-answer vectors derived from the reference's
source and against glibc's drand48 family, which
implements the same LCG with the same seed packing (simulation.xslt:669,681)."""
import ctypes
import math
import struct

import numpy as np
import pytest

from common import W, brownian_oracle
from ev_diffusion_b200 import xmml

A48 = 0x5DEECE66D
C48 = 0xB
M48 = (1 << 48) - 1


def f32(x):
    return struct.unpack("f", struct.pack("f", x))[0]


def lcg_stream(n):
    x = (123 << 16) | 0x330E
    out = []
    for _ in range(n):
        x = (A48 * x + C48) & M48
        out.append(x)
    return out


def stride_constants(N):
    A, C = 1, 0
    for _ in range(N):
        C = (C + A * C48) & ((1 << 64) - 1)
        A = (A * A48) & ((1 << 64) - 1)
    return (A & 0xFFFFFF, (A >> 24) & 0xFFFFFF, C & 0xFFFFFF, (C >> 24) & 0xFFFFFF)


def draw(state):
    r = ((state & 0xFFFFFF) >> 17) | (((state >> 24) & 0xFFFFFF) << 7)
    return r


def test_seed_kat():
    # simulation.xslt:681-686
    s = lcg_stream(2)
    assert (s[0] & 0xFFFFFF, s[0] >> 24) == (0x3B5101, 0x478E19)
    assert (s[1] & 0xFFFFFF, s[1] >> 24) == (0xF46378, 0x6A1E93)
    assert draw(s[0]) == 600247453 and draw(s[1]) == 890194426
    assert f32(np.float32(600247453) / np.float32(2147483648.0)) == f32(0.2795119881629944)
    assert np.float32(np.float32(600247453) / np.float32(2147483648.0)).view(np.uint32) == 0x3E8F1C32


@pytest.mark.parametrize("N,A,C,second", [
    (1024, (0xDE9001, 0xBA5CC7), (0x62E400, 0xABA71B), 0.7076811790466309),
    (8192, (0xF48001, 0x10B01A), (0x172000, 0x702E82), 0.5929805636405945),
    (1048576, (0x400001, 0xFEFD7A), (0x900000, 0xC3BD0B), 0.6668528914451599),
])
def test_stride_constants_kat(N, A, C, second):
    ax, ay, cx, cy = stride_constants(N)
    assert (ax, ay) == A and (cx, cy) == C
    # slot 0, second draw = state after one strided iterate
    s0 = lcg_stream(1)[0]
    Afull = ax | (ay << 24)
    Cfull = cx | (cy << 24)
    s1 = (Afull * s0 + Cfull) & M48
    assert f32(np.float32(draw(s1)) / np.float32(2147483648.0)) == f32(second)


def test_glibc_lrand48_equivalence():
    libc = ctypes.CDLL("libc.so.6")
    libc.lrand48.restype = ctypes.c_long
    libc.srand48(123)
    got = [libc.lrand48() for _ in range(64)]
    want = [draw(s) for s in lcg_stream(64)]
    assert got == want


def test_oracle_seeds_match_kat():
    sim = brownian_oracle("c2_small")
    N = sim.model.buffer_size_max
    seeds = sim.seeds()
    stream = lcg_stream(N)
    assert seeds[0].tolist() == [0x3B5101, 0x478E19]
    assert seeds[1].tolist() == [0xF46378, 0x6A1E93]
    assert np.array_equal(seeds[:, 0], np.array([s & 0xFFFFFF for s in stream], dtype=np.uint32))
    assert np.array_equal(seeds[:, 1], np.array([s >> 24 for s in stream], dtype=np.uint32))
    assert tuple(sim.rand48_constants().tolist()) == stride_constants(N)
    sim.close()


def test_oracle_draws_follow_lrand48_stream():
    """Slot i's k-th draw is the (i + k*N + 1)-th lrand48() after srand48(123).
    The Brownian `move` script draws three numbers per agent per iteration."""
    cfg = W.BROWNIAN["c2_tiny"]
    sim = brownian_oracle("c2_tiny")
    N = sim.model.buffer_size_max
    n = 1000
    st = W.brownian_state(cfg, n)
    sim.set_agents("Particle", "default", st)
    sim.set_constant("SIGMA", 0.0)  # positions must not move: isolates the RNG
    sim.step(2)
    libc = ctypes.CDLL("libc.so.6")
    libc.lrand48.restype = ctypes.c_long
    libc.srand48(123)
    draws = [libc.lrand48() for _ in range(N * 7)]
    # after 2 iterations x 3 draws, slot i holds the state whose *next* draw is number i + 6N
    seeds = sim.seeds()
    nxt = (seeds[:, 0] >> 17) | (seeds[:, 1] << 7)
    assert nxt[:n].tolist() == [draws[i + 6 * N] for i in range(n)]
    # slots beyond the population were never advanced
    assert nxt[n:n + 8].tolist() == [draws[i] for i in range(n, n + 8)]
    sim.close()


@pytest.mark.parametrize("cells,bits", [(64, 6), (1000, 10), (125000, 17), (262144, 18), (2097152, 21), (16777216, 24)])
def test_radix_bits_kat(cells, bits):
    # simulation.xslt:282
    assert int(math.ceil(math.log(cells) / math.log(2))) == bits
    m = xmml.Message("m", [], 1, "spatial", radius=1.0, bounds_min=(0, 0, 0), bounds_max=(cells, 1, 1))
    assert m.radix_bits() == bits


def test_partition_dims_kat():
    cases = [((0, 0, 0), (8, 8, 1), 1.0, (8, 8, 1)),
             ((-0.5, -0.5, -0.5), (0.5, 0.5, 0.5), 0.1, (10, 10, 10)),
             ((-0.5, -0.5, -0.5), (0.5, 0.5, 0.5), 0.02, (50, 50, 50)),
             ((0, 0, 0), (500, 500, 500), 125, (4, 4, 4)),
             ((0, 0, 0), (500, 500, 500), 15.625, (32, 32, 32)),
             ((-0.5, -0.5, -0.5), (0.5, 0.5, 0.5), 1.0 / 128, (128, 128, 128)),
             ((-0.5, -0.5, -0.5), (0.5, 0.5, 0.5), 1.0 / 256, (256, 256, 256))]
    for lo, hi, r, want in cases:
        m = xmml.Message("m", [], 1, "spatial", radius=r, bounds_min=lo, bounds_max=hi)
        assert m.partition_dims() == want
        assert m.grid_size() == want[0] * want[1] * want[2]
