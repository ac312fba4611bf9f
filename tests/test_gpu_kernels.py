"""Kernel-level parity: the hand-written LSD radix sort and the ballot/scan
compaction against numpy restatements (stable argsort / boolean mask), at the
edge sizes of the tiling (tile = 4096 keys, compaction block = 1024)."""
import numpy as np
import pytest

from ev_diffusion_b200 import runtime

pytestmark = pytest.mark.gpu

SIZES = [1, 2, 31, 32, 33, 255, 256, 257, 4095, 4096, 4097, 8191, 8192, 12289, 100003, 1 << 20]


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("bits,digit", [(6, 8), (18, 8), (22, 8), (24, 8), (13, 5), (32, 8)])
def test_radix_sort_matches_stable_argsort(n, bits, digit):
    rng = np.random.default_rng(n * 131 + bits)
    hi = (1 << bits) - 1
    keys = rng.integers(0, hi, size=n, dtype=np.uint64, endpoint=True).astype(np.uint32)
    if n > 64:
        keys[: n // 3] = keys[0]          # long runs of equal keys exercise stability
        keys[n // 2:] &= np.uint32(0xFF)  # heavy collisions in the high digits
    ko, vo = runtime.debug_sort_pairs(keys, bits, digit)
    order = np.argsort(keys, kind="stable").astype(np.uint32)
    assert np.array_equal(vo, order)
    assert np.array_equal(ko, keys[order])


def test_radix_sort_empty():
    ko, vo = runtime.debug_sort_pairs(np.zeros(0, dtype=np.uint32), 8)
    assert ko.size == 0 and vo.size == 0


def test_radix_sort_all_equal_and_sorted_inputs():
    n = 50000
    for keys in (np.zeros(n, np.uint32), np.arange(n, dtype=np.uint32), np.arange(n, dtype=np.uint32)[::-1].copy()):
        ko, vo = runtime.debug_sort_pairs(keys, 16)
        order = np.argsort(keys, kind="stable").astype(np.uint32)
        assert np.array_equal(vo, order) and np.array_equal(ko, keys[order])


@pytest.mark.parametrize("n", [1, 31, 32, 33, 1023, 1024, 1025, 2048, 5000, 1 << 20, (1 << 20) + 7])
@pytest.mark.parametrize("p", [0.0, 0.02, 0.5, 1.0])
def test_compaction_is_stable(n, p):
    rng = np.random.default_rng(n + int(p * 100))
    flags = (rng.random(n) < p).astype(np.int32)
    if n > 4:
        flags[3] = 2   # only flag == 1 is kept (scatter_<A>_Agents tests `_scan_input == 1`, krn:247)
    payload = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
    out = runtime.debug_compact(flags, payload)
    assert np.array_equal(out, payload[flags == 1])
