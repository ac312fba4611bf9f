#!/usr/bin/env python
"""Device time of the LSD radix sort for the sizes / key widths the workloads use, over the digit widths (tuning aid)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ev_diffusion_b200.runtime import load_runtime
lib = C.CDLL(os.environ["FGPU_RT_ALT"]) if os.environ.get("FGPU_RT_ALT") else load_runtime()
lib.fgpu_debug_sort_time.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
for n, bits in ((1 << 20, 18), (2162688, 19), (4194304 + 65536, 20), (8 << 20, 21), (1 << 24, 22), (1 << 22, 21), (1 << 24, 24), (1 << 19, 15)):
    row = []
    for mb in (5, 6, 7, 8, 9, 10, 11):
        ms = C.c_float(0)
        rc = lib.fgpu_debug_sort_time(n, bits, mb, 20, C.byref(ms))
        row.append("%d:%.1f" % (mb, ms.value * 1e3) if rc == 0 else "%d:err" % mb)
    print("n=%9d bits=%2d  us by max digit bits: %s" % (n, bits, "  ".join(row)), flush=True)
