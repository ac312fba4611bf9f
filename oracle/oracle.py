"""ORACLE binding.  TEST INFRASTRUCTURE ONLY -- see oracle/oracle_rt.hpp.

Loads the per-model oracle library built by ``gen_oracle.build_oracle`` and
exposes the same Python surface as ``ev_diffusion_b200.runtime.Simulation`` so
that parity tests read symmetrically.  Only tests/, __graft_entry__.smoke()
and bench.py's CPU-baseline legs import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from ev_diffusion_b200.runtime import ModelInfo, SimBase, _Model, bind_sim_api  # noqa: E402


class OracleSimulation(SimBase):
    _prefix = "orc"

    def __init__(self, oracle_so: str, threads: int = 1):
        if not os.path.exists(oracle_so):
            raise FileNotFoundError(f"{oracle_so}: build it with oracle.gen_oracle.build_oracle()")
        lib = C.CDLL(oracle_so, mode=C.RTLD_LOCAL)
        bind_sim_api(lib, "orc")
        lib.fgpu_model_get.restype = C.POINTER(_Model)
        lib.orc_sim_create.restype = C.c_void_p
        lib.orc_sim_set_threads.argtypes = [C.c_void_p, C.c_int]
        handle = lib.orc_sim_create()
        super().__init__(lib, C.c_void_p(handle), ModelInfo(lib.fgpu_model_get().contents))
        lib.orc_sim_set_threads(self._h, int(threads))

    def set_threads(self, n: int) -> None:
        self._lib.orc_sim_set_threads(self._h, int(n))

    def set_brute_force_block(self, threads_per_block: int) -> None:
        """Block size assumed for walks over non-partitioned messages (the reference starts each block at its own
        tile, FLAMEGPU_kernals.xslt:445-520); 0 = one block holds the whole list = plain list order (default)."""
        self._lib.orc_set_brute_force_block(int(threads_per_block))
