"""Shared helpers of the parity tests: build-on-demand of oracle libraries and
model plugins (both are prebuilt by __graft_entry__.build(); on the GPU box the
prebuilt files are used as they are)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from ev_diffusion_b200 import build, workloads as W  # noqa: E402
from oracle import gen_oracle  # noqa: E402
from oracle.oracle import OracleSimulation  # noqa: E402

REFERENCE = "/root/reference"


def brownian_oracle(key: str, threads: int = 1) -> OracleSimulation:
    cfg = W.BROWNIAN[key]
    so = gen_oracle.oracle_path(cfg.name)
    if not os.path.exists(so):
        so = gen_oracle.build_oracle(W.brownian_xml(cfg), cfg.name, function_files=W.brownian_function_files())
    return OracleSimulation(so, threads=threads)


def brownian_plugin(key: str, exact: bool = False) -> str:
    cfg = W.BROWNIAN[key]
    name = cfg.name + ("_exact" if exact else "")
    path = build.model_path(name)
    if not os.path.exists(path):
        path = build.build_model(W.brownian_xml(cfg), name, fmad=not exact, function_files=W.brownian_function_files())
    return path


def ulp_diff(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Distance in units in the last place between two float32 arrays."""
    ai = a.astype(np.float32).view(np.int32).astype(np.int64)
    bi = b.astype(np.float32).view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(ai & 0x7FFFFFFF), ai)
    bi = np.where(bi < 0, -(bi & 0x7FFFFFFF), bi)
    return np.abs(ai - bi)
