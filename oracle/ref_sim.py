"""ctypes wrapper over a library built by oracle/build_ref.py (the reference's own generated simulation on the host).

TEST INFRASTRUCTURE.  The reference keeps its state in globals, so one process can hold one loaded model at a time per
library; `initialise` may be called again after `close()`.  The reference prints progress on stdout from inside
initialise()/singleIteration(); `quiet=True` points fd 1 at /dev/null for the duration of each call."""
from __future__ import annotations

import ctypes
import os
from contextlib import contextmanager

import numpy as np

from ev_diffusion_b200 import xmml

_DTYPES = {"int": np.int32, "unsigned int": np.uint32, "uint": np.uint32, "float": np.float32, "double": np.float64,
           "long long": np.int64, "unsigned long long": np.uint64, "char": np.int8, "unsigned char": np.uint8,
           "short": np.int16, "unsigned short": np.uint16}


@contextmanager
def _silenced(enabled: bool):
    if not enabled:
        yield
        return
    import sys
    sys.stdout.flush()
    saved = os.dup(1)
    null = os.open(os.devnull, os.O_WRONLY)
    try:
        os.dup2(null, 1)
        yield
    finally:
        libc = ctypes.CDLL(None)
        libc.fflush(None)
        os.dup2(saved, 1)
        os.close(saved)
        os.close(null)


class RefSimulation:
    def __init__(self, lib_path: str, xml_model: str, input_xml: str, quiet: bool = True):
        self.model = xmml.parse_model(xml_model)
        self.lib = ctypes.CDLL(os.path.abspath(lib_path))
        self.quiet = quiet
        self.lib.ref_initialise.argtypes = [ctypes.c_char_p]
        self.lib.ref_read_var.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        self.lib.ref_read_message_var.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        self.lib.ref_read_pbm.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        self.lib.ref_seeds.argtypes = [ctypes.c_void_p]
        self.lib.ref_iteration.restype = ctypes.c_uint
        with _silenced(quiet):
            self.lib.ref_initialise(os.path.abspath(input_xml).encode())
        self._open = True

    def step(self, n: int = 1) -> None:
        with _silenced(self.quiet):
            self.lib.ref_step(int(n))

    @property
    def iteration(self) -> int:
        return int(self.lib.ref_iteration())

    def _agent(self, agent):
        return agent if isinstance(agent, int) else [a.name for a in self.model.agents].index(agent)

    def _state(self, ai, state):
        return state if isinstance(state, int) else self.model.agents[ai].states.index(state)

    def agent_count(self, agent, state=0) -> int:
        ai = self._agent(agent)
        return int(self.lib.ref_count(ai, self._state(ai, state)))

    def agent_var(self, agent, state, var) -> np.ndarray:
        ai = self._agent(agent)
        si = self._state(ai, state)
        a = self.model.agents[ai]
        vi = var if isinstance(var, int) else [v.name for v in a.variables].index(var)
        n, L = self.agent_count(ai, si), max(1, a.variables[vi].array_length)
        out = np.empty(n * L, dtype=_DTYPES[a.variables[vi].type])
        if self.lib.ref_read_var(ai, si, vi, out.ctypes.data) != 0:
            raise KeyError((agent, state, var))
        return np.ascontiguousarray(out.reshape(L, n).T) if L > 1 else out     # array variables: (count, length)

    def message_count(self, message) -> int:
        mi = message if isinstance(message, int) else [m.name for m in self.model.messages].index(message)
        return int(self.lib.ref_message_count(mi))

    def message_var(self, message, var) -> np.ndarray:
        mi = message if isinstance(message, int) else [m.name for m in self.model.messages].index(message)
        m = self.model.messages[mi]
        vi = var if isinstance(var, int) else [v.name for v in m.variables].index(var)
        out = np.empty(self.message_count(mi), dtype=_DTYPES[m.variables[vi].type])
        if self.lib.ref_read_message_var(mi, vi, out.ctypes.data) != 0:
            raise KeyError((message, var))
        return out

    def pbm(self, message):
        """(start, end_or_count) exactly as the reference stores them (deterministic variant: start = first index or
        0xffffffff for an empty cell, end_or_count = one past the last index)."""
        mi = message if isinstance(message, int) else [m.name for m in self.model.messages].index(message)
        cells = self.model.messages[mi].grid_size()
        start, end = np.empty(cells, np.int32), np.empty(cells, np.int32)
        if self.lib.ref_read_pbm(mi, start.ctypes.data, end.ctypes.data) != cells:
            raise KeyError(message)
        return start, end

    def save_iteration(self, outputpath: str, iteration: int) -> str:
        """The reference's saveIterationData (io.xslt:124-200): writes <outputpath><iteration>.xml; returns the path."""
        self.lib.ref_save_iteration.argtypes = [ctypes.c_char_p, ctypes.c_int]
        with _silenced(self.quiet):
            self.lib.ref_save_iteration(outputpath.encode(), int(iteration))
        return "%s%d.xml" % (outputpath, iteration)

    def seeds(self) -> np.ndarray:
        out = np.empty((int(self.lib.ref_buffer_size_max()), 2), dtype=np.uint32)
        self.lib.ref_seeds(out.ctypes.data)
        return out

    def close(self) -> None:
        if self._open:
            with _silenced(self.quiet):
                self.lib.ref_cleanup()
            self._open = False
