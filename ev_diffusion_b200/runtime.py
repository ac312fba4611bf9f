"""ctypes binding of libfgpu_rt.so -- the host-side mirror of the reference's
driver-facing API (``header.xslt:531-645``: initialise, singleIteration,
get_agent_<A>_<S>_count, get_<A>_<S>_variable_<v>, set_<constant> ...).

There is no CPU path here: if the CUDA runtime library or a model plugin has
not been built, or no GPU is present, construction raises.  The oracle under
``oracle/`` is test infrastructure and is never imported from this module.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import build

_NP_TYPES = {
    "int": np.int32,
    "unsigned int": np.uint32,
    "float": np.float32,
    "double": np.float64,
    "long long int": np.int64,
    "unsigned long long int": np.uint64,
}


class FgpuError(RuntimeError):
    pass


# ---- ctypes mirrors of include/fgpu_model.h --------------------------------
class _Var(C.Structure):
    _fields_ = [("name", C.c_char_p), ("ctype", C.c_char_p), ("size", C.c_int), ("is_float", C.c_int),
                ("list_offset", C.c_size_t), ("default_value", C.c_double), ("length", C.c_int)]


class _Agent(C.Structure):
    _fields_ = [("name", C.c_char_p), ("buffer_size", C.c_int), ("nvars", C.c_int), ("vars", C.POINTER(_Var)),
                ("nstates", C.c_int), ("states", C.POINTER(C.c_char_p)), ("initial_state", C.c_int),
                ("list_bytes", C.c_size_t), ("position_offset", C.c_size_t), ("scan_offset", C.c_size_t),
                ("carry_quads", C.c_int)]


class _Message(C.Structure):
    _fields_ = [("name", C.c_char_p), ("buffer_size", C.c_int), ("nvars", C.c_int), ("vars", C.POINTER(_Var)),
                ("rec_quads", C.c_int), ("partitioning", C.c_int), ("radius", C.c_float),
                ("min_bounds", C.c_float * 3), ("max_bounds", C.c_float * 3), ("dims", C.c_int * 3),
                ("grid_size", C.c_int), ("is_2d", C.c_int), ("radix_bits", C.c_int),
                ("xvar", C.c_int), ("yvar", C.c_int), ("zvar", C.c_int)]


class _Function(C.Structure):
    _fields_ = [("name", C.c_char_p), ("agent", C.c_int), ("current_state", C.c_int), ("next_state", C.c_int),
                ("input_message", C.c_int), ("output_message", C.c_int), ("output_type", C.c_int),
                ("xagent_output", C.c_int), ("xagent_output_state", C.c_int), ("reallocate", C.c_int),
                ("swappable", C.c_int), ("rng", C.c_int), ("condition", C.c_int), ("gc_max_iterations", C.c_int),
                ("gc_must_evaluate_to", C.c_int), ("launch", C.c_void_p), ("filter", C.c_void_p),
                ("block_size", C.c_int), ("writes_position", C.c_int), ("mirror_producer", C.c_int)]


class _Layer(C.Structure):
    _fields_ = [("nfunctions", C.c_int), ("functions", C.POINTER(C.c_int))]


class _Constant(C.Structure):
    _fields_ = [("name", C.c_char_p), ("ctype", C.c_char_p), ("size", C.c_int), ("length", C.c_int),
                ("set", C.c_void_p), ("get", C.c_void_p)]


class _Model(C.Structure):
    _fields_ = [("abi_version", C.c_int), ("name", C.c_char_p), ("buffer_size_max", C.c_int),
                ("nagents", C.c_int), ("agents", C.POINTER(_Agent)),
                ("nmessages", C.c_int), ("messages", C.POINTER(_Message)),
                ("nfunctions", C.c_int), ("functions", C.POINTER(_Function)),
                ("nlayers", C.c_int), ("layers", C.POINTER(_Layer)),
                ("nconstants", C.c_int), ("constants", C.POINTER(_Constant)),
                ("init_constants", C.c_void_p),
                ("ninit_functions", C.c_int), ("init_functions", C.c_void_p),
                ("nstep_functions", C.c_int), ("step_functions", C.c_void_p),
                ("nexit_functions", C.c_int), ("exit_functions", C.c_void_p), ("bind", C.c_void_p),
                ("init_function_names", C.c_void_p), ("step_function_names", C.c_void_p), ("exit_function_names", C.c_void_p),
                ("ids_after_load", C.c_void_p), ("ids_to_host", C.c_void_p), ("ids_to_device", C.c_void_p)]


# ---- description objects ----------------------------------------------------
class VarInfo:
    def __init__(self, v: _Var):
        self.name = v.name.decode()
        self.ctype = v.ctype.decode()
        self.size = v.size
        self.dtype = np.dtype(_NP_TYPES[self.ctype])
        self.default = v.default_value
        self.length = max(1, int(v.length))   # > 1: an agent array variable; its columns are (count, length) arrays here


class AgentInfo:
    def __init__(self, a: _Agent):
        self.name = a.name.decode()
        self.buffer_size = a.buffer_size
        self.vars = [VarInfo(a.vars[i]) for i in range(a.nvars)]
        self.states = [a.states[i].decode() for i in range(a.nstates)]
        self.initial_state = a.initial_state

    def var_index(self, name: str) -> int:
        return [v.name for v in self.vars].index(name)


class MessageInfo:
    def __init__(self, m: _Message):
        self.name = m.name.decode()
        self.buffer_size = m.buffer_size
        self.vars = [VarInfo(m.vars[i]) for i in range(m.nvars)]
        self.partitioning = m.partitioning
        self.radius = m.radius
        self.min_bounds = tuple(m.min_bounds)
        self.max_bounds = tuple(m.max_bounds)
        self.dims = tuple(m.dims)
        self.grid_size = m.grid_size
        self.is_2d = bool(m.is_2d)
        self.radix_bits = m.radix_bits

    def var_index(self, name: str) -> int:
        return [v.name for v in self.vars].index(name)


class FunctionInfo:
    def __init__(self, f: _Function):
        self.name = f.name.decode()
        self.agent = f.agent
        self.current_state = f.current_state
        self.next_state = f.next_state
        self.input_message = f.input_message
        self.output_message = f.output_message
        self.reallocate = bool(f.reallocate)
        self.swappable = bool(f.swappable)
        self.rng = bool(f.rng)
        self.condition = f.condition
        self.xagent_output = f.xagent_output


class ModelInfo:
    """Python view of an ``fgpu_model`` descriptor (shared by the CUDA runtime
    binding and, in tests, by the oracle binding)."""

    def __init__(self, m: _Model):
        self.name = m.name.decode()
        self.buffer_size_max = m.buffer_size_max
        self.agents = [AgentInfo(m.agents[i]) for i in range(m.nagents)]
        self.messages = [MessageInfo(m.messages[i]) for i in range(m.nmessages)]
        self.functions = [FunctionInfo(m.functions[i]) for i in range(m.nfunctions)]
        self.layers = [[m.layers[i].functions[k] for k in range(m.layers[i].nfunctions)] for i in range(m.nlayers)]
        self.constants = [(m.constants[i].name.decode(), m.constants[i].ctype.decode(), m.constants[i].length)
                          for i in range(m.nconstants)]

    def agent_index(self, name: str) -> int:
        return [a.name for a in self.agents].index(name)

    def message_index(self, name: str) -> int:
        return [m.name for m in self.messages].index(name)

    def function_index(self, name: str) -> int:
        return [f.name for f in self.functions].index(name)


# ---- generic binding over a C ABI with a given symbol prefix -----------------
def bind_sim_api(lib: C.CDLL, prefix: str) -> None:
    """Declare argtypes/restypes of the `<prefix>_sim_*` entry points (the oracle
    library exports the same shapes under the prefix ``orc``)."""
    vp, ci, cp = C.c_void_p, C.c_int, C.c_char_p

    def sig(name, res, *args):
        fn = getattr(lib, f"{prefix}_{name}")
        fn.restype = res
        fn.argtypes = list(args)

    sig("last_error", cp)
    sig("sim_destroy", None, vp)
    sig("sim_load_states", ci, vp, cp)
    sig("sim_set_agents", ci, vp, ci, ci, ci, C.POINTER(vp))
    sig("sim_add_agents", ci, vp, ci, ci, ci, C.POINTER(vp))
    sig("sim_run_init_functions", ci, vp)
    sig("sim_step", ci, vp, ci)
    sig("sim_begin_iteration", ci, vp)
    sig("sim_run_function", ci, vp, ci)
    sig("sim_iteration", C.c_uint, vp)
    sig("sim_agent_count", ci, vp, ci, ci)
    sig("sim_read_agent_var", ci, vp, ci, ci, ci, vp, ci)
    sig("sim_reduce_agent_var", ci, vp, ci, ci, ci, ci, vp, vp)
    sig("sim_message_count", ci, vp, ci)
    sig("sim_read_message_var", ci, vp, ci, ci, vp, ci)
    sig("sim_read_pbm", ci, vp, ci, vp, vp)
    sig("sim_read_message_keys", ci, vp, ci, vp, ci)
    sig("sim_read_message_order", ci, vp, ci, vp, ci)
    sig("sim_read_seeds", ci, vp, vp, ci)
    sig("sim_rand48_constants", ci, vp, vp)
    sig("sim_set_constant", ci, vp, cp, vp)
    sig("sim_get_constant", ci, vp, cp, vp)


class SimBase:
    """Operations common to every `<prefix>_sim_*` ABI."""

    _prefix = "fgpu"

    def __init__(self, lib: C.CDLL, handle, info: ModelInfo):
        self._lib = lib
        self._h = handle
        self.model = info

    def _fn(self, name):
        return getattr(self._lib, f"{self._prefix}_{name}")

    def _check(self, rc: int) -> int:
        if rc < 0:
            raise FgpuError(self._fn("last_error")().decode())
        return rc

    # -- lifecycle
    def close(self) -> None:
        if self._h:
            self._fn("sim_destroy")(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- state
    def _columns(self, agent: int, columns: Dict[str, np.ndarray], count: Optional[int]):
        a = self.model.agents[agent]
        arrs: List[Optional[np.ndarray]] = []
        n = count
        for v in a.vars:
            if v.name in columns and columns[v.name] is not None:
                arr = np.ascontiguousarray(columns[v.name], dtype=v.dtype)
                if n is None:
                    n = arr.shape[0]
                if arr.shape[0] != n:
                    raise FgpuError(f"column {v.name}: {arr.shape[0]} rows, expected {n}")
                if v.length > 1:   # (count, length) here; the C ABI takes `length` blocks of `count` values
                    if arr.shape != (n, v.length):
                        raise FgpuError(f"column {v.name}: shape {arr.shape}, expected ({n}, {v.length})")
                    arr = np.ascontiguousarray(arr.T)
                arrs.append(arr)
            else:
                arrs.append(None)
        n = 0 if n is None else int(n)
        ptrs = (C.c_void_p * len(arrs))(*[(x.ctypes.data if x is not None else None) for x in arrs])
        return n, ptrs, arrs

    def set_agents(self, agent, state, columns: Dict[str, np.ndarray], count: Optional[int] = None) -> None:
        agent = self._agent(agent)
        state = self._state(agent, state)
        n, ptrs, keep = self._columns(agent, columns, count)
        self._check(self._fn("sim_set_agents")(self._h, agent, state, n, ptrs))

    def add_agents(self, agent, state, columns: Dict[str, np.ndarray], count: Optional[int] = None) -> None:
        agent = self._agent(agent)
        state = self._state(agent, state)
        n, ptrs, keep = self._columns(agent, columns, count)
        self._check(self._fn("sim_add_agents")(self._h, agent, state, n, ptrs))

    def load_states(self, xml_path: str) -> None:
        self._check(self._fn("sim_load_states")(self._h, os.fsencode(xml_path)))

    def run_init_functions(self) -> None:
        self._check(self._fn("sim_run_init_functions")(self._h))

    # -- stepping
    def step(self, n: int = 1) -> None:
        self._check(self._fn("sim_step")(self._h, int(n)))

    def begin_iteration(self) -> None:
        self._check(self._fn("sim_begin_iteration")(self._h))

    def run_function(self, function) -> None:
        if isinstance(function, str):
            function = self.model.function_index(function)
        self._check(self._fn("sim_run_function")(self._h, int(function)))

    @property
    def iteration(self) -> int:
        return int(self._fn("sim_iteration")(self._h))

    # -- queries
    def _agent(self, agent) -> int:
        return self.model.agent_index(agent) if isinstance(agent, str) else int(agent)

    def _state(self, agent: int, state) -> int:
        return self.model.agents[agent].states.index(state) if isinstance(state, str) else int(state)

    def _message(self, message) -> int:
        return self.model.message_index(message) if isinstance(message, str) else int(message)

    def agent_count(self, agent, state=0) -> int:
        agent = self._agent(agent)
        return self._check(self._fn("sim_agent_count")(self._h, agent, self._state(agent, state)))

    def agent_var(self, agent, state, var, count: Optional[int] = None) -> np.ndarray:
        agent = self._agent(agent)
        state = self._state(agent, state)
        a = self.model.agents[agent]
        vi = a.var_index(var) if isinstance(var, str) else int(var)
        if count is None:
            count = self.agent_count(agent, state)
        L = a.vars[vi].length
        out = np.empty(count * L, dtype=a.vars[vi].dtype)
        self._check(self._fn("sim_read_agent_var")(self._h, agent, state, vi, out.ctypes.data, count))
        return np.ascontiguousarray(out.reshape(L, count).T) if L > 1 else out

    def agents(self, agent, state=0) -> Dict[str, np.ndarray]:
        agent = self._agent(agent)
        return {v.name: self.agent_var(agent, state, v.name) for v in self.model.agents[agent].vars}

    def message_count(self, message) -> int:
        return self._check(self._fn("sim_message_count")(self._h, self._message(message)))

    def message_var(self, message, var) -> np.ndarray:
        message = self._message(message)
        m = self.model.messages[message]
        vi = m.var_index(var) if isinstance(var, str) else int(var)
        n = self.message_count(message)
        out = np.empty(n, dtype=m.vars[vi].dtype)
        self._check(self._fn("sim_read_message_var")(self._h, message, vi, out.ctypes.data, n))
        return out

    def messages(self, message) -> Dict[str, np.ndarray]:
        message = self._message(message)
        return {v.name: self.message_var(message, v.name) for v in self.model.messages[message].vars}

    def pbm(self, message):
        """(start, count) per cell, normalised form."""
        message = self._message(message)
        g = self.model.messages[message].grid_size
        start = np.zeros(g, dtype=np.int32)
        count = np.zeros(g, dtype=np.int32)
        self._check(self._fn("sim_read_pbm")(self._h, message, start.ctypes.data, count.ctypes.data))
        return start, count

    def message_keys(self, message) -> np.ndarray:
        message = self._message(message)
        n = self.message_count(message)
        out = np.empty(n, dtype=np.uint32)
        self._check(self._fn("sim_read_message_keys")(self._h, message, out.ctypes.data, n))
        return out

    def message_order(self, message) -> np.ndarray:
        message = self._message(message)
        n = self.message_count(message)
        out = np.empty(n, dtype=np.uint32)
        self._check(self._fn("sim_read_message_order")(self._h, message, out.ctypes.data, n))
        return out

    def seeds(self, count: Optional[int] = None) -> np.ndarray:
        if count is None:
            count = self.model.buffer_size_max
        out = np.empty((count, 2), dtype=np.uint32)
        self._check(self._fn("sim_read_seeds")(self._h, out.ctypes.data, count))
        return out

    def rand48_constants(self) -> np.ndarray:
        out = np.empty(4, dtype=np.uint32)
        self._check(self._fn("sim_rand48_constants")(self._h, out.ctypes.data))
        return out

    REDUCE_OPS = {"sum": 0, "min": 1, "max": 2, "count": 3}

    def reduce(self, agent, state, var, op: str, count_value=None):
        """reduce_/min_/max_/count_<A>_<S>_<v>_variable(): one agent variable folded over a state list;
        the result has the variable's own type."""
        agent = self._agent(agent)
        state = self._state(agent, state)
        a = self.model.agents[agent]
        vi = a.var_index(var) if isinstance(var, str) else int(var)
        dt = a.vars[vi].dtype
        res = np.zeros(1, dtype=dt)
        cv = np.asarray([0 if count_value is None else count_value], dtype=dt)
        self._check(self._fn("sim_reduce_agent_var")(self._h, agent, state, vi, self.REDUCE_OPS[op],
                                                     cv.ctypes.data if op == "count" else None, res.ctypes.data))
        return res[0]

    def set_constant(self, name: str, value) -> None:
        ctype = {c[0]: c[1] for c in self.model.constants}[name]
        arr = np.ascontiguousarray(value, dtype=_NP_TYPES[ctype]).reshape(-1)
        self._check(self._fn("sim_set_constant")(self._h, name.encode(), arr.ctypes.data))

    def get_constant(self, name: str):
        ctype, length = {c[0]: (c[1], c[2]) for c in self.model.constants}[name]
        arr = np.zeros(max(1, length), dtype=_NP_TYPES[ctype])
        self._check(self._fn("sim_get_constant")(self._h, name.encode(), arr.ctypes.data))
        return arr[0] if length <= 1 else arr


# ---- the CUDA runtime ---------------------------------------------------------
_RT: Optional[C.CDLL] = None


def load_runtime() -> C.CDLL:
    """Load libfgpu_rt.so; raise if it has not been built (no fallback)."""
    global _RT
    if _RT is not None:
        return _RT
    path = build.runtime_path()
    if not os.path.exists(path):
        raise FgpuError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(the CUDA runtime is required; there is no CPU fallback)")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    bind_sim_api(lib, "fgpu")
    lib.fgpu_model_load.restype = C.POINTER(_Model)
    lib.fgpu_model_load.argtypes = [C.c_char_p]
    lib.fgpu_sim_create.restype = C.c_void_p
    lib.fgpu_sim_create.argtypes = [C.POINTER(_Model), C.c_int]
    lib.fgpu_sim_step_async.restype = C.c_int
    lib.fgpu_sim_step_async.argtypes = [C.c_void_p, C.c_int]
    lib.fgpu_sim_sync.restype = C.c_int
    lib.fgpu_sim_sync.argtypes = [C.c_void_p]
    lib.fgpu_sim_step_timed.restype = C.c_int
    lib.fgpu_sim_step_timed.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p]
    lib.fgpu_sim_launch_count.restype = C.c_longlong
    lib.fgpu_sim_launch_count.argtypes = [C.c_void_p]
    lib.fgpu_sim_last_step_ms.restype = C.c_float
    lib.fgpu_sim_last_step_ms.argtypes = [C.c_void_p]
    lib.fgpu_sim_set_option.restype = C.c_int
    lib.fgpu_sim_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    lib.fgpu_sim_profile_iteration.restype = C.c_int
    lib.fgpu_sim_profile_iteration.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int]
    lib.fgpu_sim_profile_iterations.restype = C.c_int
    lib.fgpu_sim_profile_iterations.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int]
    for fn in ("fgpu_sim_save_snapshot", "fgpu_sim_load_snapshot"):
        getattr(lib, fn).restype = C.c_int
        getattr(lib, fn).argtypes = [C.c_void_p, C.c_char_p]
    lib.fgpu_sim_save_states.restype = C.c_int
    lib.fgpu_sim_save_states.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    lib.fgpu_sim_read_agents.restype = C.c_int
    lib.fgpu_sim_read_agents.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.fgpu_sim_device_agents.restype = C.c_void_p
    lib.fgpu_sim_device_agents.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.fgpu_nccl_unique_id.restype = C.c_int
    lib.fgpu_nccl_unique_id.argtypes = [C.c_void_p]
    lib.fgpu_sim_slab_init.restype = C.c_int
    lib.fgpu_sim_slab_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
    lib.fgpu_sim_slab_info.restype = C.c_int
    lib.fgpu_sim_slab_info.argtypes = [C.c_void_p, C.c_void_p]
    lib.fgpu_states_file_count.restype = C.c_int
    lib.fgpu_states_file_count.argtypes = [C.POINTER(_Model), C.c_char_p, C.c_int]
    lib.fgpu_states_file_read_var.restype = C.c_int
    lib.fgpu_states_file_read_var.argtypes = [C.POINTER(_Model), C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    lib.fgpu_debug_sort_pairs.restype = C.c_int
    lib.fgpu_debug_sort_pairs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.fgpu_debug_compact.restype = C.c_int
    lib.fgpu_debug_compact.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    _RT = lib
    return lib


def load_model(plugin_path: str):
    lib = load_runtime()
    if not os.path.exists(plugin_path):
        raise FgpuError(f"model plugin {plugin_path} is missing: build it with ev_diffusion_b200.build.build_model()")
    m = lib.fgpu_model_load(os.fsencode(plugin_path))
    if not m:
        raise FgpuError(lib.fgpu_last_error().decode())
    return m


def read_states_file(plugin_path: str, xml_path: str) -> Dict[str, Dict[str, np.ndarray]]:
    """{agent: {variable: array}} of a states file (0.xml), parsed by the runtime's own reader -- the one
    ``Simulation.load_states`` uses (readInitialStates, io.xslt:211-621) -- without a device: host code only."""
    lib = load_runtime()
    model = load_model(plugin_path)
    info = ModelInfo(model.contents)
    out: Dict[str, Dict[str, np.ndarray]] = {}
    for ai, a in enumerate(info.agents):
        n = lib.fgpu_states_file_count(model, os.fsencode(xml_path), ai)
        if n < 0:
            raise FgpuError(lib.fgpu_last_error().decode())
        out[a.name] = {}
        for vi, v in enumerate(a.vars):
            col = np.empty(n * v.length, dtype=v.dtype)
            if lib.fgpu_states_file_read_var(model, os.fsencode(xml_path), ai, vi, col.ctypes.data, n) != 0:
                raise FgpuError(lib.fgpu_last_error().decode())
            out[a.name][v.name] = np.ascontiguousarray(col.reshape(v.length, n).T) if v.length > 1 else col
    return out


class Simulation(SimBase):
    """One simulation on one B200: ``initialise`` / ``singleIteration`` /
    ``cleanup`` of the reference (``simulation.xslt:472,878,779``) behind the C ABI."""

    _prefix = "fgpu"

    def __init__(self, plugin_path: str, device: int = 0):
        lib = load_runtime()
        mptr = load_model(plugin_path)
        handle = lib.fgpu_sim_create(mptr, int(device))
        if not handle:
            raise FgpuError(lib.fgpu_last_error().decode())
        super().__init__(lib, C.c_void_p(handle), ModelInfo(mptr.contents))
        self._model_ptr = mptr

    def step_async(self, n: int = 1) -> None:
        self._check(self._lib.fgpu_sim_step_async(self._h, int(n)))

    def set_agents_async(self, agent, state, columns: Dict[str, np.ndarray], count: Optional[int] = None) -> None:
        """set_agents without the final synchronisation (fgpu_sim_set_agents_async): every variable must be given,
        ideally from pinned arrays, which must stay untouched until sync()."""
        agent = self._agent(agent)
        state = self._state(agent, state)
        n, ptrs, keep = self._columns(agent, columns, count)
        self._keep_async = keep
        self._lib.fgpu_sim_set_agents_async.restype = C.c_int
        self._lib.fgpu_sim_set_agents_async.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        self._check(self._lib.fgpu_sim_set_agents_async(self._h, agent, state, n, ptrs))

    def read_agents_async(self, agent, state, out: Dict[str, np.ndarray], count: int) -> int:
        """read_agents, enqueued only: `out` is complete after the next sync()."""
        return self.read_agents(agent, state, out, count, _async=True)

    def read_agents(self, agent, state, out: Dict[str, np.ndarray], count: Optional[int] = None, _async: bool = False) -> int:
        """Fill caller-owned arrays (ideally pinned) with the named variables of a state list: all copies
        are issued before one synchronisation (fgpu_sim_read_agents).  Returns the number of agents read."""
        agent = self._agent(agent)
        state = self._state(agent, state)
        a = self.model.agents[agent]
        if count is None:
            count = self.agent_count(agent, state)
        ptrs = (C.c_void_p * len(a.vars))()
        for name, arr in out.items():
            vi = a.var_index(name)
            # an array variable is delivered as `length` blocks of `count` values (the C ABI's element-major layout)
            if arr.dtype != a.vars[vi].dtype or arr.size < count * a.vars[vi].length or not arr.flags["C_CONTIGUOUS"]:
                raise ValueError(f"read_agents: bad destination for '{name}'")
            ptrs[vi] = arr.ctypes.data
        if _async:
            self._lib.fgpu_sim_read_agents_async.restype = C.c_int
            self._lib.fgpu_sim_read_agents_async.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
            self._check(self._lib.fgpu_sim_read_agents_async(self._h, agent, state, count, ptrs))
        else:
            self._check(self._lib.fgpu_sim_read_agents(self._h, agent, state, count, ptrs))
        return count

    def run_exit_functions(self) -> None:
        """The model's exit functions (simulation.xslt:779-790; cleanup() runs them before freeing)."""
        self._lib.fgpu_sim_run_exit_functions.restype = C.c_int
        self._lib.fgpu_sim_run_exit_functions.argtypes = [C.c_void_p]
        self._check(self._lib.fgpu_sim_run_exit_functions(self._h))

    def save_states(self, outputpath: str, iteration_number: int) -> str:
        """saveIterationData(): writes <outputpath><iteration_number>.xml in the reference's format."""
        self._check(self._lib.fgpu_sim_save_states(self._h, os.fsencode(outputpath), int(iteration_number)))
        return f"{outputpath}{int(iteration_number)}.xml"

    def save_snapshot(self, path: str) -> None:
        """Exact binary checkpoint: state lists, rand48 seeds, environment constants, iteration counter."""
        self._check(self._lib.fgpu_sim_save_snapshot(self._h, os.fsencode(path)))

    def load_snapshot(self, path: str) -> None:
        self._check(self._lib.fgpu_sim_load_snapshot(self._h, os.fsencode(path)))

    def sync(self) -> None:
        self._check(self._lib.fgpu_sim_sync(self._h))

    def step_timed(self, n: int, flush_bytes: int = 0) -> np.ndarray:
        """n iterations, device milliseconds of each (cold L2 when flush_bytes > 0)."""
        ms = np.zeros(n, dtype=np.float32)
        self._check(self._lib.fgpu_sim_step_timed(self._h, int(n), int(flush_bytes), ms.ctypes.data))
        return ms

    @property
    def launch_count(self) -> int:
        return int(self._lib.fgpu_sim_launch_count(self._h))

    @property
    def last_step_ms(self) -> float:
        return float(self._lib.fgpu_sim_last_step_ms(self._h))

    def set_option(self, key: str, value: int) -> None:
        self._check(self._lib.fgpu_sim_set_option(self._h, key.encode(), int(value)))

    # -- 1-D slab decomposition (one process per GPU)
    def slab_init(self, rank: int, world: int, unique_id: bytes, halo_cap: int = 0, mig_cap: int = 0) -> None:
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._check(self._lib.fgpu_sim_slab_init(self._h, int(rank), int(world), buf, int(halo_cap), int(mig_cap)))

    def slab_info(self) -> dict:
        out = np.zeros(8, dtype=np.int32)
        self._check(self._lib.fgpu_sim_slab_info(self._h, out.ctypes.data))
        keys = ("enabled", "axis", "nplanes", "p0", "p1", "halo_cap", "mig_cap", "plane_cells")
        d = dict(zip(keys, (int(v) for v in out)))
        d["p2p"] = bool(d["plane_cells"] >> 30)
        d["plane_cells"] &= (1 << 30) - 1
        return d

    def profile_iteration(self):
        return self.profile_iterations(1)

    def profile_iterations(self, iters: int):
        """Steady-state device time per stage over `iters` back-to-back iterations (stream mode)."""
        cap = 256
        names = (C.c_char_p * cap)()
        ms = (C.c_float * cap)()
        n = self._check(self._lib.fgpu_sim_profile_iterations(self._h, int(iters), names, ms, cap))
        return [(names[i].decode(), float(ms[i])) for i in range(n)]


def set_output_dir(path: str) -> None:
    """getOutputDir() of step/exit functions (process-wide, like the reference's main.cu)."""
    lib = load_runtime()
    lib.fgpu_set_output_dir.argtypes = [C.c_char_p]
    lib.fgpu_set_output_dir.restype = None
    lib.fgpu_set_output_dir(os.fsencode(path))


def nccl_unique_id() -> bytes:
    """ncclGetUniqueId() through the runtime (rank 0 calls this; the host program broadcasts it)."""
    lib = load_runtime()
    buf = C.create_string_buffer(128)
    if lib.fgpu_nccl_unique_id(buf) < 0:
        raise FgpuError(lib.fgpu_last_error().decode())
    return buf.raw


def debug_sort_pairs(keys: np.ndarray, key_bits: int, max_digit_bits: int = 8):
    lib = load_runtime()
    keys = np.ascontiguousarray(keys, dtype=np.uint32)
    ko = np.empty_like(keys)
    vo = np.empty_like(keys)
    rc = lib.fgpu_debug_sort_pairs(keys.ctypes.data, keys.shape[0], key_bits, max_digit_bits, ko.ctypes.data, vo.ctypes.data)
    if rc < 0:
        raise FgpuError(lib.fgpu_last_error().decode())
    return ko, vo


def debug_compact(flags: np.ndarray, payload: np.ndarray):
    lib = load_runtime()
    flags = np.ascontiguousarray(flags, dtype=np.int32)
    payload = np.ascontiguousarray(payload, dtype=np.uint32)
    out = np.empty_like(payload)
    cnt = C.c_int(0)
    rc = lib.fgpu_debug_compact(flags.ctypes.data, payload.ctypes.data, flags.shape[0], out.ctypes.data, C.byref(cnt))
    if rc < 0:
        raise FgpuError(lib.fgpu_last_error().decode())
    return out[:cnt.value].copy()
