"""Glue generator: XMML -> header.h + model plugin source.

This is the build-time replacement for the reference's XSLT templates
(`FLAMEGPU/templates/header.xslt`, `FLAMEGPU_kernals.xslt`, `simulation.xslt`).
It emits only *thin* code: the model API header that ``functions.c`` compiles
against unchanged, one wrapper kernel per ``<gpu:function>`` and the C
descriptor table of ``include/fgpu_model.h``.  Everything that executes per
iteration apart from the user's agent scripts lives in the model-agnostic
runtime (``csrc/fgpu_rt.cu``).

Two back ends share the model walk:
  * ``emit_cuda``   -- sm_100a plugin (device proxies, packed message records);
  * ``emit_oracle`` -- host-C++ restatement with the reference's own struct
    layouts and cursor logic (used only by tests / the CPU baseline).
"""
from __future__ import annotations

import os
import re
from typing import Dict, List, Optional, Set, Tuple

from .looprewrite import rewrite_message_loops
from .xmml import Agent, Function, Message, Model, ModelError, Variable, _f32

# --------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------


def _float_lit(v: float) -> str:
    """A C float literal that round-trips the binary32 value."""
    f = _f32(v)
    s = repr(float(f))
    if "e" not in s and "." not in s and "inf" not in s and "nan" not in s:
        s += ".0"
    return s + "f"


def _cast_lit(ctype: str, literal: str) -> str:
    return f"({ctype})({literal})"


def record_layout(msg: Message) -> Tuple[Dict[str, int], int]:
    """Byte offset of each message variable inside the packed record and the
    number of 16-byte quads.  XML order, natural alignment."""
    off = 0
    offsets = {}
    for v in msg.variables:
        sz = v.size
        off = (off + sz - 1) // sz * sz
        offsets[v.name] = off
        off += sz
    quads = max(1, (off + 15) // 16)
    return offsets, quads


def agent_list_offsets(agent: Agent) -> Tuple[Dict[str, int], int]:
    """offsetof() of each column in xmachine_memory_<A>_list (header.xslt:187-194)."""
    n = agent.buffer_size
    off = 8 * n  # int _position[MAX]; int _scan_input[MAX];
    offsets = {}
    maxalign = 4
    for v in agent.variables:
        sz = v.size
        maxalign = max(maxalign, sz)
        off = (off + sz - 1) // sz * sz
        offsets[v.name] = off
        off += sz * n * max(1, v.array_length)
    total = (off + maxalign - 1) // maxalign * maxalign
    return offsets, total


# --------------------------------------------------------------------------
# functions.c analysis: which agent variables does a script write?
# --------------------------------------------------------------------------

_COMMENT_RE = re.compile(r"//[^\n]*|/\*.*?\*/", re.S)
_STRING_RE = re.compile(r'"(?:\\.|[^"\\])*"|\'(?:\\.|[^\'\\])*\'')


def _strip(src: str) -> str:
    src = _COMMENT_RE.sub(" ", src)
    return _STRING_RE.sub('""', src)


def find_function_body(src: str, fname: str) -> Optional[Tuple[str, str]]:
    """Return (parameter list text, body text) of ``int fname(...) {...}``."""
    for m in re.finditer(r"\bint\s+" + re.escape(fname) + r"\s*\(", src):
        i = m.end()
        depth = 1
        while i < len(src) and depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        params = src[m.end():i - 1]
        j = i
        while j < len(src) and src[j] in " \t\r\n":
            j += 1
        if j >= len(src) or src[j] != "{":
            continue  # a prototype
        k = j + 1
        depth = 1
        while k < len(src) and depth:
            depth += {"{": 1, "}": -1}.get(src[k], 0)
            k += 1
        return params, src[j:k]
    return None


def written_variables(func: Function, agent: Agent, sources: List[str]) -> Set[str]:
    """Conservative write set of an agent script.  The reference scatters every
    variable back (krn:1383-1384); we only store those the script can modify.
    Anything we cannot prove read-only is treated as written."""
    allv = {v.name for v in agent.variables}
    for raw in sources:
        src = _strip(raw)
        # a macro that hides a member access defeats the textual analysis
        if re.search(r"#\s*define[^\n]*->", src):
            return allv
        found = find_function_body(src, func.name)
        if not found:
            continue
        params, body = found
        first = params.split(",")[0].strip()
        pm = re.search(r"([A-Za-z_]\w*)\s*$", first)
        if not pm:
            return allv
        p = pm.group(1)
        # the pointer escapes (passed on, dereferenced whole, reassigned)
        for use in re.finditer(r"\b" + re.escape(p) + r"\b", body):
            tail = body[use.end():use.end() + 8].lstrip()
            if not tail.startswith("->"):
                return allv
        # fail closed: a member name after `p->` that is not a variable of the agent is a macro (or a typo the compiler
        # will report); its expansion is invisible here, so every variable counts as written
        for member in re.finditer(r"\b" + re.escape(p) + r"\s*->\s*([A-Za-z_]\w*)", body):
            if member.group(1) not in allv:
                return allv
        written = set()
        assign = r"\s*(?:\[[^\]]*\])?\s*(?:=(?!=)|\+=|-=|\*=|/=|%=|&=|\|=|\^=|<<=|>>=|\+\+|--)"
        for v in agent.variables:
            ref = r"\b" + re.escape(p) + r"\s*->\s*" + re.escape(v.name) + r"\b"
            if (re.search(ref + assign, body) or re.search(r"(?:\+\+|--)\s*" + ref, body)
                    or re.search(r"&\s*" + ref, body) or v.array_length):
                written.add(v.name)
        return written
    return allv


def message_mirror(func: Function, agent: Agent, message, sources: List[str]) -> Dict[str, str]:
    """Agent variables whose value a producer's message carries unchanged: {agent variable: message variable}.

    A script of the usual shape ends with ``add_<M>_message(messages, agent->id, agent->x, agent->y, agent->z)``.  When the call is
    the function's one unconditional output, an argument is literally ``agent->v``, the types agree and the function never
    writes ``v``, message k holds the value variable v of list index k has when the function returns.  A consumer that runs in
    cell order right behind it (thread t <-> sorted message t <-> list index order[t]) can then take v from its own message
    -- one coalesced 16-byte load -- instead of one random 4-byte gather per variable (fgpu_launch.mirror).  Anything that
    cannot be proved from the text gives an empty map: macros hiding member accesses, a conditional or repeated call, a
    ``return`` before the call, an argument that is any other expression."""
    if func.output_message != message.name or (func.output_type or "single_message") != "single_message":
        return {}
    written = written_variables(func, agent, sources)
    scalars = {v.name: v for v in agent.variables if not v.array_length}
    if written >= set(scalars):
        return {}
    for raw in sources:
        src = _strip(raw)
        found = find_function_body(src, func.name)
        if not found:
            continue
        params, body = found
        pm = re.search(r"([A-Za-z_]\w*)\s*$", params.split(",")[0].strip())
        if not pm:
            return {}
        p = pm.group(1)
        calls = list(re.finditer(r"\badd_" + re.escape(message.name) + r"_message\s*\(", body))
        if len(calls) != 1 or re.search(r"^[ \t]*#", body, re.M):     # conditional compilation inside the body: the text is not the program
            return {}
        c = calls[0]
        before = body[:c.start()]
        # the call is a top-level statement of the function body (brace depth 1) and nothing returns before it
        if before.count("{") - before.count("}") != 1 or re.search(r"\breturn\b", before):
            return {}
        i, depth = c.end(), 1
        while i < len(body) and depth:
            depth += {"(": 1, ")": -1}.get(body[i], 0)
            i += 1
        args, cur, depth = [], "", 0
        for ch in body[c.end():i - 1] + ",":
            if ch == "," and depth == 0:
                args.append(cur.strip())
                cur = ""
                continue
            depth += ch in "([{"
            depth -= ch in ")]}"
            cur += ch
        if len(args) != len(message.variables) + 1:
            return {}
        out: Dict[str, str] = {}
        for arg, mv in zip(args[1:], message.variables):
            m = re.fullmatch(re.escape(p) + r"\s*->\s*([A-Za-z_]\w*)", arg)
            if m and m.group(1) in scalars and m.group(1) not in written and scalars[m.group(1)].type == mv.type \
                    and m.group(1) not in out:
                out[m.group(1)] = mv.name
        return out
    return {}


def mirror_producer(model: Model, f: Function, sources: List[str]) -> Tuple[Optional[Function], Dict[str, str]]:
    """The one function whose message `f` consumes and the variables of f's agent that message carries (see message_mirror);
    (None, {}) unless the message is spatially partitioned and has exactly one producer, on f's own agent type."""
    if not f.input_message or model.message(f.input_message).partitioning != "spatial":
        return None, {}
    producers = [g for g in model.all_functions() if g.output_message == f.input_message]
    if len(producers) != 1 or producers[0].agent != f.agent:
        return None, {}
    m = message_mirror(producers[0], model.agent(f.agent), model.message(f.input_message), sources)
    return (producers[0], m) if m else (None, {})


# --------------------------------------------------------------------------
# shared header pieces
# --------------------------------------------------------------------------

_COLOURS = """
#define FLAME_GPU_VISUALISATION_COLOUR_BLACK 0
#define FLAME_GPU_VISUALISATION_COLOUR_RED 1
#define FLAME_GPU_VISUALISATION_COLOUR_GREEN 2
#define FLAME_GPU_VISUALISATION_COLOUR_BLUE 3
#define FLAME_GPU_VISUALISATION_COLOUR_YELLOW 4
#define FLAME_GPU_VISUALISATION_COLOUR_CYAN 5
#define FLAME_GPU_VISUALISATION_COLOUR_MAGENTA 6
#define FLAME_GPU_VISUALISATION_COLOUR_WHITE 7
#define FLAME_GPU_VISUALISATION_COLOUR_BROWN 8
"""


def _common_defs(model: Model) -> str:
    out = []
    out.append("#define THREADS_PER_TILE 64")
    out.append("#define USE_CUDA_STREAMS")
    out.append("#define FLAME_GPU_MAJOR_VERSION 1\n#define FLAME_GPU_MINOR_VERSION 5\n#define FLAME_GPU_PATCH_VERSION 0")
    out.append("#define MAX_FILEPATH_LENGTH 4096")
    out.append("typedef unsigned int uint;")
    for n in ("vec2", "vec3", "vec4", "ivec2", "ivec3", "ivec4", "uvec2", "uvec3", "uvec4", "dvec2", "dvec3", "dvec4"):
        out.append(f"typedef glm::{n} {n};")
    for n in ("2", "3", "4"):
        out.append(f"typedef glm::vec{n} fvec{n};")
    if any(v.type == "double" for a in model.agents for v in a.variables) or \
            any(v.type == "double" for m in model.messages for v in m.variables):
        out.append("#define _DOUBLE_SUPPORT_REQUIRED_")
    out.append(f"#define buffer_size_MAX {model.buffer_size_max}")
    for a in model.agents:
        out.append(f"#define xmachine_memory_{a.name}_MAX {a.buffer_size}")
        for v in a.variables:
            if v.array_length:
                out.append(f"#define xmachine_memory_{a.name}_{v.name}_LENGTH {v.array_length}")
    for m in model.messages:
        out.append(f"#define xmachine_message_{m.name}_MAX {m.buffer_size}")
        tag = {"spatial": "partitioningSpatial", "none": "partitioningNone"}[m.partitioning]
        out.append(f"#define xmachine_message_{m.name}_{tag}")
        if m.partitioning == "spatial":
            out.append(f"#define xmachine_message_{m.name}_grid_size {m.grid_size()}")
    out.append(_COLOURS)
    out.append("enum MESSAGE_OUTPUT{\n\tsingle_message,\n\toptional_message,\n};")
    out.append("enum AGENT_TYPE{\n\tCONTINUOUS,\n\tDISCRETE_2D\n};")
    return "\n".join(out) + "\n"


def _agent_structs(model: Model, align: str) -> str:
    out = []
    for a in model.agents:
        out.append(f"struct {align} xmachine_memory_{a.name}\n{{")
        for v in a.variables:
            star = "*" if v.array_length else ""
            out.append(f"    {v.type} {star}{v.name};")
        out.append("};\n")
        out.append(f"struct xmachine_memory_{a.name}_list\n{{")
        out.append(f"    int _position [xmachine_memory_{a.name}_MAX];")
        out.append(f"    int _scan_input [xmachine_memory_{a.name}_MAX];")
        for v in a.variables:
            mul = f"*{v.array_length}" if v.array_length else ""
            out.append(f"    {v.type} {v.name} [xmachine_memory_{a.name}_MAX{mul}];")
        out.append("};\n")
    return "\n".join(out)


def _function_signature(model: Model, f: Function, prefix: str = "") -> str:
    """Agent script signature rule (header.xslt:375-378)."""
    args = [f"xmachine_memory_{f.agent}* agent"]
    if f.xagent_output:
        args.append(f"xmachine_memory_{f.xagent_output}_list* {f.xagent_output}_agents")
    if f.input_message:
        args.append(f"xmachine_message_{f.input_message}_list* {f.input_message}_messages")
        if model.message(f.input_message).partitioning == "spatial":
            args.append(f"xmachine_message_{f.input_message}_PBM* partition_matrix")
    if f.output_message:
        args.append(f"xmachine_message_{f.output_message}_list* {f.output_message}_messages")
    if f.rng:
        args.append("RNG_rand48* rand48")
    return f"{prefix}int {f.name}({', '.join(args)})"


def _msg_args(m: Message) -> str:
    return ", ".join(f"{v.type} {v.name}" for v in m.variables)


def _agent_args(a: Agent) -> str:
    return ", ".join(f"{v.type} {v.name}" for v in a.variables if not v.array_length)


# --------------------------------------------------------------------------
# CUDA back end
# --------------------------------------------------------------------------

def _cuda_message_api(m: Message) -> str:
    name = m.name
    T = f"fgpu_traits_{name}"
    S = f"xmachine_message_{name}"
    out = []
    # user-visible message struct == the record stored in HBM (payload only; the
    # cursor lives in the proxy).  __align__(16) pads sizeof to a 16-byte multiple.
    out.append(f"struct __align__(16) {S}\n{{")
    for v in m.variables:
        out.append(f"    {v.type} {v.name};")
    out.append("};")
    out.append(f"struct {T} {{")
    out.append(f"    static constexpr int MAX = {m.buffer_size};")
    out.append(f"    static constexpr int NQ = (int)(sizeof({S}) / 16);")
    if m.partitioning == "spatial":
        dx, dy, dz = m.partition_dims()
        out.append(f"    static constexpr int DIMX = {dx}, DIMY = {dy}, DIMZ = {dz};")
        out.append(f"    static constexpr bool IS2D = {'true' if m.is_2d() else 'false'};")
        for ax, lo, hi in zip("XYZ", m.min_literals, m.max_literals):
            out.append(f"    static constexpr float MIN{ax} = (float){lo}, MAX{ax} = (float){hi};")
    else:
        out.append("    static constexpr int DIMX = 1, DIMY = 1, DIMZ = 1;")
        out.append("    static constexpr bool IS2D = false;")
        for ax in "XYZ":
            out.append(f"    static constexpr float MIN{ax} = 0.0f, MAX{ax} = 1.0f;")
    out.append("};")
    out.append(f"struct {S}_list : fgpu::MessageProxy<{S}, {T}> {{}};")
    if m.partitioning == "spatial":
        out.append(f"struct {S}_PBM {{ int _unused; }};")
    # add_<M>_message (krn:382-404) fused with the cell hash (krn:899-911 / 944-953)
    out.append(f"FGPU_DEV void add_{name}_message({S}_list* messages, {_msg_args(m)}){{")
    out.append(f"    union {{ {S} m; uint4 q[{T}::NQ]; }} r;")
    out.append(f"    for (int _q = 0; _q < {T}::NQ; ++_q) r.q[_q] = make_uint4(0u, 0u, 0u, 0u);")
    for v in m.variables:
        out.append(f"    r.m.{v.name} = {v.name};")
    out.append(f"    uint4* _dst = (uint4*)(messages->out_recs + messages->out_index);")
    out.append(f"    for (int _q = 0; _q < {T}::NQ; ++_q) _dst[_q] = r.q[_q];")
    if m.partitioning == "spatial":
        out.append("    int _cx, _cy, _cz;")
        out.append(f"    fgpu::grid_position<{T}>((float)x, (float)y, (float)z, _cx, _cy, _cz);")
        out.append(f"    messages->out_keys[messages->out_index] = fgpu::cell_hash<{T}>(_cx, _cy, _cz);")
    out.append("    if (messages->out_flags != nullptr) messages->out_flags[messages->out_index] = 1;   /* optional_message, krn:394-397 */")
    out.append("}")
    if m.partitioning == "spatial":
        out.append(f"""FGPU_DEV {S}* get_first_{name}_message({S}_list* messages, {S}_PBM* partition_matrix, float x, float y, float z){{
    (void)partition_matrix;
    if (messages->in_count == 0) return nullptr;   /* krn:1117 */
    fgpu::begin_walk<{S}, {T}>(messages, x, y, z);
    return fgpu::open_next_range<{S}, {T}>(messages);
}}
FGPU_DEV {S}* get_next_{name}_message({S}* current, {S}_list* messages, {S}_PBM* partition_matrix){{
    (void)partition_matrix;
    return fgpu::next_message<{S}, {T}>(current, messages);
}}
/* the two levels of the walk, for message loops rewritten by looprewrite.py */
FGPU_DEV {S}* fgpu_range_end_{name}({S}_list* messages){{ return const_cast<{S}*>(messages->end); }}
FGPU_DEV int fgpu_range_count_{name}({S}_list* messages){{ return messages->cnt; }}
FGPU_DEV {S}* fgpu_next_range_{name}({S}_list* messages){{ return fgpu::open_next_range<{S}, {T}>(messages); }}""")
    else:
        out.append(f"""FGPU_DEV {S}* get_first_{name}_message({S}_list* messages){{
    if (messages->in_count == 0) return nullptr;
    messages->end = messages->in_recs + messages->in_count;
    messages->cnt = messages->in_count;
    return const_cast<{S}*>(messages->in_recs);
}}
FGPU_DEV {S}* get_next_{name}_message({S}* current, {S}_list* messages){{
    {S}* next = current + 1;
    return next != messages->end ? next : nullptr;
}}
FGPU_DEV {S}* fgpu_range_end_{name}({S}_list* messages){{ return const_cast<{S}*>(messages->end); }}
FGPU_DEV int fgpu_range_count_{name}({S}_list* messages){{ return messages->cnt; }}
FGPU_DEV {S}* fgpu_next_range_{name}({S}_list* messages){{ (void)messages; return nullptr; }}""")
    return "\n".join(out) + "\n"


def _cuda_agent_api(a: Agent) -> str:
    scal = [v for v in a.variables if not v.array_length]
    out = [f"""template <int AGENT_TYPE>
FGPU_DEV void add_{a.name}_agent(xmachine_memory_{a.name}_list* agents, {_agent_args(a)}){{
    /* krn:287-314: newborn of parent k goes to sparse slot k of d_{a.name}s_new, flag = 1 */
    const fgpu::NewAgentProxy* _p = (const fgpu::NewAgentProxy*)(const void*)agents;
    xmachine_memory_{a.name}_list* _l = (xmachine_memory_{a.name}_list*)_p->list;
    const int index = _p->index;
    _l->_position[index] = 0;
    _l->_scan_input[index] = 1;"""]
    for v in scal:
        out.append(f"    _l->{v.name}[index] = {v.name};")
    out.append("}")
    out.append(f"FGPU_DEV void add_{a.name}_agent(xmachine_memory_{a.name}_list* agents, {_agent_args(a)}){{")
    out.append(f"    add_{a.name}_agent<DISCRETE_2D>(agents, {', '.join(v.name for v in scal)});\n}}")
    if any(v.array_length for v in a.variables):
        out.append(f"""/* krn:336-367: `array` points at element 0 of this agent; elements are xmachine_memory_{a.name}_MAX apart */
template<typename T> FGPU_DEV T get_{a.name}_agent_array_value(T *array, unsigned int index){{
    return array != nullptr ? array[(size_t)index * xmachine_memory_{a.name}_MAX] : T(0);
}}
template<typename T> FGPU_DEV void set_{a.name}_agent_array_value(T *array, unsigned int index, T value){{
    if (array != nullptr) array[(size_t)index * xmachine_memory_{a.name}_MAX] = value;
}}""")
    return "\n".join(out) + "\n"


def emit_cuda_header(model: Model) -> str:
    out = ["/* generated by ev_diffusion_b200.codegen -- model API header (replaces header.xslt output) */",
           "#ifndef __HEADER\n#define __HEADER",
           "#include <cuda_runtime.h>",
           '#include "fgpu_glm.h"',
           '#include "fgpu_device.cuh"',
           "#define __FLAME_GPU_FUNC__ __device__",
           "#define __FLAME_GPU_INIT_FUNC__\n#define __FLAME_GPU_STEP_FUNC__\n#define __FLAME_GPU_EXIT_FUNC__",
           "#define __FLAME_GPU_HOST_FUNC__ __host__",
           "#define FGPU_B200 1"]
    out.append(_common_defs(model))
    out.append(_agent_structs(model, "__align__(16)"))
    for m in model.messages:
        out.append(_cuda_message_api(m))
    out.append("struct RNG_rand48 : fgpu::Rand48Proxy {};")
    out.append("""template <int AGENT_TYPE> FGPU_DEV float rnd(RNG_rand48* rand48){ return fgpu::rand48_draw(rand48); }
FGPU_DEV float rnd(RNG_rand48* rand48){ return fgpu::rand48_draw(rand48); }""")
    for a in model.agents:
        out.append(_cuda_agent_api(a))
    out.append("/* Agent function prototypes (header.xslt:364-378) */")
    for f in model.all_functions():
        out.append(_function_signature(model, f, "__FLAME_GPU_FUNC__ ") + ";")
    out.append("/* environment constants (header.xslt:763-773) */")
    for c in model.constants:
        dim = f"[{c.array_length}]" if c.array_length else ""
        out.append(f"__constant__ {c.type} {c.name}{dim};")
        out.append(f"extern {c.type} h_env_{c.name}{dim};")
        out.append(f"void set_{c.name}({c.type}* h_{c.name});")
        out.append(f"const {c.type}* get_{c.name}();")
    out.append("/* agent id generation (header.xslt:500-512) */")
    out.append(_id_decls(model, cuda=True))
    out.append(_host_api_decls(model))
    out.append("#endif /* __HEADER */")
    return "\n".join(out) + "\n"


def _host_api_decls(model: Model) -> str:
    """Face-2 names (header.xslt:531-645) implemented by the generated shim on
    top of libfgpu_rt."""
    out = ["/* host API (header.xslt:531-645) -- implemented by the plugin shim over libfgpu_rt */",
           "extern unsigned int getIterationNumber();",
           "extern void initialise(char * input);",
           "extern void cleanup();",
           "extern void singleIteration();",
           "extern void set_exit_early();",
           "extern bool get_exit_early();",
           "extern const char* getOutputDir();",
           "/* bounds of the agents read by the last initialise()/readInitialStates(): start at 0, grow with every x, y, z (io.xslt:256-261,418-439,623-629) */",
           "extern glm::vec3 getMaximumBounds();",
           "extern glm::vec3 getMinimumBounds();"]
    # io.xslt:124,211: one (host list, device list, count) triple per agent state / one (host list, count) pair per agent
    save_args = "".join(f", xmachine_memory_{a.name}_list* h_{a.name}s_{s}, xmachine_memory_{a.name}_list* d_{a.name}s_{s}, int h_xmachine_memory_{a.name}_{s}_count"
                        for a in model.agents for s in a.states)
    read_args = "".join(f", xmachine_memory_{a.name}_list* h_{a.name}s, int* h_xmachine_memory_{a.name}_count" for a in model.agents)
    out.append(f"extern void saveIterationData(char* outputpath, int iteration_number{save_args});")
    out.append(f"extern void readInitialStates(char* inputpath{read_args});")
    for a in model.agents:
        out.append(f"extern int get_agent_{a.name}_MAX_count();")
        for s in a.states:
            out.append(f"extern int get_agent_{a.name}_{s}_count();")
            out.append(f"extern void reset_{a.name}_{s}_count();")
            out.append(f"extern xmachine_memory_{a.name}_list* get_device_{a.name}_{s}_agents();")
            out.append(f"extern xmachine_memory_{a.name}_list* get_host_{a.name}_{s}_agents();")
            out.append(f"void sort_{a.name}s_{s}(void (*generate_key_value_pairs)(unsigned int* keys, unsigned int* values, xmachine_memory_{a.name}_list* agents));")
            for v in a.variables:
                if v.array_length:   # header.xslt:648-658
                    out.append(f"{v.type} get_{a.name}_{s}_variable_{v.name}(unsigned int index, unsigned int element);")
                if not v.array_length:
                    out.append(f"{v.type} get_{a.name}_{s}_variable_{v.name}(unsigned int index);")
                    # analytics (header.xslt:729-751): sum for every scalar, count for integers, min/max for non-vectors
                    out.append(f"{v.type} reduce_{a.name}_{s}_{v.name}_variable();")
                    if "int" in v.type:
                        out.append(f"{v.type} count_{a.name}_{s}_{v.name}_variable({v.type} count_value);")
                    if "vec" not in v.type:
                        out.append(f"{v.type} min_{a.name}_{s}_{v.name}_variable();")
                        out.append(f"{v.type} max_{a.name}_{s}_{v.name}_variable();")
    # host-based agent creation (header.xslt:665-706)
    for a in model.agents:
        out.append(f"xmachine_memory_{a.name}* h_allocate_agent_{a.name}();")
        out.append(f"void h_free_agent_{a.name}(xmachine_memory_{a.name}** agent);")
        out.append(f"xmachine_memory_{a.name}** h_allocate_agent_{a.name}_array(unsigned int count);")
        out.append(f"void h_free_agent_{a.name}_array(xmachine_memory_{a.name}*** agents, unsigned int count);")
        for s in a.states:
            out.append(f"void h_add_agent_{a.name}_{s}(xmachine_memory_{a.name}* agent);")
            out.append(f"void h_add_agents_{a.name}_{s}(xmachine_memory_{a.name}** agents, unsigned int count);")
    return "\n".join(out)


def _host_list_defs(model: Model) -> List[str]:
    """get_host_<A>_<S>_agents (header.xslt:605-613), readInitialStates / saveIterationData with the reference's
    parameter lists (io.xslt:124,211) and getMaximum/MinimumBounds (io.xslt:623-629), all on top of the C ABI.

    The reference keeps one host copy of every state list and refreshes it only when it writes an XML file; here the
    host copy is filled from the device every time it is asked for, which is what a caller of the reference has to
    arrange by hand (cudaMemcpy from get_device_*) before the pointer means anything."""
    out = ["/* the reader accumulates the bounds while it parses; here they are folded on the device the first time they are",
           "   asked for after a read (callers -- the visualisation -- ask right after initialise()) */",
           "static glm::vec3 fgpu_agent_maximum(0.0f, 0.0f, 0.0f), fgpu_agent_minimum(0.0f, 0.0f, 0.0f);",
           "static bool fgpu_bounds_stale = false;",
           "static void fgpu_fold_bounds();",
           "static void fgpu_update_bounds(){ fgpu_bounds_stale = true; }",
           "glm::vec3 getMaximumBounds(){ if (fgpu_bounds_stale) fgpu_fold_bounds(); return fgpu_agent_maximum; }",
           "glm::vec3 getMinimumBounds(){ if (fgpu_bounds_stale) fgpu_fold_bounds(); return fgpu_agent_minimum; }"]
    fill: List[str] = []
    for ai, a in enumerate(model.agents):
        cols = ", ".join(("nullptr" if v.array_length else f"(void*)h->{v.name}") for v in a.variables)
        # array variables arrive as `length` blocks of `count` values; the host list keeps the device layout
        # (element e of agent i at var[e * MAX + i], header.xslt:187-194)
        arrays = "".join(f"""
    if (count > 0) {{ std::vector<{v.type}> t((size_t)count * {v.array_length});
        if (fgpu_sim_read_agent_var(fgpu_cur, {ai}, state, {vi}, t.data(), count) != FGPU_OK) fgpu_die();
        for (int e = 0; e < {v.array_length}; ++e) memcpy(h->{v.name} + (size_t)e * xmachine_memory_{a.name}_MAX, t.data() + (size_t)e * count, sizeof({v.type}) * (size_t)count); }}"""
                         for vi, v in enumerate(a.variables) if v.array_length)
        out.append(f"""static void fgpu_fill_host_{a.name}(xmachine_memory_{a.name}_list* h, int state, int count){{
    void* const columns[] = {{{cols}}};
    if (count > 0 && fgpu_sim_read_agents(fgpu_cur, {ai}, state, count, columns) != FGPU_OK) fgpu_die();{arrays}
}}""")
        for si, st in enumerate(a.states):
            out.append(f"""xmachine_memory_{a.name}_list* get_host_{a.name}_{st}_agents(){{
    static xmachine_memory_{a.name}_list* h = nullptr;
    if (!h) h = (xmachine_memory_{a.name}_list*)calloc(1, sizeof(xmachine_memory_{a.name}_list));
    if (!h) {{ fprintf(stderr, "get_host_{a.name}_{st}_agents: out of host memory\\n"); exit(EXIT_FAILURE); }}
    fgpu_fill_host_{a.name}(h, {si}, fgpu_sim_agent_count(fgpu_cur, {ai}, {si}));
    return h;
}}""")
        # bounds over the variables called x, y, z of the initial state list, as the reader accumulates them
        init = a.states.index(a.initial_state)
        for axis in "xyz":
            for vi, v in enumerate(a.variables):
                if v.name == axis and not v.array_length:
                    fill.append(f"    if (fgpu_sim_agent_count(fgpu_cur, {ai}, {init}) > 0) {{ {v.type} lo = 0, hi = 0;\n"
                                f"        if (fgpu_sim_reduce_agent_var(fgpu_cur, {ai}, {init}, {vi}, FGPU_REDUCE_MIN, nullptr, &lo) != FGPU_OK) fgpu_die();\n"
                                f"        if (fgpu_sim_reduce_agent_var(fgpu_cur, {ai}, {init}, {vi}, FGPU_REDUCE_MAX, nullptr, &hi) != FGPU_OK) fgpu_die();\n"
                                f"        if (fgpu_agent_minimum.{axis} > (float)lo) fgpu_agent_minimum.{axis} = (float)lo;\n"
                                f"        if (fgpu_agent_maximum.{axis} < (float)hi) fgpu_agent_maximum.{axis} = (float)hi; }}")
    out.append("static void fgpu_fold_bounds(){\n    fgpu_bounds_stale = false;\n    fgpu_agent_maximum = glm::vec3(0.0f, 0.0f, 0.0f); fgpu_agent_minimum = glm::vec3(0.0f, 0.0f, 0.0f);\n"
               + "\n".join(fill) + "\n}")
    save_args = "".join(f", xmachine_memory_{a.name}_list* h_{a.name}s_{s}, xmachine_memory_{a.name}_list* d_{a.name}s_{s}, int h_xmachine_memory_{a.name}_{s}_count"
                        for a in model.agents for s in a.states)
    read_args = "".join(f", xmachine_memory_{a.name}_list* h_{a.name}s, int* h_xmachine_memory_{a.name}_count" for a in model.agents)
    body = []
    for ai, a in enumerate(model.agents):
        for si, st in enumerate(a.states):
            body.append(f"    if (h_{a.name}s_{st}) fgpu_fill_host_{a.name}(h_{a.name}s_{st}, {si}, fgpu_sim_agent_count(fgpu_cur, {ai}, {si}));")
            body.append(f"    (void)d_{a.name}s_{st}; (void)h_xmachine_memory_{a.name}_{st}_count;")
    out.append(f"""void saveIterationData(char* outputpath, int iteration_number{save_args}){{
    /* io.xslt:124-200: device lists -> host lists -> <outputpath><iteration_number>.xml.  The lists written are the
       simulation's own (the device pointers a caller can pass are the ones get_device_* returned). */
{chr(10).join(body)}
    if (fgpu_sim_save_states(fgpu_cur, outputpath, iteration_number) != FGPU_OK) fgpu_die();
}}""")
    body = []
    for ai, a in enumerate(model.agents):
        init = a.states.index(a.initial_state)
        body.append(f"    {{ int n = fgpu_sim_agent_count(fgpu_cur, {ai}, {init}); if (h_xmachine_memory_{a.name}_count) *h_xmachine_memory_{a.name}_count = n;")
        body.append(f"      if (h_{a.name}s) fgpu_fill_host_{a.name}(h_{a.name}s, {init}, n); }}")
    out.append(f"""void readInitialStates(char* inputpath{read_args}){{
    /* io.xslt:211-621: parse the file into the initial state lists (and the environment constants).  Here the lists
       live on the device, so the file is loaded into the simulation and the caller's host lists are filled from it. */
    if (fgpu_sim_load_states(fgpu_cur, inputpath) != FGPU_OK) fgpu_die();
    fgpu_update_bounds();
{chr(10).join(body)}
}}""")
    return out


def _host_creation_defs(model: Model, add_call: str, count_call: str) -> List[str]:
    """h_allocate_/h_free_/h_add_agent(s)_<A>_<S> (simulation.xslt:1192-1320): host structs with the XML default
    values, unpacked AoS -> SoA columns and appended to the state list through
    `add_call`(agent, state, count, const void* const* columns); `count_call`(agent, state) is the current
    population for the reference's buffer-size guard (message + exit).  Array variables travel as `length` blocks of
    `count` values."""
    out = []
    for ai, a in enumerate(model.agents):
        A = a.name
        inits = "".join(f" agent->{v.name} = ({v.type}){v.default};" for v in a.variables if v.default is not None and not v.array_length)
        arr_alloc = "".join(f" agent->{v.name} = ({v.type}*)calloc({v.array_length}, sizeof({v.type}));"
                            + (f" for (unsigned int index = 0; index < {v.array_length}; index++) agent->{v.name}[index] = ({v.type}){v.default};"
                               if v.default is not None else "")      # simulation.xslt:1200-1207
                            for v in a.variables if v.array_length)
        arr_free = "".join(f" free((*agent)->{v.name});" for v in a.variables if v.array_length)
        out.append(f"xmachine_memory_{A}* h_allocate_agent_{A}(){{ xmachine_memory_{A}* agent = (xmachine_memory_{A}*)malloc(sizeof(xmachine_memory_{A})); "
                   f"memset(agent, 0, sizeof(xmachine_memory_{A}));{inits}{arr_alloc} return agent; }}")
        out.append(f"void h_free_agent_{A}(xmachine_memory_{A}** agent){{{arr_free} free(*agent); *agent = NULL; }}")
        out.append(f"xmachine_memory_{A}** h_allocate_agent_{A}_array(unsigned int count){{ xmachine_memory_{A}** agents = "
                   f"(xmachine_memory_{A}**)malloc(count * sizeof(xmachine_memory_{A}*)); for (unsigned int i = 0; i < count; i++) agents[i] = h_allocate_agent_{A}(); return agents; }}")
        out.append(f"void h_free_agent_{A}_array(xmachine_memory_{A}*** agents, unsigned int count){{ for (unsigned int i = 0; i < count; i++) "
                   f"h_free_agent_{A}(&((*agents)[i])); free(*agents); *agents = NULL; }}")
        for si, st in enumerate(a.states):
            if True:
                # arrays: `length` blocks of `count` values (simulation.xslt:1238-1241 unpacks them element by element too)
                cols = "".join((f" std::vector<{v.type}> c_{v.name}((size_t)count * {v.array_length}); for (unsigned int i = 0; i < count; i++) "
                                f"for (unsigned int j = 0; j < {v.array_length}; j++) c_{v.name}[(size_t)j * count + i] = agents[i]->{v.name}[j];")
                               if v.array_length else
                               f" std::vector<{v.type}> c_{v.name}(count); for (unsigned int i = 0; i < count; i++) c_{v.name}[i] = agents[i]->{v.name};"
                               for v in a.variables)
                ptrs = ", ".join(f"c_{v.name}.data()" for v in a.variables)
                body = (f"{{ if (count == 0) return; if ({count_call}({ai}, {si}) + (int)count > xmachine_memory_{A}_MAX){{ "
                        f'printf("Error: Buffer size of {A} agents in state {st} will be exceeded by h_add_agents_{A}_{st}\\n"); exit(EXIT_FAILURE); }}'
                        f"{cols} const void* columns[] = {{{ptrs}}}; if ({add_call}({ai}, {si}, (int)count, columns) != 0) fgpu_die(); }}")
            out.append(f"void h_add_agents_{A}_{st}(xmachine_memory_{A}** agents, unsigned int count){body}")
            out.append(f"void h_add_agent_{A}_{st}(xmachine_memory_{A}* agent){{ h_add_agents_{A}_{st}(&agent, 1); }}")
    return out


def _reduction_defs(model: Model, call: str) -> List[str]:
    """Definitions of reduce_/count_/min_/max_<A>_<S>_<v>_variable() (simulation.xslt:1325-1357) over
    `call`(agent, state, var, op, count_value*, result*) -> status."""
    out = []
    for ai, a in enumerate(model.agents):
        for si, st in enumerate(a.states):
            for vi, v in enumerate(a.variables):
                if v.array_length:
                    continue
                def body(op, arg="nullptr"):
                    return (f"{{ {v.type} t = 0; if ({call}({ai}, {si}, {vi}, {op}, {arg}, &t) != 0) fgpu_die(); return t; }}")
                out.append(f"{v.type} reduce_{a.name}_{st}_{v.name}_variable(){body(0)}")
                if "int" in v.type:
                    out.append(f"{v.type} count_{a.name}_{st}_{v.name}_variable({v.type} count_value){body(3, '&count_value')}")
                if "vec" not in v.type:
                    out.append(f"{v.type} min_{a.name}_{st}_{v.name}_variable(){body(1)}")
                    out.append(f"{v.type} max_{a.name}_{st}_{v.name}_variable(){body(2)}")
    return out


_ID_TYPES = {"int": "2147483647", "unsigned int": "4294967295u", "unsigned long long": "18446744073709551615ull",
             "unsigned long long int": "18446744073709551615ull"}


def _id_agents(model: Model) -> List[Tuple[int, Agent, Variable]]:
    """Agents that get generate_<A>_id(): an integer scalar variable called `id` (header.xslt:500-512).  atomicAdd
    limits the types to the ones it supports."""
    out = []
    for ai, a in enumerate(model.agents):
        for v in a.variables:
            if v.name == "id" and not v.array_length and v.type in _ID_TYPES:
                out.append((ai, a, v))
    return out


def _id_decls(model: Model, cuda: bool) -> str:
    out = []
    for _, a, v in _id_agents(model):
        out.append(f"extern {v.type} h_current_value_generate_{a.name}_id;")
        if cuda:
            out.append(f"__device__ {v.type} d_current_value_generate_{a.name}_id;")
        out.append(f"__FLAME_GPU_HOST_FUNC__ __FLAME_GPU_FUNC__ {v.type} generate_{a.name}_id();")
        out.append(f"void set_initial_{a.name}_id({v.type} firstID);")
    return "\n".join(out)


def _id_defs(model: Model, cuda: bool) -> str:
    """generate_<A>_id / set_initial_<A>_id (FLAMEGPU_kernals.xslt:1400-1411, simulation.xslt:336-356) and the three
    hooks the runtime calls where the reference synchronises the counters."""
    ids = _id_agents(model)
    out = []
    for ai, a, v in ids:
        A, T = a.name, v.type
        out.append(f"{T} h_current_value_generate_{A}_id = 0;")
        out.append(f"static {T} h_last_value_generate_{A}_id = {_ID_TYPES[T]};")
        if cuda:
            out.append(f"""__FLAME_GPU_HOST_FUNC__ __FLAME_GPU_FUNC__ {T} generate_{A}_id(){{
#if defined(__CUDA_ARCH__)
    return atomicAdd(&d_current_value_generate_{A}_id, ({T})1);
#else
    {T} new_id = h_current_value_generate_{A}_id;
    h_current_value_generate_{A}_id++;
    return new_id;
#endif
}}""")
        else:
            out.append(f"{T} generate_{A}_id(){{ {T} new_id = h_current_value_generate_{A}_id; h_current_value_generate_{A}_id++; return new_id; }}")
        out.append(f"void set_initial_{A}_id({T} firstID){{ h_current_value_generate_{A}_id = firstID; }}")
    if not cuda:
        return "\n".join(out) + "\n"
    if not ids:
        out.append("static void (*const fgpu_ids_after_load)(void*) = nullptr;\nstatic void (*const fgpu_ids_to_host)(void) = nullptr;\n"
                   "static void (*const fgpu_ids_to_device)(void) = nullptr;")
        return "\n".join(out) + "\n"
    out.append("static void fgpu_ids_to_device(void){")
    for ai, a, v in ids:   # update_device_generate_<A>_id: only when the host value moved (simulation.xslt:346-350)
        out.append(f"    if (h_current_value_generate_{a.name}_id != h_last_value_generate_{a.name}_id){{ "
                   f"cudaMemcpyToSymbol(d_current_value_generate_{a.name}_id, &h_current_value_generate_{a.name}_id, sizeof({v.type})); "
                   f"h_last_value_generate_{a.name}_id = h_current_value_generate_{a.name}_id; }}")
    out.append("}")
    out.append("static void fgpu_ids_to_host(void){")
    for ai, a, v in ids:   # update_host_generate_<A>_id (simulation.xslt:352-356)
        out.append(f"    cudaMemcpyFromSymbol(&h_current_value_generate_{a.name}_id, d_current_value_generate_{a.name}_id, sizeof({v.type})); "
                   f"h_last_value_generate_{a.name}_id = h_current_value_generate_{a.name}_id;")
    out.append("}")
    out.append("static void fgpu_ids_after_load(void* sim){")
    for ai, a, v in ids:   # io.xslt:599-616
        init = a.states.index(a.initial_state)
        vi = [x.name for x in a.variables].index("id")
        out.append(f"    if (fgpu_sim_agent_count((fgpu_sim*)sim, {ai}, {init}) > 0){{ {v.type} mx = 0; "
                   f"if (fgpu_sim_reduce_agent_var((fgpu_sim*)sim, {ai}, {init}, {vi}, FGPU_REDUCE_MAX, nullptr, &mx) != FGPU_OK) fgpu_die(); "
                   f"set_initial_{a.name}_id(mx + 1); }} else set_initial_{a.name}_id(0);")
    out.append("    fgpu_ids_to_device();\n}")
    return "\n".join(out) + "\n"


def _carry_vars(a: Agent) -> List[Variable]:
    """Variables of the carried agent record: every scalar variable (none if the agent has array variables)."""
    if any(v.array_length for v in a.variables):
        return []
    return list(a.variables)


def _cuda_carry_struct(a: Agent) -> str:
    """The packed record a message producer stores beside its message (fgpu_launch.carry_out): the build moves it with
    the message, so the consumer that runs in cell order reads its agents' variables with coalesced 16-byte loads
    instead of one random 4-byte read per variable (each of which costs a 64-byte DRAM burst at 16M agents)."""
    vs = _carry_vars(a)
    if not vs:
        return ""
    out = [f"struct __align__(16) fgpu_carry_{a.name}\n{{"]
    for v in vs:
        out.append(f"    {v.type} {v.name};")
    out.append("};")
    return "\n".join(out) + "\n"


def _cuda_kernel(model: Model, f: Function, written: Set[str], mirror: Optional[Dict[str, str]] = None) -> str:
    a = model.agent(f.agent)
    A = a.name
    bs = int(os.environ.get("FGPU_CONSUMER_BLOCK", "128")) if f.input_message else 256   # tuning hook for A/B builds
    minb = int(os.environ.get("FGPU_CONSUMER_MINBLOCKS", "0")) if f.input_message else 0    # ... and a register cap through min blocks per SM
    bounds = f"{bs}, {minb}" if minb else f"{bs}"
    carry = _carry_vars(a)
    out = [f"/* GPUFLAME_{f.name} (krn:1320-1387) */",
           f"__global__ void __launch_bounds__({bounds}) fgpu_k_{f.name}(const fgpu_launch L)\n{{",
           '    asm volatile("griddepcontrol.wait;" ::: "memory");   /* programmatic dependent launch: the previous kernel of the iteration has completed */',
           "    int t = blockIdx.x * blockDim.x + threadIdx.x;"]
    if f.input_message and model.message(f.input_message).partitioning == "spatial":
        out += ["    if (L.win_mode) {   /* slab mode: interior planes first, boundary planes once the halo has arrived */",
                "        const int w0 = (int)(L.in_cell_start[L.win_cell_lo] - L.win_base), w1 = (int)(L.in_cell_start[L.win_cell_hi] - L.win_base);",
                "        if (L.win_mode == 1) { t += w0; if (t >= w1) return; }",
                "        else if (t >= w0) t = w1 + (t - w0);",
                "    }"]
    out += ["    if (t >= (L.d_n ? *L.d_n : L.n)) return;",
            "    const int index = L.order ? (int)__ldg(L.order + t) : t;",
            f"    xmachine_memory_{A}_list* __restrict__ agents = (xmachine_memory_{A}_list*)L.agents;",
            f"    xmachine_memory_{A} agent;"]
    def mirror_branch(indent: str) -> List[str]:
        # cell order, straight behind the message's producer: sorted message t is this agent's own, and it carries these
        # variables unchanged (codegen.message_mirror) -- one coalesced record load instead of a random gather per variable
        M = f.input_message
        lines = [f"{indent}const xmachine_message_{M} fgpu_self = fgpu::load_message((const xmachine_message_{M}*)L.in_recs + ((size_t)L.mirror_base + (size_t)t));"]
        for v in a.variables:
            if v.name in mirror:
                lines.append(f"{indent}agent.{v.name} = fgpu_self.{mirror[v.name]};")
            elif v.array_length:
                lines.append(f"{indent}agent.{v.name} = &(agents->{v.name}[index]);")
            else:
                lines.append(f"{indent}agent.{v.name} = agents->{v.name}[index];")
        return lines
    if carry and f.input_message:
        out.append(f"    if (L.carry_in) {{   /* cell order: the producer's packed records, moved by the message build */")
        out.append(f"        constexpr int CQ = (int)(sizeof(fgpu_carry_{A}) / 16);")
        out.append(f"        union {{ fgpu_carry_{A} c; uint4 q[CQ]; }} r;")
        out.append(f"        for (int _q = 0; _q < CQ; ++_q) r.q[_q] = __ldg((const uint4*)L.carry_in + (size_t)t * CQ + _q);")
        for v in carry:
            out.append(f"        agent.{v.name} = r.c.{v.name};")
        if mirror:
            out.append("    } else if (L.mirror) {")
            out += mirror_branch("        ")
        out.append("    } else {")
        for v in a.variables:
            out.append(f"        agent.{v.name} = agents->{v.name}[index];")
        out.append("    }")
    elif mirror:
        out.append("    if (L.mirror) {")
        out += mirror_branch("        ")
        out.append("    } else {")
        for v in a.variables:
            if v.array_length:
                out.append(f"        agent.{v.name} = &(agents->{v.name}[index]);")
            else:
                out.append(f"        agent.{v.name} = agents->{v.name}[index];")
        out.append("    }")
    else:
        for v in a.variables:
            if v.array_length:
                out.append(f"    agent.{v.name} = &(agents->{v.name}[index]);")
            else:
                out.append(f"    agent.{v.name} = agents->{v.name}[index];")
    call = ["&agent"]
    if f.xagent_output:
        out.append("    fgpu::NewAgentProxy newp; newp.list = L.new_agents; newp.index = index;")
        call.append(f"(xmachine_memory_{f.xagent_output}_list*)(void*)&newp")
    if f.input_message:
        M = f.input_message
        out.append(f"    xmachine_message_{M}_list inp;")
        out.append(f"    inp.in_recs = (const xmachine_message_{M}*)L.in_recs; inp.cell_start = L.in_cell_start; inp.in_count = L.d_in_count ? *L.d_in_count : L.in_count;")
        out.append("    inp.out_recs = nullptr; inp.out_keys = nullptr; inp.out_flags = nullptr; inp.out_index = 0; inp.end = nullptr;")
        out.append("    inp.step = 0; inp.cx = 0; inp.cy = 0; inp.cz = 0; inp.mode = fgpu::WALK_CELL; inp.cnt = 0; inp.nlo = 0; inp.nhi = 0;")
        call.append("&inp")
        if model.message(M).partitioning == "spatial":
            call.append(f"(xmachine_message_{M}_PBM*)nullptr")
    if f.output_message:
        M = f.output_message
        if f.input_message == M:
            raise ModelError(f"function '{f.name}' reads and writes message '{M}'")
        out.append(f"    xmachine_message_{M}_list outp;")
        out.append("    outp.in_recs = nullptr; outp.cell_start = nullptr; outp.in_count = 0; outp.end = nullptr;")
        out.append("    outp.step = 0; outp.cx = 0; outp.cy = 0; outp.cz = 0; outp.mode = fgpu::WALK_CELL; outp.cnt = 0; outp.nlo = 0; outp.nhi = 0;")
        out.append(f"    outp.out_recs = (xmachine_message_{M}*)L.out_recs; outp.out_keys = L.out_keys; outp.out_flags = L.out_flags; outp.out_index = L.out_base + index;")
        call.append("&outp")
    if f.rng:
        out.append("    RNG_rand48 rng; rng.Ax = L.Ax; rng.Ay = L.Ay; rng.Cx = L.Cx; rng.Cy = L.Cy;")
        out.append("    { const uint2 s = ((const uint2*)L.seeds)[index]; rng.sx = s.x; rng.sy = s.y; }")
        call.append("&rng")
    out.append(f"    const int dead = !{f.name}({', '.join(call)});")
    if f.reallocate:
        out.append("    agents->_scan_input[index] = dead;   /* krn:1381: flag 1 = keep */")
    else:
        out.append("    (void)dead;")
    for v in a.variables:
        if not v.array_length and v.name in written:
            out.append(f"    agents->{v.name}[index] = agent.{v.name};")
    if f.rng:
        out.append("    ((uint2*)L.seeds)[index] = make_uint2(rng.sx, rng.sy);")
    if carry and f.output_message:
        out.append(f"    if (L.carry_out) {{   /* the variables as this function leaves them, for the consumer of its message */")
        out.append(f"        constexpr int CQ = (int)(sizeof(fgpu_carry_{A}) / 16);")
        out.append(f"        union {{ fgpu_carry_{A} c; uint4 q[CQ]; }} r;")
        out.append(f"        r.q[CQ - 1] = make_uint4(0u, 0u, 0u, 0u);")
        for v in carry:
            out.append(f"        r.c.{v.name} = agent.{v.name};")
        out.append(f"        for (int _q = 0; _q < CQ; ++_q) ((uint4*)L.carry_out)[(size_t)index * CQ + _q] = r.q[_q];")
        out.append("    }")
    out.append("}")
    out.append(f"""static void fgpu_launch_{f.name}(const fgpu_launch* a, void* stream){{
    const int threads = a->win_mode ? a->launch_n : a->n;
    if (threads <= 0) return;
    const int grid = (threads + {bs} - 1) / {bs};
    cudaLaunchConfig_t cfg = {{}};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3({bs}); cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = a->pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, fgpu_k_{f.name}, *a);
}}""")
    return "\n".join(out) + "\n"


def _cuda_filter(model: Model, f: Function) -> str:
    a = model.agent(f.agent)
    A = a.name
    if f.condition is not None:
        expr = f.condition.c_expr("currentState->{0}[index]")
        copies = "\n".join(
            (f"        for (int i = 0; i < {v.array_length}; i++) nextState->{v.name}[(i*xmachine_memory_{A}_MAX)+index] = currentState->{v.name}[(i*xmachine_memory_{A}_MAX)+index];"
             if v.array_length else f"        nextState->{v.name}[index] = currentState->{v.name}[index];")
            for v in a.variables)
        return f"""/* {f.name}_function_filter (krn:162-184); also performs the two
 * reset_{A}_scan_input launches of simulation.xslt:1447-1452 by writing both flags */
__global__ void __launch_bounds__(256) fgpu_filter_{f.name}(const fgpu_filter F)
{{
    const int index = blockIdx.x * blockDim.x + threadIdx.x;
    if (index >= F.n) return;
    xmachine_memory_{A}_list* __restrict__ currentState = (xmachine_memory_{A}_list*)F.current;
    xmachine_memory_{A}_list* __restrict__ nextState = (xmachine_memory_{A}_list*)F.working;
    if {expr}
    {{
{copies}
        nextState->_scan_input[index] = 1;
        currentState->_scan_input[index] = 0;
    }}
    else
    {{
        nextState->_scan_input[index] = 0;
        currentState->_scan_input[index] = 1;
    }}
}}
static void fgpu_launch_filter_{f.name}(const fgpu_filter* a, void* stream){{
    if (a->n <= 0) return;
    fgpu_filter_{f.name}<<<(a->n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*a);
}}
"""
    expr = f.global_condition.c_expr("currentState->{0}[index]")
    return f"""/* {f.name}_function_filter, global condition (krn:192-210) */
__global__ void __launch_bounds__(256) fgpu_filter_{f.name}(const fgpu_filter F)
{{
    const int index = blockIdx.x * blockDim.x + threadIdx.x;
    if (index >= F.n) return;
    xmachine_memory_{A}_list* __restrict__ currentState = (xmachine_memory_{A}_list*)F.current;
    if {expr}
        currentState->_scan_input[index] = 1;
    else
        currentState->_scan_input[index] = 0;
}}
static void fgpu_launch_filter_{f.name}(const fgpu_filter* a, void* stream){{
    if (a->n <= 0) return;
    fgpu_filter_{f.name}<<<(a->n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*a);
}}
"""


def _descriptor_tables(model: Model, mode: str, launch_prefix: str, sources: Optional[List[str]] = None) -> str:
    """The fgpu_model tables (shared by both back ends; the oracle passes its
    own thunks in place of kernel launchers)."""
    out = []
    aidx = {a.name: i for i, a in enumerate(model.agents)}
    midx = {m.name: i for i, m in enumerate(model.messages)}
    for a in model.agents:
        rows = []
        for v in a.variables:
            dv = v.default_literal()
            rows.append(f'    {{"{v.name}", "{v.type}", {v.size}, {1 if v.type in ("float", "double") else 0}, '
                        f'offsetof(xmachine_memory_{a.name}_list, {v.name}), (double)({dv}), {max(1, v.array_length)}}},')
        out.append(f"static const fgpu_var fgpu_vars_agent_{a.name}[] = {{\n" + "\n".join(rows) + "\n};")
        states = ", ".join(f'"{s}"' for s in a.states)
        out.append(f"static const char* const fgpu_states_{a.name}[] = {{{states}}};")
    for m in model.messages:
        rows = []
        for v in m.variables:
            rows.append(f'    {{"{v.name}", "{v.type}", {v.size}, {1 if v.type in ("float", "double") else 0}, '
                        f'offsetof(xmachine_message_{m.name}{"_list" if mode == "oracle" else ""}, {v.name}), 0.0, 1}},')
        out.append(f"static const fgpu_var fgpu_vars_message_{m.name}[] = {{\n" + "\n".join(rows) + "\n};")
    rows = []
    for a in model.agents:
        rows.append(f'    {{"{a.name}", {a.buffer_size}, {len(a.variables)}, fgpu_vars_agent_{a.name}, {len(a.states)}, '
                    f'fgpu_states_{a.name}, {a.states.index(a.initial_state)}, sizeof(xmachine_memory_{a.name}_list), '
                    f'offsetof(xmachine_memory_{a.name}_list, _position), offsetof(xmachine_memory_{a.name}_list, _scan_input), '
                    + (f'(int)(sizeof(fgpu_carry_{a.name}) / 16)' if mode == "cuda" and _carry_vars(a) else '0') + '},')
    out.append("static const fgpu_agent fgpu_agents[] = {\n" + "\n".join(rows) + "\n};")
    rows = []
    for m in model.messages:
        quads = f"(int)(sizeof(xmachine_message_{m.name}) / 16)"
        names = [v.name for v in m.variables]
        if m.partitioning == "spatial":
            dx, dy, dz = m.partition_dims()
            mn = ", ".join(f"(float){x}" for x in m.min_literals)
            mx = ", ".join(f"(float){x}" for x in m.max_literals)
            rows.append(f'    {{"{m.name}", {m.buffer_size}, {len(m.variables)}, fgpu_vars_message_{m.name}, {quads}, FGPU_PART_SPATIAL, '
                        f'(float){m.radius_literal}, {{{mn}}}, {{{mx}}}, {{{dx}, {dy}, {dz}}}, {m.grid_size()}, {1 if m.is_2d() else 0}, '
                        f'{m.radix_bits()}, {names.index("x")}, {names.index("y")}, {names.index("z")}}},')
        else:
            rows.append(f'    {{"{m.name}", {m.buffer_size}, {len(m.variables)}, fgpu_vars_message_{m.name}, {quads}, FGPU_PART_NONE, '
                        f'0.0f, {{0,0,0}}, {{0,0,0}}, {{1,1,1}}, 1, 0, 1, -1, -1, -1}},')
    if rows:
        out.append("static const fgpu_message fgpu_messages[] = {\n" + "\n".join(rows) + "\n};")
    else:
        out.append("static const fgpu_message* const fgpu_messages = nullptr;")
    funcs = model.all_functions()
    fidx = {f.name: i for i, f in enumerate(funcs)}
    rows = []
    for f in funcs:
        a = model.agent(f.agent)
        cond = "FGPU_COND_FUNCTION" if f.condition is not None else ("FGPU_COND_GLOBAL" if f.global_condition is not None else "FGPU_COND_NONE")
        # the literal test of simulation.xslt:1945 (note gpu:reallocate='false' is a string compare)
        swappable = int(f.current_state == f.next_state and f.condition is None and f.global_condition is None
                        and f.reallocate_literal == "false" and f.xagent_output is None)
        filt = f"{launch_prefix}filter_{f.name}" if cond != "FGPU_COND_NONE" else "nullptr"
        rows.append(
            f'    {{"{f.name}", {aidx[f.agent]}, {a.states.index(f.current_state)}, {a.states.index(f.next_state)}, '
            f'{midx[f.input_message] if f.input_message else -1}, {midx[f.output_message] if f.output_message else -1}, '
            f'{"FGPU_OUT_OPTIONAL" if f.output_type == "optional_message" else "FGPU_OUT_SINGLE"}, '
            f'{aidx[f.xagent_output] if f.xagent_output else -1}, '
            f'{model.agent(f.xagent_output).states.index(f.xagent_output_state) if f.xagent_output else -1}, '
            f'{int(f.reallocate)}, {swappable}, {int(f.rng)}, {cond}, {f.gc_max_iterations}, {int(f.gc_must_evaluate_to)}, '
            f'{launch_prefix}{f.name}, {filt}, {128 if f.input_message else 256}, '
            f'{int(bool(written_variables(f, a, sources or []) & {"x", "y", "z"}))}, '
            # the wrapper can take agent variables from the agent's own message when this function produced it
            # (cuda back end only: the oracle reads the list)
            + (str(fidx[mirror_producer(model, f, sources or [])[0].name]) if mode == "cuda" and mirror_producer(model, f, sources or [])[0] else "-1") + '},')
    out.append("static const fgpu_function fgpu_functions[] = {\n" + "\n".join(rows) + "\n};")
    lrows = []
    for i, layer in enumerate(model.layers):
        out.append(f"static const int fgpu_layer_{i}[] = {{{', '.join(str(fidx[n]) for n in layer)}}};")
        lrows.append(f"    {{{len(layer)}, fgpu_layer_{i}}},")
    out.append("static const fgpu_layer fgpu_layers[] = {\n" + "\n".join(lrows) + "\n};")
    return "\n".join(out) + "\n"


def _constants_code(model: Model, cuda: bool) -> str:
    out = []
    rows = []
    for c in model.constants:
        dim = f"[{c.array_length}]" if c.array_length else ""
        n = max(1, c.array_length)
        out.append(f"{c.type} h_env_{c.name}{dim};")
        if cuda:
            out.append(f"void set_{c.name}({c.type}* h_{c.name}){{\n"
                       f"    cudaMemcpyToSymbol({c.name}, h_{c.name}, sizeof({c.type})*{n});\n"
                       f"    memcpy(&h_env_{c.name}, h_{c.name}, sizeof({c.type})*{n});\n}}")
        else:
            out.append(f"void set_{c.name}({c.type}* h_{c.name}){{\n"
                       f"    memcpy(&{c.name}, h_{c.name}, sizeof({c.type})*{n});\n"
                       f"    memcpy(&h_env_{c.name}, h_{c.name}, sizeof({c.type})*{n});\n}}")
        out.append(f"const {c.type}* get_{c.name}(){{ return {'h_env_' + c.name if c.array_length else '&h_env_' + c.name}; }}")
        out.append(f"static void fgpu_cset_{c.name}(const void* v){{ set_{c.name}(({c.type}*)v); }}")
        out.append(f"static const void* fgpu_cget_{c.name}(void){{ return (const void*)get_{c.name}(); }}")
        rows.append(f'    {{"{c.name}", "{c.type}", (int)sizeof({c.type}), {n}, fgpu_cset_{c.name}, fgpu_cget_{c.name}}},')
    # initEnvVars (io.xslt:201-209): T t = (T)<literal>; set_<c>(&t)
    out.append("static void fgpu_init_constants(void){")
    for c in model.constants:
        if c.default is not None and c.default != "" and not c.array_length:
            out.append(f"    {{ {c.type} t = ({c.type}){c.default}; set_{c.name}(&t); }}")
        elif not c.array_length:
            out.append(f"    {{ {c.type} t = ({c.type})0; set_{c.name}(&t); }}")
    out.append("}")
    if rows:
        out.append("static const fgpu_constant fgpu_constants[] = {\n" + "\n".join(rows) + "\n};")
    else:
        out.append("static const fgpu_constant* const fgpu_constants = nullptr;")
    return "\n".join(out) + "\n"


def _face2_shim(model: Model, cuda: bool) -> str:
    """Face-2 names of the reference (header.xslt:531-645) on top of the C ABI, preserving
    the reference's error behaviour (message + exit(), simulation.xslt:198-206).  The oracle
    build gets no-op bind only (its driver is the test harness)."""
    if not cuda:
        return "static void fgpu_plugin_bind(void*){}\n"
    out = ['#include "fgpu.h"',
           "static fgpu_sim* fgpu_cur = nullptr;",

           "static void fgpu_plugin_bind(void* sim){ fgpu_cur = (fgpu_sim*)sim; }",
           "static void fgpu_update_bounds();",
           'static void fgpu_die(){ fprintf(stderr, "%s\\n", fgpu_last_error()); fflush(stderr); exit(EXIT_FAILURE); }',
           "unsigned int getIterationNumber(){ return fgpu_sim_iteration(fgpu_cur); }",
           "void set_exit_early(){ fgpu_set_exit_early(1); }",
           "bool get_exit_early(){ return fgpu_get_exit_early() != 0; }",
           "const char* getOutputDir(){ return fgpu_get_output_dir(); }",
           """void initialise(char * inputfile){
    /* simulation.xslt:472-747: allocate, read the initial states, seed rand48, run init functions */
    const char* dev = getenv("FGPU_DEVICE");
    fgpu_cur = fgpu_sim_create(fgpu_model_get(), dev ? atoi(dev) : 0);
    if (!fgpu_cur) fgpu_die();
    if (fgpu_sim_load_states(fgpu_cur, inputfile) != FGPU_OK) fgpu_die();
    fgpu_update_bounds();
    if (fgpu_sim_run_init_functions(fgpu_cur) != FGPU_OK) fgpu_die();
}
void singleIteration(){ if (fgpu_sim_step(fgpu_cur, 1) != FGPU_OK) fgpu_die(); }
void cleanup(){ fgpu_sim_run_exit_functions(fgpu_cur); fgpu_sim_destroy(fgpu_cur); fgpu_cur = nullptr; }"""]
    for ai, a in enumerate(model.agents):
        out.append(f"int get_agent_{a.name}_MAX_count(){{ return xmachine_memory_{a.name}_MAX; }}")
        for si, st in enumerate(a.states):
            out.append(f"int get_agent_{a.name}_{st}_count(){{ return fgpu_sim_agent_count(fgpu_cur, {ai}, {si}); }}")
            out.append(f"void reset_{a.name}_{st}_count(){{ fgpu_sim_set_agents(fgpu_cur, {ai}, {si}, 0, nullptr); }}")
            out.append(f"xmachine_memory_{a.name}_list* get_device_{a.name}_{st}_agents(){{ return (xmachine_memory_{a.name}_list*)fgpu_sim_device_agents(fgpu_cur, {ai}, {si}); }}")
            out.append(f"void sort_{a.name}s_{st}(void (*generate_key_value_pairs)(unsigned int* keys, unsigned int* values, xmachine_memory_{a.name}_list* agents)){{ "
                       f"if (fgpu_sim_sort_agents(fgpu_cur, {ai}, {si}, (const void*)generate_key_value_pairs) != FGPU_OK) fgpu_die(); }}")
            for vi, v in enumerate(a.variables):
                if not v.array_length:
                    out.append(f"{v.type} get_{a.name}_{st}_variable_{v.name}(unsigned int index){{ {v.type} t = 0; "
                               f"if (fgpu_sim_read_agent_value(fgpu_cur, {ai}, {si}, {vi}, index, &t) != FGPU_OK) fgpu_die(); return t; }}")
                else:   # simulation.xslt:1093-1140
                    out.append(f"{v.type} get_{a.name}_{st}_variable_{v.name}(unsigned int index, unsigned int element){{ {v.type} t = 0; "
                               f"if (fgpu_sim_read_agent_element(fgpu_cur, {ai}, {si}, {vi}, index, element, &t) != FGPU_OK) fgpu_die(); return t; }}")
    out += _host_list_defs(model)
    out.append("static int fgpu_reduce_cur(int a, int s, int v, int op, const void* cv, void* r){ return fgpu_sim_reduce_agent_var(fgpu_cur, a, s, v, op, cv, r); }")
    out += _reduction_defs(model, "fgpu_reduce_cur")
    out.append("static int fgpu_add_cur(int a, int s, int n, const void* const* cols){ return fgpu_sim_add_agents(fgpu_cur, a, s, n, cols); }")
    out.append("static int fgpu_count_cur(int a, int s){ return fgpu_sim_agent_count(fgpu_cur, a, s); }")
    out += _host_creation_defs(model, "fgpu_add_cur", "fgpu_count_cur")
    out.append(_id_defs(model, cuda=True))
    return "\n".join(out) + "\n"


def _model_struct(model: Model, cuda: bool = True) -> str:
    def fnlist(kind, names):
        if not names:
            return f"static void (*const *fgpu_{kind}_functions)(void) = nullptr;"
        return f"static void (*const fgpu_{kind}_functions[])(void) = {{{', '.join(names)}}};"
    def namelist(kind, names):
        if not names:
            return f"static const char* const* fgpu_{kind}_function_names = nullptr;"
        return f"static const char* const fgpu_{kind}_function_names[] = {{{', '.join(chr(34) + n + chr(34) for n in names)}}};"
    out = [fnlist("init", model.init_functions), fnlist("step", model.step_functions), fnlist("exit", model.exit_functions),
           namelist("init", model.init_functions), namelist("step", model.step_functions), namelist("exit", model.exit_functions)]
    safe = model.name.replace('"', "'")
    out.append(f"""static const fgpu_model fgpu_the_model = {{
    FGPU_MODEL_ABI_VERSION, "{safe}", {model.buffer_size_max},
    {len(model.agents)}, fgpu_agents,
    {len(model.messages)}, fgpu_messages,
    {len(model.all_functions())}, fgpu_functions,
    {len(model.layers)}, fgpu_layers,
    {len(model.constants)}, fgpu_constants,
    fgpu_init_constants,
    {len(model.init_functions)}, fgpu_init_functions,
    {len(model.step_functions)}, fgpu_step_functions,
    {len(model.exit_functions)}, fgpu_exit_functions,
    fgpu_plugin_bind,
    fgpu_init_function_names, fgpu_step_function_names, fgpu_exit_function_names,
    {'fgpu_ids_after_load, fgpu_ids_to_host, fgpu_ids_to_device' if cuda else 'nullptr, nullptr, nullptr'}
}};
extern "C" const fgpu_model* fgpu_model_get(void){{ return &fgpu_the_model; }}
""")
    return "\n".join(out)


def emit_cuda_glue(model: Model, function_files: List[str], include_files: Optional[List[str]] = None) -> str:
    """`function_files` are the author's scripts (analysed for write sets etc.); `include_files` the files
    actually compiled in their place (the loop-rewritten copies), default the scripts themselves."""
    sources = [open(p, "r", errors="replace").read() for p in function_files]
    out = ["/* generated by ev_diffusion_b200.codegen -- model plugin for sm_100a */",
           "#include <cstddef>\n#include <cstring>\n#include <cstdio>\n#include <cmath>\n#include <cstdlib>\n#include <vector>",
           '#include "header.h"',
           '#include "fgpu_model.h"']
    for p in (include_files or function_files):
        out.append(f'#include "{os.path.abspath(p)}"')
    for a in model.agents:
        out.append(_cuda_carry_struct(a))
    for f in model.all_functions():
        a = model.agent(f.agent)
        out.append(_cuda_kernel(model, f, written_variables(f, a, sources), mirror_producer(model, f, sources)[1]))
        if f.condition is not None or f.global_condition is not None:
            out.append(_cuda_filter(model, f))
    out.append(_constants_code(model, cuda=True))
    out.append(_descriptor_tables(model, "cuda", "fgpu_launch_", sources))
    out.append('extern "C" const fgpu_model* fgpu_model_get(void);')
    out.append(_face2_shim(model, cuda=True))
    out.append(_model_struct(model))
    return "\n".join(out) + "\n"


def emit_cuda(model: Model, function_files: List[str], outdir: str, rewrite_loops: bool = True,
              loop_unroll: Optional[int] = None) -> Dict[str, str]:
    """Writes header.h, model_glue.cu and -- unless `rewrite_loops` is off -- a copy of every script with
    its canonical message loops restructured (looprewrite.py), which is what model_glue.cu then includes.
    `loops.txt` lists every message loop found and whether it was rewritten."""
    os.makedirs(outdir, exist_ok=True)
    files: Dict[str, str] = {}
    include_files = None
    if rewrite_loops:
        include_files, log = [], []
        for i, p in enumerate(function_files):
            text, reports = rewrite_message_loops(open(p, "r", errors="replace").read(), [m.name for m in model.messages],
                                                  unroll=loop_unroll)
            name = f"rewritten_{i}_{os.path.basename(p)}"
            files[name] = f'#line 1 "{os.path.abspath(p)}"\n' + text
            include_files.append(os.path.join(outdir, name))
            for r in reports:
                log.append(f"{p}:{r.line}: message {r.message}: " + ("rewritten" if r.rewritten else f"generic get_next ({r.reason})"))
        files["loops.txt"] = "\n".join(log) + "\n"
    files["header.h"] = emit_cuda_header(model)
    files["model_glue.cu"] = emit_cuda_glue(model, function_files, include_files)
    for name, text in files.items():
        path = os.path.join(outdir, name)
        if not os.path.exists(path) or open(path).read() != text:
            with open(path, "w") as fh:
                fh.write(text)
    return {k: os.path.join(outdir, k) for k in files}
