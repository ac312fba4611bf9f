"""ORACLE glue generator.  TEST INFRASTRUCTURE ONLY -- see oracle/oracle_rt.hpp.

Emits, for one XMML model, a host-mode ``header.h`` whose structs have the
reference's own layouts (``header.xslt:151-225,334-338``) and an
``oracle_glue.cpp`` that restates the device layer of
``FLAMEGPU_kernals.xslt`` for one sequential "thread" at a time:

  add_<M>_message            krn:382-404
  get_first/next_<M>_message krn:1037-1160 (deterministic-sort PBM form), krn:105-154
  add_<A>_agent              krn:287-314
  rnd<>                      krn:1488-1547
  GPUFLAME_<f>               krn:1320-1387 (gather all vars, call, scatter all vars)
  <f>_function_filter        krn:162-210

The agent scripts (functions.c) are compiled unchanged with
``__FLAME_GPU_FUNC__`` empty.  The "thread index" of the reference
(blockIdx.x*blockDim.x+threadIdx.x == list index for continuous agents,
krn:1326) is a thread_local integer set by the agent loop.
"""
from __future__ import annotations

import os
import subprocess
import sys
from typing import List

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from ev_diffusion_b200 import codegen, xmml  # noqa: E402  (the oracle may read product helpers; never the reverse)
from ev_diffusion_b200.xmml import Function, Message, Model  # noqa: E402

OUT = os.path.join(HERE, "_build")
LIBDIR = os.path.join(HERE, "_ref")


def _message_structs(model: Model) -> str:
    out = []
    for m in model.messages:
        out.append(f"struct xmachine_message_{m.name}\n{{")
        if m.partitioning == "spatial":
            # hdr:170-173
            out.append("    glm::ivec3 _relative_cell;\n    int _cell_index_max;\n    glm::ivec3 _agent_grid_cell;\n    int _cell_index;")
        else:
            out.append("    int _position;")
        for v in m.variables:
            out.append(f"    {v.type} {v.name};")
        out.append("};")
        out.append(f"struct xmachine_message_{m.name}_list\n{{")
        out.append(f"    int _position [xmachine_message_{m.name}_MAX];")
        out.append(f"    int _scan_input [xmachine_message_{m.name}_MAX];")
        for v in m.variables:
            out.append(f"    {v.type} {v.name} [xmachine_message_{m.name}_MAX];")
        out.append("};")
        if m.partitioning == "spatial":
            out.append(f"struct xmachine_message_{m.name}_PBM\n{{\n    int start[xmachine_message_{m.name}_grid_size];\n"
                       f"    int end_or_count[xmachine_message_{m.name}_grid_size];\n}};")
    return "\n".join(out) + "\n"


def _message_api(m: Message) -> str:
    n = m.name
    S = f"xmachine_message_{n}"
    out = [f"int d_message_{n}_count = 0;", f"int d_message_{n}_output_type = 0;"]
    # add_<M>_message, krn:382-404
    out.append(f"void add_{n}_message({S}_list* messages, {codegen._msg_args(m)}){{")
    out.append(f"    int index = orc_tid + d_message_{n}_count;")
    out.append("    int _position; int _scan_input;")
    out.append(f"    if (d_message_{n}_output_type == single_message){{ _position = index; _scan_input = 0; }}")
    out.append("    else { _position = 0; _scan_input = 1; }")
    out.append("    messages->_scan_input[index] = _scan_input;\n    messages->_position[index] = _position;")
    for v in m.variables:
        out.append(f"    messages->{v.name}[index] = {v.name};")
    out.append("}")
    if m.partitioning != "spatial":
        # brute force: every message, in list order 0..N-1.  krn:445-520 visits all messages starting at the calling
        # block's own tile and wrapping round, so the reference's order is 0..N-1 exactly when one block holds the
        # whole list; with smaller blocks it is a per-block rotation of it (tests/test_oracle_vs_reference.py)
        out.append(f"""static thread_local {S} orc_sm_{n};
static {S}* orc_load_{n}({S}_list* messages, int i){{
    if (i >= d_message_{n}_count) return nullptr;
    orc_sm_{n}._position = i;""")
        for v in m.variables:
            out.append(f"    orc_sm_{n}.{v.name} = messages->{v.name}[i];")
        out.append(f"    return &orc_sm_{n};\n}}")
        out.append(f"""/* get_first / get_next, krn:445-520.  The reference starts at the calling block's own tile and walks round:
 * block b of B threads meets messages (b*B) % wrap ... count-1, 0 ... start-1, with wrap = ceil(count / B) * B (computed
 * in float, as there).  B is the launch's block size: orc_brute_force_block, capped by the number of agents launched
 * like the occupancy call's limit (sim:1601-1606); 0 = one block holds everything = plain list order. */
static thread_local int orc_bf_start_{n}, orc_bf_wrap_{n};
{S}* get_first_{n}_message({S}_list* messages){{
    if (d_message_{n}_count == 0) return nullptr;
    int B = orc_brute_force_block;
    if (B <= 0 || B > orc_launch_threads) B = orc_launch_threads > 0 ? orc_launch_threads : 1;
    orc_bf_wrap_{n} = (int)(ceil((float)d_message_{n}_count / B) * B);
    orc_bf_start_{n} = ((orc_tid / B) * B) % orc_bf_wrap_{n};
    return orc_load_{n}(messages, orc_bf_start_{n});
}}
{S}* get_next_{n}_message({S}* current, {S}_list* messages){{
    int i = (current->_position + 1) % orc_bf_wrap_{n};
    if (i >= d_message_{n}_count) i = 0;
    if (i == orc_bf_start_{n}) return nullptr;
    return orc_load_{n}(messages, i);
}}""")
        return "\n".join(out) + "\n"
    dx, dy, dz = m.partition_dims()
    mn = ", ".join(f"(float){x}" for x in m.min_literals)
    mx = ", ".join(f"(float){x}" for x in m.max_literals)
    nextcell = "orc_next_cell2D" if m.is_2d() else "orc_next_cell3D"
    out.append(f"""static const glm::vec3 d_message_{n}_min_bounds = glm::vec3({mn});
static const glm::vec3 d_message_{n}_max_bounds = glm::vec3({mx});
static const glm::ivec3 d_message_{n}_partitionDim = glm::ivec3({dx}, {dy}, {dz});
/* message_{n}_grid_position, krn:860-871 */
static glm::ivec3 message_{n}_grid_position(glm::vec3 position)
{{
    glm::ivec3 gridPos;
    gridPos.x = floor((position.x - d_message_{n}_min_bounds.x) * (float)d_message_{n}_partitionDim.x / (d_message_{n}_max_bounds.x - d_message_{n}_min_bounds.x));
    gridPos.y = floor((position.y - d_message_{n}_min_bounds.y) * (float)d_message_{n}_partitionDim.y / (d_message_{n}_max_bounds.y - d_message_{n}_min_bounds.y));
    gridPos.z = floor((position.z - d_message_{n}_min_bounds.z) * (float)d_message_{n}_partitionDim.z / (d_message_{n}_max_bounds.z - d_message_{n}_min_bounds.z));
    return gridPos;
}}
/* message_{n}_hash, krn:877-889 */
static unsigned int message_{n}_hash(glm::ivec3 gridPos)
{{
    gridPos.x = (gridPos.x<0)? d_message_{n}_partitionDim.x-1: gridPos.x;
    gridPos.x = (gridPos.x>=d_message_{n}_partitionDim.x)? 0 : gridPos.x;
    gridPos.y = (gridPos.y<0)? d_message_{n}_partitionDim.y-1 : gridPos.y;
    gridPos.y = (gridPos.y>=d_message_{n}_partitionDim.y)? 0 : gridPos.y;
    gridPos.z = (gridPos.z<0)? d_message_{n}_partitionDim.z-1: gridPos.z;
    gridPos.z = (gridPos.z>=d_message_{n}_partitionDim.z)? 0 : gridPos.z;
    return ((gridPos.z * d_message_{n}_partitionDim.y) * d_message_{n}_partitionDim.x) + (gridPos.y * d_message_{n}_partitionDim.x) + gridPos.x;
}}
static thread_local {S} orc_sm_{n};   /* the thread's shared-memory slot, krn:1100-1103 */
/* load_next_{n}_message, krn:1037-1106, non-FAST_ATOMIC_SORTING branch */
static bool load_next_{n}_message({S}_list* messages, {S}_PBM* partition_matrix, glm::ivec3 relative_cell, int cell_index_max, glm::ivec3 agent_grid_cell, int cell_index)
{{
    int move_cell = true;
    cell_index ++;
    if(cell_index < cell_index_max)
        move_cell = false;
    while(move_cell)
    {{
        if ({nextcell}(&relative_cell))
        {{
            glm::ivec3 next_cell_position = agent_grid_cell + relative_cell;
            int next_cell_hash = message_{n}_hash(next_cell_position);
            int cell_index_min = partition_matrix->start[next_cell_hash];
            cell_index_max = partition_matrix->end_or_count[next_cell_hash];
            if ((unsigned int)cell_index_min != 0xffffffffu)
            {{
                cell_index = cell_index_min;
                move_cell = false;
            }}
        }}
        else
        {{
            return false;
        }}
    }}
    {S} temp_message;
    temp_message._relative_cell = relative_cell;
    temp_message._cell_index_max = cell_index_max;
    temp_message._cell_index = cell_index;
    temp_message._agent_grid_cell = agent_grid_cell;""")
    for v in m.variables:
        out.append(f"    temp_message.{v.name} = messages->{v.name}[cell_index];")
    out.append(f"""    orc_sm_{n} = temp_message;
    return true;
}}
/* get_first_{n}_message, krn:1111-1136 */
{S}* get_first_{n}_message({S}_list* messages, {S}_PBM* partition_matrix, float x, float y, float z){{
    if(d_message_{n}_count == 0){{
        return nullptr;
    }}
    glm::ivec3 relative_cell = glm::ivec3(-2, -1, -1);
    int cell_index_max = 0;
    int cell_index = 0;
    glm::vec3 position = glm::vec3(x, y, z);
    glm::ivec3 agent_grid_cell = message_{n}_grid_position(position);
    if (load_next_{n}_message(messages, partition_matrix, relative_cell, cell_index_max, agent_grid_cell, cell_index))
        return &orc_sm_{n};
    return nullptr;
}}
/* get_next_{n}_message, krn:1141-1160 */
{S}* get_next_{n}_message({S}* message, {S}_list* messages, {S}_PBM* partition_matrix){{
    if(d_message_{n}_count == 0){{
        return nullptr;
    }}
    if (load_next_{n}_message(messages, partition_matrix, message->_relative_cell, message->_cell_index_max, message->_agent_grid_cell, message->_cell_index))
        return &orc_sm_{n};
    return nullptr;
}}""")
    return "\n".join(out) + "\n"


def emit_header(model: Model) -> str:
    out = ["/* generated by oracle/gen_oracle.py -- HOST-mode model header (oracle, test infrastructure) */",
           "#ifndef __HEADER\n#define __HEADER",
           "#include <cmath>\n#include <cstdlib>\n#include <cstring>\n#include <cstdio>\n#include <algorithm>",
           '#include "fgpu_glm.h"',
           "#define __FLAME_GPU_FUNC__\n#define __FLAME_GPU_INIT_FUNC__\n#define __FLAME_GPU_STEP_FUNC__\n#define __FLAME_GPU_EXIT_FUNC__\n#define __FLAME_GPU_HOST_FUNC__",
           "#define FGPU_ORACLE 1",
           "/* device built-ins the scripts use unqualified (SURVEY.md 7, hard part 9) */",
           "using std::max; using std::min; using std::abs; using std::pow; using std::sqrt; using std::floor; using std::ceil;",
           "using std::sin; using std::cos; using std::exp; using std::log; using std::fabs; using std::fmod;"]
    out.append(codegen._common_defs(model))
    out.append(codegen._agent_structs(model, ""))
    out.append(_message_structs(model))
    out.append("struct RNG_rand48\n{\n    glm::uvec2 A, C;\n    glm::uvec2 seeds[buffer_size_MAX];\n};")
    out.append("extern thread_local int orc_tid;")
    out.append("extern int orc_launch_threads, orc_brute_force_block;")
    out.append("template <int AGENT_TYPE> float rnd(RNG_rand48* rand48);\nfloat rnd(RNG_rand48* rand48);")
    for m in model.messages:
        S = f"xmachine_message_{m.name}"
        out.append(f"void add_{m.name}_message({S}_list* messages, {codegen._msg_args(m)});")
        if m.partitioning == "spatial":
            out.append(f"{S}* get_first_{m.name}_message({S}_list* messages, {S}_PBM* partition_matrix, float x, float y, float z);")
            out.append(f"{S}* get_next_{m.name}_message({S}* current, {S}_list* messages, {S}_PBM* partition_matrix);")
        else:
            out.append(f"{S}* get_first_{m.name}_message({S}_list* messages);")
            out.append(f"{S}* get_next_{m.name}_message({S}* current, {S}_list* messages);")
    for a in model.agents:
        out.append(f"template <int AGENT_TYPE> void add_{a.name}_agent(xmachine_memory_{a.name}_list* agents, {codegen._agent_args(a)});")
        out.append(f"void add_{a.name}_agent(xmachine_memory_{a.name}_list* agents, {codegen._agent_args(a)});")
    for f in model.all_functions():
        out.append(codegen._function_signature(model, f) + ";")
    for c in model.constants:
        dim = f"[{c.array_length}]" if c.array_length else ""
        out.append(f"extern {c.type} {c.name}{dim};")
        out.append(f"extern {c.type} h_env_{c.name}{dim};")
        out.append(f"void set_{c.name}({c.type}* h_{c.name});")
        out.append(f"const {c.type}* get_{c.name}();")
    # host API the scripts' init/step/exit functions may call; the oracle defines the counts and the analytics
    out.append(codegen._host_api_decls(model))
    out.append(codegen._id_decls(model, cuda=False))
    out.append("#endif")
    return "\n".join(out) + "\n"


def _thunk(model: Model, f: Function) -> str:
    a = model.agent(f.agent)
    A = a.name
    out = [f"/* GPUFLAME_{f.name}, krn:1320-1387: one loop trip == one CUDA thread */",
           f"static void orc_thunk_{f.name}(const fgpu_launch* L, void*){{",
           f"    xmachine_memory_{A}_list* agents = (xmachine_memory_{A}_list*)L->agents;"]
    call = ["&agent"]
    if f.xagent_output:
        call.append(f"(xmachine_memory_{f.xagent_output}_list*)L->new_agents")
    if f.input_message:
        M = f.input_message
        out.append(f"    d_message_{M}_count = L->in_count;")
        call.append(f"(xmachine_message_{M}_list*)L->in_recs")
        if model.message(M).partitioning == "spatial":
            call.append(f"(xmachine_message_{M}_PBM*)L->in_cell_start")
    if f.output_message:
        M = f.output_message
        out.append(f"    d_message_{M}_count = L->out_base;")
        call.append(f"(xmachine_message_{M}_list*)L->out_recs")
    if f.rng:
        call.append("(RNG_rand48*)L->seeds")
    out.append("    orc_launch_threads = L->n;")
    out.append("    #pragma omp parallel for schedule(static)")
    out.append("    for (int index = 0; index < L->n; ++index){")
    out.append("        orc_tid = index;")
    out.append(f"        xmachine_memory_{A} agent;")
    for v in a.variables:
        if v.array_length:
            out.append(f"        agent.{v.name} = &(agents->{v.name}[index]);")
        else:
            out.append(f"        agent.{v.name} = agents->{v.name}[index];")
    out.append(f"        int dead = !{f.name}({', '.join(call)});")
    out.append("        agents->_scan_input[index] = dead;")
    for v in a.variables:
        if not v.array_length:
            out.append(f"        agents->{v.name}[index] = agent.{v.name};")
    out.append("    }\n}")
    return "\n".join(out) + "\n"


def _filter(model: Model, f: Function) -> str:
    a = model.agent(f.agent)
    A = a.name
    if f.condition is not None:
        expr = f.condition.c_expr("currentState->{0}[index]")
        # krn:172-173 copies `var[index]` for every variable, i.e. only element 0 of an array variable, and leaves the rest of
        # the working list's arrays as allocated (uninitialised in the reference).  Both back ends of this repository copy
        # the whole array -- what scatter/append/reorder do (krn:251-254) -- so that the result is defined.
        copies = "\n".join(
            (f"            for (int i = 0; i < {v.array_length}; i++) nextState->{v.name}[(i*xmachine_memory_{A}_MAX)+index] = currentState->{v.name}[(i*xmachine_memory_{A}_MAX)+index];"
             if v.array_length else f"            nextState->{v.name}[index] = currentState->{v.name}[index];") for v in a.variables)
        return f"""/* {f.name}_function_filter, krn:162-184 */
static void orc_filter_{f.name}(const fgpu_filter* F, void*){{
    xmachine_memory_{A}_list* currentState = (xmachine_memory_{A}_list*)F->current;
    xmachine_memory_{A}_list* nextState = (xmachine_memory_{A}_list*)F->working;
    for (int index = 0; index < F->n; ++index){{
        if {expr}
        {{
{copies}
            nextState->_scan_input[index] = 1;
        }}
        else
        {{
            currentState->_scan_input[index] = 1;
        }}
    }}
}}
"""
    expr = f.global_condition.c_expr("currentState->{0}[index]")
    return f"""/* {f.name}_function_filter (global condition), krn:192-210 */
static void orc_filter_{f.name}(const fgpu_filter* F, void*){{
    xmachine_memory_{A}_list* currentState = (xmachine_memory_{A}_list*)F->current;
    for (int index = 0; index < F->n; ++index){{
        if {expr}
            currentState->_scan_input[index] = 1;
        else
            currentState->_scan_input[index] = 0;
    }}
}}
"""


def emit_glue(model: Model, function_files: List[str]) -> str:
    out = ["/* generated by oracle/gen_oracle.py -- CPU ORACLE glue (test infrastructure) */",
           "/* every standard header the runtime needs comes BEFORE the scripts: a script may define macros",
           " * such as `length(x)` that would otherwise rewrite libstdc++ itself */",
           "#include <cstddef>\n#include <algorithm>\n#include <cmath>\n#include <cstdarg>\n#include <cstdio>\n#include <cstdlib>",
           "#include <cstring>\n#include <string>\n#include <vector>\n#include <omp.h>",
           '#include "header.h"',
           '#include "fgpu_model.h"',
           "thread_local int orc_tid = 0;",
           "int orc_launch_threads = 0;          /* threads of the agent function being run (L->n) */",
           "int orc_brute_force_block = 0;       /* block size assumed for non-partitioned message walks; 0 = one block */",
           'extern "C" void orc_set_brute_force_block(int threads_per_block){ orc_brute_force_block = threads_per_block; }',
           """/* next_cell3D / next_cell2D, krn:105-154 */
static bool orc_next_cell3D(glm::ivec3* relative_cell)
{
    if (relative_cell->x < 1) { relative_cell->x++; return true; }
    relative_cell->x = -1;
    if (relative_cell->y < 1) { relative_cell->y++; return true; }
    relative_cell->y = -1;
    if (relative_cell->z < 1) { relative_cell->z++; return true; }
    relative_cell->z = -1;
    return false;
}
static bool orc_next_cell2D(glm::ivec3* relative_cell)
{
    if (relative_cell->x < 1) { relative_cell->x++; return true; }
    relative_cell->x = -1;
    if (relative_cell->y < 1) { relative_cell->y++; return true; }
    relative_cell->y = -1;
    return false;
}
/* RNG_rand48_iterate_single, krn:1488-1514 (__umul24 / __umulhi restated on 64-bit integers) */
static glm::uvec2 RNG_rand48_iterate_single(glm::uvec2 Xn, glm::uvec2 A, glm::uvec2 C)
{
    unsigned int R0, R1;
    const unsigned long long p00 = (unsigned long long)(Xn.x & 0xFFFFFF) * (unsigned long long)(A.x & 0xFFFFFF);
    const unsigned int lo00 = (unsigned int)p00;                                   /* __umul24 */
    const unsigned int hi00 = (unsigned int)(((unsigned long long)Xn.x * (unsigned long long)A.x) >> 32); /* __umulhi */
    R0 = (lo00 & 0xFFFFFF);
    R1 = (lo00 >> 24) | (hi00 << 8);
    R0 += C.x; R1 += C.y;
    R1 += (R0 >> 24);
    R0 &= 0xFFFFFF;
    R1 += (unsigned int)((unsigned long long)(Xn.y & 0xFFFFFF) * (unsigned long long)(A.x & 0xFFFFFF));
    R1 += (unsigned int)((unsigned long long)(Xn.x & 0xFFFFFF) * (unsigned long long)(A.y & 0xFFFFFF));
    R1 &= 0xFFFFFF;
    return glm::uvec2(R0, R1);
}
/* rnd<>, krn:1516-1547: the index is the thread's list position */
template <int AGENT_TYPE> float rnd(RNG_rand48* rand48){
    int index = orc_tid;
    glm::uvec2 state = rand48->seeds[index];
    glm::uvec2 A = rand48->A;
    glm::uvec2 C = rand48->C;
    int rand = ( state.x >> 17 ) | ( state.y << 7);
    state = RNG_rand48_iterate_single(state, A, C);
    rand48->seeds[index] = state;
    return (float)rand/2147483647;
}
float rnd(RNG_rand48* rand48){ return rnd<DISCRETE_2D>(rand48); }
template float rnd<CONTINUOUS>(RNG_rand48*);
"""]
    for c in model.constants:
        dim = f"[{c.array_length}]" if c.array_length else ""
        out.append(f"{c.type} {c.name}{dim};")
    for m in model.messages:
        out.append(_message_api(m))
    for a in model.agents:
        scal = [v for v in a.variables if not v.array_length]
        out.append(f"""/* add_{a.name}_agent, krn:287-314 */
template <int AGENT_TYPE> void add_{a.name}_agent(xmachine_memory_{a.name}_list* agents, {codegen._agent_args(a)}){{
    int index = orc_tid;
    agents->_position[index] = 0;
    agents->_scan_input[index] = 1;""")
        for v in scal:
            out.append(f"    agents->{v.name}[index] = {v.name};")
        out.append("}")
        out.append(f"void add_{a.name}_agent(xmachine_memory_{a.name}_list* agents, {codegen._agent_args(a)}){{")
        out.append(f"    add_{a.name}_agent<DISCRETE_2D>(agents, {', '.join(v.name for v in scal)});\n}}")
        if any(v.array_length for v in a.variables):
            out.append(f"""/* get/set_{a.name}_agent_array_value, krn:336-367 */
template<typename T> T get_{a.name}_agent_array_value(T *array, unsigned int index){{
    if (array != nullptr) return array[index*xmachine_memory_{a.name}_MAX];
    return T(0);
}}
template<typename T> void set_{a.name}_agent_array_value(T *array, unsigned int index, T value){{
    if (array != nullptr) array[index*xmachine_memory_{a.name}_MAX] = value;
}}""")
    for p in function_files:
        out.append(f'#include "{os.path.abspath(p)}"')
    for f in model.all_functions():
        out.append(_thunk(model, f))
        if f.condition is not None or f.global_condition is not None:
            out.append(_filter(model, f))
    out.append(codegen._constants_code(model, cuda=False))
    out.append(codegen._descriptor_tables(model, "oracle", "orc_thunk_").replace("orc_thunk_filter_", "orc_filter_"))
    out.append(codegen._face2_shim(model, cuda=False))
    out.append(codegen._id_defs(model, cuda=False))
    for a in model.agents:      # sort_<A>s_<S> takes a CUDA kernel: declared for the scripts, not executable on the host
        for st in a.states:
            out.append(f"void sort_{a.name}s_{st}(void (*)(unsigned int*, unsigned int*, xmachine_memory_{a.name}_list*)){{ "
                       f'fprintf(stderr, "sort_{a.name}s_{st}: the oracle cannot launch a CUDA kernel\\n"); abort(); }}')
    out.append(codegen._model_struct(model, cuda=False))
    rows = []
    for m in model.messages:
        S = f"xmachine_message_{m.name}"
        if m.partitioning == "spatial":
            rows.append(f"    {{sizeof({S}_list), sizeof({S}_PBM), offsetof({S}_PBM, end_or_count), &d_message_{m.name}_count, &d_message_{m.name}_output_type}},")
        else:
            rows.append(f"    {{sizeof({S}_list), 0, 0, &d_message_{m.name}_count, &d_message_{m.name}_output_type}},")
    out.append('#include "oracle_rt.hpp"')
    # host API the scripts' init/step/exit functions may call (header.xslt:588-596, 729-751), on the oracle's lists
    out.append('static void fgpu_die(){ fprintf(stderr, "%s\\n", orc_last_error()); exit(EXIT_FAILURE); }')
    out.append("unsigned int getIterationNumber(){ return orc_cur ? orc_cur->iteration : 0u; }")
    out.append('static bool orc_exit_early = false;\nvoid set_exit_early(){ orc_exit_early = true; }\nbool get_exit_early(){ return orc_exit_early; }')
    out.append('const char* getOutputDir(){ const char* d = getenv("ORACLE_OUTPUT_DIR"); return d ? d : ""; }')
    for ai, a in enumerate(model.agents):
        out.append(f"int get_agent_{a.name}_MAX_count(){{ return xmachine_memory_{a.name}_MAX; }}")
        for si, st in enumerate(a.states):
            out.append(f"int get_agent_{a.name}_{st}_count(){{ return orc_sim_agent_count(orc_cur, {ai}, {si}); }}")
    out += codegen._reduction_defs(model, "orc_host_reduce")
    out.append("static int orc_add_cur(int a, int s, int n, const void* const* cols){ return orc_sim_add_agents(orc_cur, a, s, n, cols); }")
    out.append("static int orc_count_cur(int a, int s){ return orc_sim_agent_count(orc_cur, a, s); }")
    out += codegen._host_creation_defs(model, "orc_add_cur", "orc_count_cur")
    out.append("const orc_message_hooks orc_hooks[] = {\n" + "\n".join(rows) + "\n    {0, 0, 0, nullptr, nullptr}\n};")
    return "\n".join(out) + "\n"


def oracle_path(name: str) -> str:
    return os.path.join(LIBDIR, f"oracle_{name}.so")


def build_oracle(xml_path: str, name: str, function_files=None, force: bool = False) -> str:
    """g++ -O3 -march=native -ffp-contract=off -fopenmp."""
    model = xmml.parse_model(xml_path)
    mdir = os.path.dirname(os.path.abspath(xml_path))
    if function_files is None:
        function_files = [os.path.join(mdir, f) for f in model.function_files]
    gen = os.path.join(OUT, name)
    os.makedirs(gen, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    files = {"header.h": emit_header(model), "oracle_glue.cpp": emit_glue(model, function_files)}
    for fn, text in files.items():
        p = os.path.join(gen, fn)
        if not os.path.exists(p) or open(p).read() != text:
            with open(p, "w") as fh:
                fh.write(text)
    out = oracle_path(name)
    deps = [os.path.join(gen, "header.h"), os.path.join(gen, "oracle_glue.cpp"), os.path.join(HERE, "oracle_rt.hpp"),
            os.path.join(ROOT, "include", "fgpu_model.h")] + list(function_files)
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    cmd = ["g++", "-std=c++17", "-O3", "-march=x86-64-v3", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-w",
           "-I", gen, "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "ev_diffusion_b200", "csrc"),
           "-I", HERE, "-I", mdir, os.path.join(gen, "oracle_glue.cpp"), "-o", out]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("oracle build failed: %s\n%s" % (" ".join(cmd), proc.stdout[-6000:]))
    return out
