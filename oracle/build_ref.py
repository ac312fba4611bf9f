"""Build the reference's OWN generated simulation for a model as a host library (oracle/_ref/<name>/libref_<name>.so).

TEST INFRASTRUCTURE.  Pins the CPU oracle (oracle_rt.hpp + gen_oracle.py, a restatement) against the reference itself:

  1. expand the reference's templates -- ``FLAMEGPU/templates/{header,FLAMEGPU_kernals,simulation,io}.xslt``, read where
     they lie under /root/reference -- over the model's XMLModelFile.xml with oracle/xslt_subset.py (this image has no
     xsltproc), into ``oracle/_ref/<name>/dynamic/`` (git-ignored, and deleted again once the compile has succeeded:
     reference text never enters the repository);
  2. rewrite the ``kernel<<<grid, block, smem, stream>>>(args);`` launch statements of the expanded simulation.cu into
     ``cuda_on_host::Launch(grid, block, smem, stream).run(barriers, [&]{ kernel(args); });`` -- the only edit;
  3. compile simulation.cu (which includes FLAMEGPU_kernals.cu, which includes the model's own functions.c) and io.cu
     with g++ against oracle/cuda_on_host/ (host meaning of the CUDA runtime, cub and thrust entry points used) and
     the reference's vendored glm, plus a generated accessor shim (``ref_*`` C functions: initialise from a 0.xml,
     step, read any agent variable / the rand48 seeds / the sorted message list and PBM).

The result executes the reference's own host orchestration, kernels and agent scripts, thread by thread on the CPU
(-ffp-contract=off, as the oracle).  The generated header.h defines FAST_ATOMIC_SORTING (``header.xslt:35``): the default
message build is a counting sort with global atomics (``FLAMEGPU_kernals.xslt:899-934``); on one host thread its atomics
resolve in thread order, so it is deterministic here.  ``fast_atomic=False`` removes that one #define and builds the
deterministic variant (hash + radix sort + reorder,
``FLAMEGPU_kernals.xslt:944-1021``) -- the one the oracle restates.  Both must give the same states.
"""
from __future__ import annotations

import os
import re
import shutil
import subprocess
import sys
from typing import Dict, List, Optional, Set

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from ev_diffusion_b200 import xmml                                   # noqa: E402
from ev_diffusion_b200.looprewrite import mask_comments_and_strings  # noqa: E402
from oracle import xslt_subset                                       # noqa: E402

REFERENCE = os.environ.get("FGPU_REFERENCE", "/root/reference")
TEMPLATES = os.path.join(REFERENCE, "FLAMEGPU", "templates")
OUT = os.path.join(HERE, "_ref")
EXPANDED = {"header.xslt": "header.h", "FLAMEGPU_kernals.xslt": "FLAMEGPU_kernals.cu", "simulation.xslt": "simulation.cu",
            "io.xslt": "io.cu"}


def reference_available() -> bool:
    return os.path.isdir(TEMPLATES)


def ref_path(name: str) -> str:
    return os.path.join(OUT, name, f"libref_{name}.so")


# ---------------------------------------------------------------------------------------------------------------------
# launch rewriting and barrier analysis
# ---------------------------------------------------------------------------------------------------------------------

def _matching(text: str, pos: int, open_c: str, close_c: str) -> int:
    depth = 0
    for i in range(pos, len(text)):
        if text[i] == open_c:
            depth += 1
        elif text[i] == close_c:
            depth -= 1
            if depth == 0:
                return i
    raise ValueError("unbalanced %s%s" % (open_c, close_c))


def _active_branches_only(masked: str, defined: Set[str]) -> str:
    """Blank the inactive branches of preprocessor conditionals (same length, newlines kept): the expanded sources open
    a brace in both branches of an #ifdef and close it once, so only one branch may be seen when matching.  #ifdef and
    #ifndef and a lone `#if [!]defined(X)` are decided against `defined`; any other #if keeps its first branch."""
    out, live = [], []                      # live[k]: the current branch of conditional k is compiled
    for line in masked.split("\n"):
        d = re.match(r"\s*#\s*(ifdef|ifndef|if|else|elif|endif)\b\s*(\w*)", line)
        if d:
            kind = d.group(1)
            if kind == "ifdef":
                live.append(d.group(2) in defined)
            elif kind == "ifndef":
                live.append(d.group(2) not in defined)
            elif kind == "if":
                single = re.match(r"\s*#\s*if\s+(!?)\s*defined\s*\(?\s*(\w+)\s*\)?\s*$", line)
                live.append(((single.group(2) in defined) != (single.group(1) == "!")) if single else True)
            elif kind in ("else", "elif") and live:
                live[-1] = not live[-1]
            elif kind == "endif" and live:
                live.pop()
            out.append(" " * len(line))
        else:
            out.append(line if all(live) else " " * len(line))
    return "\n".join(out)


def function_bodies(src: str, defined: Set[str] = frozenset()) -> Dict[str, str]:
    """name -> concatenated bodies of every top-level function definition called `name` (comments/strings masked)."""
    masked = _active_branches_only(mask_comments_and_strings(src), set(defined))
    bodies: Dict[str, str] = {}
    depth, i, n = 0, 0, len(masked)
    while i < n:
        ch = masked[i]
        if ch == "{":
            if depth == 0:
                # look back for `name ( ... )` immediately before the brace
                j = i - 1
                while j >= 0 and masked[j].isspace():
                    j -= 1
                if j >= 0 and masked[j] == ")":
                    k, d = j, 0
                    while k >= 0:
                        if masked[k] == ")":
                            d += 1
                        elif masked[k] == "(":
                            d -= 1
                            if d == 0:
                                break
                        k -= 1
                    m = re.search(r"([A-Za-z_]\w*)\s*$", masked[:k])
                    if m and m.group(1) not in ("if", "for", "while", "switch"):
                        end = _matching(masked, i, "{", "}")
                        bodies[m.group(1)] = bodies.get(m.group(1), "") + masked[i:end + 1]
                        i = end + 1
                        continue
            depth += 1
        elif ch == "}":
            depth -= 1
        i += 1
    return bodies


def functions_reaching_barrier(sources: List[str], defined: Set[str] = frozenset()) -> Set[str]:
    bodies: Dict[str, str] = {}
    for s in sources:
        for k, v in function_bodies(s, defined).items():
            bodies[k] = bodies.get(k, "") + v
    calls = {k: set(re.findall(r"\b([A-Za-z_]\w*)\s*(?:<(?!<)[^;(){}]*>)?\s*\(", v)) & set(bodies) for k, v in bodies.items()}
    reach = {k for k, v in bodies.items() if re.search(r"\b__syncthreads\s*\(", v)}
    changed = True
    while changed:
        changed = False
        for k, c in calls.items():
            if k not in reach and c & reach:
                reach.add(k)
                changed = True
    return reach


def rewrite_launches(src: str, barrier_kernels: Set[str]) -> str:
    masked = mask_comments_and_strings(src)
    out, last = [], 0
    for m in re.finditer(r"\b([A-Za-z_]\w*)\s*<<<", masked):
        close = masked.index(">>>", m.end())
        paren = close + 3
        while masked[paren].isspace():
            paren += 1
        if masked[paren] != "(":
            raise ValueError("launch of %s is not followed by an argument list" % m.group(1))
        end = _matching(masked, paren, "(", ")")
        kernel = m.group(1)
        config, args = src[m.end():close], src[paren + 1:end]
        out.append(src[last:m.start()])
        out.append("cuda_on_host::Launch(%s).run(%s, [&]{ %s(%s); })" % (
            config.strip(), "true" if kernel in barrier_kernels else "false", kernel, args))
        last = end + 1
    out.append(src[last:])
    return "".join(out)


# ---------------------------------------------------------------------------------------------------------------------
# accessor shim (appended to the simulation translation unit, so it sees that file's globals)
# ---------------------------------------------------------------------------------------------------------------------

def emit_shim(model) -> str:
    L: List[str] = ["", "// ---- accessor shim generated by oracle/build_ref.py (not reference text) ----",
                    "// main.cu is not part of this build; it owns the output path (directory of the 0.xml input)",
                    "static char ref_output_dir[4096] = \"./\";",
                    "const char* getOutputDir() { return ref_output_dir; }",
                    'extern "C" {',
                    "void ref_initialise(const char* input) {",
                    "    std::snprintf(ref_output_dir, sizeof(ref_output_dir), \"%s\", input);",
                    "    char* slash = std::strrchr(ref_output_dir, '/');",
                    "    if (slash) slash[1] = 0; else std::snprintf(ref_output_dir, sizeof(ref_output_dir), \"./\");",
                    "    initialise(const_cast<char*>(input));",
                    "}",
                    "void ref_step(int n) { for (int i = 0; i < n; ++i) singleIteration(); }",
                    "// the reference's own iteration-file writer, called the way its main.cu calls it (main.xslt:367)",
                    "void ref_save_iteration(const char* outputpath, int iteration) {",
                    "    saveIterationData(const_cast<char*>(outputpath), iteration, " + ", ".join(
                        f"get_host_{a.name}_{s}_agents(), get_device_{a.name}_{s}_agents(), get_agent_{a.name}_{s}_count()"
                        for a in model.agents for s in a.states) + ");",
                    "}",
                    "void ref_cleanup() { cleanup(); }",
                    "unsigned ref_iteration() { return getIterationNumber(); }",
                    "int ref_buffer_size_max() { return buffer_size_MAX; }",
                    "// rand48 state per list slot as the reference holds it: (low 24 bits, high 24 bits)",
                    "void ref_seeds(unsigned int* dst) {",
                    "    for (int i = 0; i < buffer_size_MAX; ++i) { dst[2 * i] = d_rand48->seeds[i].x; dst[2 * i + 1] = d_rand48->seeds[i].y; }",
                    "}",
                    "int ref_count(int agent, int state) {", "    switch (agent * 64 + state) {"]
    for ai, a in enumerate(model.agents):
        for si, s in enumerate(a.states):
            L.append(f"    case {ai * 64 + si}: return h_xmachine_memory_{a.name}_{s}_count;")
    L += ["    }", "    return -1;", "}", "int ref_read_var(int agent, int state, int var, void* dst) {"]
    L.append("    switch ((agent * 64 + state) * 256 + var) {")
    for ai, a in enumerate(model.agents):
        for si, s in enumerate(a.states):
            for vi, v in enumerate(a.variables):
                if v.array_length:     # `length` blocks of `count` values (element e of agent i at var[e * MAX + i])
                    L.append(f"    case {(ai * 64 + si) * 256 + vi}: for (int e = 0; e < {v.array_length}; ++e) "
                             f"std::memcpy((char*)dst + sizeof({v.type}) * (size_t)e * h_xmachine_memory_{a.name}_{s}_count, "
                             f"d_{a.name}s_{s}->{v.name} + (size_t)e * xmachine_memory_{a.name}_MAX, "
                             f"sizeof({v.type}) * h_xmachine_memory_{a.name}_{s}_count); return 0;")
                    continue
                L.append(f"    case {(ai * 64 + si) * 256 + vi}: std::memcpy(dst, d_{a.name}s_{s}->{v.name}, "
                         f"sizeof({v.type}) * h_xmachine_memory_{a.name}_{s}_count); return 0;")
    L += ["    }", "    return -1;", "}", "int ref_message_count(int message) {", "    switch (message) {"]
    for mi, m in enumerate(model.messages):
        L.append(f"    case {mi}: return h_message_{m.name}_count;")
    L += ["    }", "    return -1;", "}", "int ref_read_message_var(int message, int var, void* dst) {",
          "    switch (message * 256 + var) {"]
    for mi, m in enumerate(model.messages):
        for vi, v in enumerate(m.variables):
            if v.array_length:
                continue
            L.append(f"    case {mi * 256 + vi}: std::memcpy(dst, d_{m.name}s->{v.name}, sizeof({v.type}) * h_message_{m.name}_count); return 0;")
    L += ["    }", "    return -1;", "}",
          "// PBM of a spatially partitioned message list: start[] and end_or_count[] as the reference stores them",
          "int ref_read_pbm(int message, int* start, int* end_or_count) {", "    switch (message) {"]
    for mi, m in enumerate(model.messages):
        if m.partitioning == "spatial":
            L.append(f"    case {mi}: std::memcpy(start, d_{m.name}_partition_matrix->start, sizeof(int) * xmachine_message_{m.name}_grid_size);")
            L.append(f"        std::memcpy(end_or_count, d_{m.name}_partition_matrix->end_or_count, sizeof(int) * xmachine_message_{m.name}_grid_size); return xmachine_message_{m.name}_grid_size;")
    L += ["    }", "    return -1;", "}", "}  // extern \"C\"", ""]
    return "\n".join(L)


# ---------------------------------------------------------------------------------------------------------------------
# build
# ---------------------------------------------------------------------------------------------------------------------

def expand(xml_path: str, outdir: str) -> Dict[str, str]:
    os.makedirs(outdir, exist_ok=True)
    written = {}
    for xslt, target in EXPANDED.items():
        text = xslt_subset.Stylesheet(os.path.join(TEMPLATES, xslt)).transform(xml_path)
        path = os.path.join(outdir, target)
        with open(path, "w") as f:
            f.write(text)
        written[target] = path
    return written


def ref_console_path(name: str) -> str:
    return os.path.join(OUT, name, f"ref_console_{name}")


def build_ref_console(xml_path: str, name: str, function_dir: Optional[str] = None, force: bool = False) -> str:
    """The reference's console PROGRAM on the host: its own main.cu (argument handling, messages, output paths,
    iteration loop, XML output frequency) linked with the host build of simulation.cu / io.cu.  Used to check the
    behaviour of this repository's console program (csrc/fgpu_console.cpp) where no device is involved."""
    if not reference_available():
        raise RuntimeError("the reference tree is not mounted (%s)" % REFERENCE)
    model = xmml.parse_model(xml_path)
    mdir = function_dir or os.path.dirname(os.path.abspath(xml_path))
    out = ref_console_path(name)
    if not force and os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(os.path.abspath(__file__)), os.path.getmtime(xml_path)):
        return out
    dyn = os.path.join(OUT, name, "dynamic_console")
    files = expand(xml_path, dyn)
    main = xslt_subset.Stylesheet(os.path.join(TEMPLATES, "main.xslt")).transform(xml_path)
    kernels = open(files["FLAMEGPU_kernals.cu"]).read()
    scripts = [open(os.path.join(mdir, f)).read() for f in model.function_files]
    reach = functions_reaching_barrier([kernels] + scripts, {"FAST_ATOMIC_SORTING"})
    with open(os.path.join(dyn, "simulation_host.cpp"), "w") as f:
        f.write("CUDA_ON_HOST_SHARED_MEMORY_DEFINITION\n" + rewrite_launches(open(files["simulation.cu"]).read(), reach))
    with open(os.path.join(dyn, "io_host.cpp"), "w") as f:
        f.write(rewrite_launches(open(files["io.cu"]).read(), reach))
    with open(os.path.join(dyn, "main_host.cpp"), "w") as f:
        f.write(rewrite_launches(main, reach))
    cmd = ["g++", "-std=c++17", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-w", "-fpermissive",
           "-x", "c++", "-include", os.path.join(HERE, "cuda_on_host", "cuda_on_host.h"),
           "-I", os.path.join(HERE, "cuda_on_host"), "-I", dyn, "-I", mdir, "-I", os.path.join(REFERENCE, "include"),
           "-DGLM_FORCE_PURE", "-Wl,--allow-multiple-definition",
           os.path.join(dyn, "main_host.cpp"), os.path.join(dyn, "simulation_host.cpp"), os.path.join(dyn, "io_host.cpp"), "-o", out]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(out + ".log", "w") as f:
        f.write(" ".join(cmd) + "\n" + proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("reference console build failed (%s):\n%s" % (name, proc.stdout[-8000:]))
    _discard_expansion(dyn)
    return out


def _discard_expansion(directory: str) -> None:
    """The expanded sources are the reference's text: they exist only for the duration of the compile (kept after a
    failed one, and with FGPU_KEEP_REF_SOURCES=1, for reading)."""
    if os.environ.get("FGPU_KEEP_REF_SOURCES") != "1":
        shutil.rmtree(directory, ignore_errors=True)


def _select_sort_variant(header_path: str, fast_atomic: bool) -> None:
    if fast_atomic:
        return
    text = open(header_path).read()
    patched = re.sub(r"^[ \t]*#define[ \t]+FAST_ATOMIC_SORTING[ \t]*$", "/* #define FAST_ATOMIC_SORTING  (removed by oracle/build_ref.py) */",
                     text, count=1, flags=re.M)
    if patched == text:
        raise RuntimeError("header.h does not define FAST_ATOMIC_SORTING where expected")
    with open(header_path, "w") as f:
        f.write(patched)


def build_ref(xml_path: str, name: str, function_dir: Optional[str] = None, fast_atomic: bool = True, force: bool = False,
              opt: str = "-O2", device_patch: bool = False) -> str:
    """Expand, rewrite and compile; returns the library path.  Needs /root/reference (templates + glm).
    `device_patch` applies the texture-reference rewrite of the sm_100a build (build_ref_device) first, so that the
    very sources nvcc compiles for the GPU can be checked on the CPU (library name gets a `_patched` suffix)."""
    if not reference_available():
        raise RuntimeError("the reference tree is not mounted (%s): oracle/_ref can only be built where it is" % REFERENCE)
    model = xmml.parse_model(xml_path)
    mdir = function_dir or os.path.dirname(os.path.abspath(xml_path))
    out = ref_path(name + ("_patched" if device_patch else "") + ("" if fast_atomic else "_radix"))
    os.makedirs(os.path.dirname(out), exist_ok=True)
    deps = [xml_path, os.path.abspath(__file__), os.path.join(HERE, "xslt_subset.py"), os.path.join(HERE, "cuda_on_host", "cuda_on_host.h")]
    deps += [os.path.join(mdir, f) for f in model.function_files] + [os.path.join(TEMPLATES, x) for x in EXPANDED]
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps if os.path.exists(d)):
        return out
    dyn = os.path.join(os.path.dirname(out), "dynamic")
    files = expand(xml_path, dyn)
    _select_sort_variant(files["header.h"], fast_atomic)
    prelude = []
    if device_patch:
        for target in ("FLAMEGPU_kernals.cu", "simulation.cu", "io.cu"):
            text = patch_texture_references(open(files[target]).read())
            with open(files[target], "w") as f:
                f.write(text)
        with open(os.path.join(dyn, "texture_prelude.h"), "w") as f:
            f.write(DEVICE_PRELUDE)
        prelude = ["-include", os.path.join(dyn, "texture_prelude.h")]
    kernels = open(files["FLAMEGPU_kernals.cu"]).read()
    scripts = [open(os.path.join(mdir, f)).read() for f in model.function_files]
    reach = functions_reaching_barrier([kernels] + scripts, {"FAST_ATOMIC_SORTING"} if fast_atomic else set())
    sim = rewrite_launches(open(files["simulation.cu"]).read(), reach)
    with open(os.path.join(dyn, "simulation_host.cpp"), "w") as f:
        f.write("CUDA_ON_HOST_SHARED_MEMORY_DEFINITION\n" + sim + emit_shim(model))
    with open(os.path.join(dyn, "io_host.cpp"), "w") as f:
        f.write(rewrite_launches(open(files["io.cu"]).read(), reach))
    with open(os.path.join(dyn, "barriers.txt"), "w") as f:
        f.write("\n".join(sorted(reach)) + "\n")
    cmd = ["g++", "-std=c++17", opt, "-march=x86-64-v3", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-fpermissive",
           "-x", "c++", "-include", os.path.join(HERE, "cuda_on_host", "cuda_on_host.h")] + prelude + [
           "-I", os.path.join(HERE, "cuda_on_host"), "-I", dyn, "-I", mdir, "-I", os.path.join(REFERENCE, "include"),
           "-DGLM_FORCE_PURE",
           # header.h defines `__device__` variables; nvcc gives every translation unit its own copy, only the simulation
           # unit's is ever used -- on the host the two definitions may simply be merged
           "-Wl,--allow-multiple-definition"]
    cmd += [os.path.join(dyn, "simulation_host.cpp"), os.path.join(dyn, "io_host.cpp"), "-o", out]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(out + ".log", "w") as f:
        f.write(" ".join(cmd) + "\n" + proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("reference host build failed (%s):\n%s" % (name, proc.stdout[-8000:]))
    _discard_expansion(dyn)
    return out


# ---------------------------------------------------------------------------------------------------------------------
# device build: the reference's generated CUDA compiled for sm_100a (timing baseline on the GPU box)
# ---------------------------------------------------------------------------------------------------------------------

DEVICE_PRELUDE = r"""
// ---- prelude written by oracle/build_ref.py (not reference text) ----------------------------------------------------
// CUDA 12 removed texture REFERENCES (texture<T,1,cudaReadModeElementType> + cudaBindTexture + tex1Dfetch(ref, i)).  The
// generated sources use them for one thing only: cached reads of a linear device array.  build_ref.py rewrites each
// `texture<T, 1, cudaReadModeElementType> NAME;` into `__device__ const T* NAME;`; the three calls keep their spelling:
#include <cuda_runtime.h>
// thrust entry points the sources call but, with today's thrust, no longer receive through their own includes
#include <thrust/count.h>
#include <thrust/reduce.h>
template <class T> static __device__ __forceinline__ T tex1Dfetch(const T* p, int i) { return __ldg(p + i); }
template <class T> static cudaError_t fgpu_ref_bind(size_t* offset, T*& symbol, const void* p) {
    if (offset) *offset = 0;
    return cudaMemcpyToSymbol(symbol, &p, sizeof(p));
}
#define cudaBindTexture(offset, tex, ptr, size) fgpu_ref_bind(offset, tex, (const void*)(ptr))
#define cudaUnbindTexture(tex) cudaSuccess
// ---------------------------------------------------------------------------------------------------------------------
"""


def patch_texture_references(src: str) -> str:
    tex = r"\btexture\s*<\s*([\w:\s]+?)\s*,\s*1\s*,\s*cudaReadModeElementType\s*>"
    src = re.sub(tex + r"\s*(\w+)\s*;", r"__device__ const \1* \2;", src)        # the references themselves (file scope)
    return re.sub(tex, r"const \1*", src)                                        # helpers that take one by value


def ref_device_path(name: str) -> str:
    return os.path.join(OUT, name, f"ref_gpu_{name}")


def build_ref_device(xml_path: str, name: str, function_dir: Optional[str] = None, fast_atomic: bool = True,
                     force: bool = False) -> str:
    """The reference's console program for a model (its own main.cu, simulation.cu, kernels, io.cu and the model's
    functions.c) compiled by nvcc for sm_100a: ``ref_gpu_<name> <0.xml> <iterations>`` prints the reference's own
    "Total Processing time".  Edits: texture references (see DEVICE_PRELUDE) -- nothing else.  Default build options are
    the reference's (FAST_ATOMIC_SORTING on, -use_fast_math off, Release -O3)."""
    if not reference_available():
        raise RuntimeError("the reference tree is not mounted (%s)" % REFERENCE)
    model = xmml.parse_model(xml_path)
    mdir = function_dir or os.path.dirname(os.path.abspath(xml_path))
    out = ref_device_path(name)
    dyn = os.path.join(OUT, name, "dynamic_device")
    os.makedirs(dyn, exist_ok=True)
    sources = dict(EXPANDED)
    sources["main.xslt"] = "main.cu"
    deps = [xml_path, os.path.abspath(__file__), os.path.join(HERE, "xslt_subset.py")]
    deps += [os.path.join(mdir, f) for f in model.function_files] + [os.path.join(TEMPLATES, x) for x in sources]
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps if os.path.exists(d)):
        return out
    for xslt, target in sources.items():
        text = xslt_subset.Stylesheet(os.path.join(TEMPLATES, xslt)).transform(xml_path)
        if target.endswith(".cu"):
            text = patch_texture_references(text)
        with open(os.path.join(dyn, target), "w") as f:
            f.write(text)
    _select_sort_variant(os.path.join(dyn, "header.h"), fast_atomic)
    with open(os.path.join(dyn, "texture_prelude.h"), "w") as f:
        f.write(DEVICE_PRELUDE)
    # the reference vendors glm AND a 2016 cub under include/; only glm may be seen (cub and thrust are the toolkit's)
    inc = os.path.join(dyn, "include")
    os.makedirs(inc, exist_ok=True)
    if not os.path.islink(os.path.join(inc, "glm")):
        os.symlink(os.path.join(REFERENCE, "include", "glm"), os.path.join(inc, "glm"))
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-w",
           "-include", os.path.join(dyn, "texture_prelude.h"), "-I", dyn, "-I", mdir, "-I", inc, "-DGLM_FORCE_PURE"]
    cmd += [os.path.join(dyn, "main.cu"), os.path.join(dyn, "simulation.cu"), os.path.join(dyn, "io.cu"), "-o", out]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(out + ".log", "w") as f:
        f.write(" ".join(cmd) + "\n" + proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("reference device build failed (%s):\n%s" % (name, proc.stdout[-8000:]))
    _discard_expansion(dyn)
    return out


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("xml")
    ap.add_argument("name")
    ap.add_argument("--radix", action="store_true", help="the deterministic message build (FAST_ATOMIC_SORTING removed)")
    ap.add_argument("--device", action="store_true", help="nvcc build of the reference's console program for sm_100a")
    a = ap.parse_args()
    if a.device:
        print(build_ref_device(a.xml, a.name, fast_atomic=not a.radix, force=True))
    else:
        print(build_ref(a.xml, a.name, fast_atomic=not a.radix, force=True))
