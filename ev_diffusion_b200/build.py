"""Build helpers: nvcc for sm_100a, in-tree outputs.

Replaces the make -> xsltproc -> nvcc chain of ``tools/common.mk:252-451``:

  build_runtime()            csrc/fgpu_rt.cu              -> _lib/libfgpu_rt.so
  build_model(xml, name)     XMLModelFile.xml+functions.c -> _lib/models/<name>.so

All outputs are git-ignored ``.so`` files inside the package, so they travel
with the repository snapshot to the GPU box.  Nothing is JIT-compiled at run
time; a missing library is an error, never a fallback.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from typing import List, Optional

from . import codegen, xmml

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB = os.path.join(PKG, "_lib")
GEN = os.path.join(PKG, "_gen")
MODELS = os.path.join(PKG, "models")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared",
          "-Xlinker", "-rpath=/usr/local/cuda/lib64", "-diag-suppress", "20012,177,550"]


class BuildError(RuntimeError):
    pass


def _run(cmd: List[str], log: Optional[str] = None) -> str:
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log:
        with open(log, "w") as fh:
            fh.write(" ".join(cmd) + "\n" + proc.stdout)
    if proc.returncode != 0:
        raise BuildError("command failed: %s\n%s" % (" ".join(cmd), proc.stdout[-6000:]))
    return proc.stdout


def _stamp(paths: List[str], extra: str = "") -> str:
    h = hashlib.sha256(extra.encode())
    for p in sorted(paths):
        with open(p, "rb") as fh:
            h.update(p.encode())
            h.update(fh.read())
    return h.hexdigest()


def _fresh(out: str, stamp: str) -> bool:
    sp = out + ".stamp"
    return os.path.exists(out) and os.path.exists(sp) and open(sp).read() == stamp


def _mark(out: str, stamp: str) -> None:
    with open(out + ".stamp", "w") as fh:
        fh.write(stamp)


def runtime_path() -> str:
    return os.path.join(LIB, "libfgpu_rt.so")


def model_path(name: str) -> str:
    return os.path.join(LIB, "models", name + ".so")


def build_runtime(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB, exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.startswith("fgpu_rt") or f == "fgpu_kernels.cuh"]
    hdrs = [os.path.join(INCLUDE, f) for f in ("fgpu.h", "fgpu_model.h")]
    out = runtime_path()
    extra = []
    cmd = [NVCC] + ARCH_FLAGS + COMMON + ["-Xptxas", "-v", "-I", INCLUDE, "-I", CSRC] + extra + \
          [os.path.join(CSRC, "fgpu_rt.cu"), "-o", out, "-ldl"]
    stamp = _stamp(srcs + hdrs, " ".join(cmd))
    if not force and _fresh(out, stamp):
        return out
    text = _run(cmd, log=out + ".log")
    if verbose:
        print(text)
    _mark(out, stamp)
    return out


IF_CONVERT_DEFAULT = int(os.environ.get("FGPU_IF_CONVERT", "16"))
_if_convert_ok: Optional[bool] = None


def if_convert_flags(threshold: int) -> List[str]:
    """nvcc flags that raise NVVM's if-conversion budget (SimplifyCFG's two-entry-phi folding threshold, default 4
    instructions) to `threshold`.  A message-loop body of the usual shape -- `if (msg->id != agent->id) { a dozen float
    operations; if (d2 < r2) count++; }` -- is otherwise compiled as a branch per message: BSSY/BSYNC + BRA and a
    three-instruction conditional increment at the reconvergence point, 35 SASS instructions per pair of messages on
    C2 against 27 when the body is predicated (lanes of a warp almost never skip the body together, so the branch
    saves nothing).  Arithmetic and results are unchanged (same operations, same order).  The option goes to cicc's
    optimiser (`--Xopt`); a toolchain that rejects it builds without (probe below), so it is tuning, never a dependency."""
    global _if_convert_ok
    if threshold <= 0:
        return []
    flags = ["-Xcicc", "--Xopt", "-Xcicc", "-two-entry-phi-node-folding-threshold=%d" % threshold]
    if _if_convert_ok is None:
        import tempfile
        with tempfile.TemporaryDirectory() as td:
            src = os.path.join(td, "probe.cu")
            with open(src, "w") as fh:
                fh.write("__global__ void k(int* p){ if (p[0]) p[1] += 1; }\n")
            proc = subprocess.run([NVCC] + ARCH_FLAGS + flags + ["-ptx", src, "-o", os.path.join(td, "probe.ptx")],
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            _if_convert_ok = proc.returncode == 0
    return flags if _if_convert_ok else []


def build_model(xml_path: str, name: str, fmad: bool = True, force: bool = False, verbose: bool = False,
                function_files: Optional[List[str]] = None, extra_flags: Optional[List[str]] = None,
                rewrite_loops: bool = True, loop_unroll: Optional[int] = 2, if_convert: Optional[int] = None) -> str:
    """Generate and compile the plugin of one model.

    ``fmad=False`` builds the exact-arithmetic variant (no FMA contraction) whose
    float results are bit-comparable with the host oracle compiled with
    ``-ffp-contract=off``.  ``rewrite_loops=False`` compiles the
    scripts byte for byte (generic get_next) instead of the loop-restructured
    copies (looprewrite.py); results are identical either way.  ``loop_unroll``
    is the unroll factor requested for the rewritten inner loops: 2 measured
    best on C2 (count_neighbours 162 us; 1: 185, 3: 175, compiler's choice 4: 170,
    8: 181 -- a row holds ~12 messages and the lanes of a warp sit in ~8 cells
    with different row lengths, so short unrolled bodies waste fewer lanes).
    ``if_convert`` is the if-conversion budget of `if_convert_flags` (None: IF_CONVERT_DEFAULT, 0: the compiler's own).
    """
    build_runtime()   # the plugin's Face-2 shim links against libfgpu_rt.so
    model = xmml.parse_model(xml_path)
    mdir = os.path.dirname(os.path.abspath(xml_path))
    if function_files is None:
        function_files = [os.path.join(mdir, f) for f in model.function_files]
    for f in function_files:
        if not os.path.exists(f):
            raise BuildError(f"function file {f} not found")
    gen = os.path.join(GEN, name)
    files = codegen.emit_cuda(model, function_files, gen, rewrite_loops=rewrite_loops, loop_unroll=loop_unroll)
    os.makedirs(os.path.join(LIB, "models"), exist_ok=True)
    out = model_path(name)
    cmd = [NVCC] + ARCH_FLAGS + COMMON + list(extra_flags or []) + \
          if_convert_flags(IF_CONVERT_DEFAULT if if_convert is None else if_convert) + \
          ["-Xptxas", "-v", "-fmad=" + ("true" if fmad else "false"),
                                          "-I", gen, "-I", INCLUDE, "-I", CSRC, "-I", mdir,
                                          files["model_glue.cu"], "-o", out,
                                          "-L", LIB, "-lfgpu_rt", "-Xlinker", "-rpath=$ORIGIN/.."]
    deps = sorted(files.values()) + [os.path.join(CSRC, "fgpu_device.cuh"),
            os.path.join(CSRC, "fgpu_glm.h"), os.path.join(INCLUDE, "fgpu_model.h"), os.path.join(INCLUDE, "fgpu.h")] + list(function_files)
    stamp = _stamp(deps, " ".join(cmd))
    # scripts from outside this repository (the reference's examples) are compiled where they lie; their
    # loop-rewritten copies are compiler input only and do not stay in the tree
    transient = [p for k, p in files.items() if k.startswith("rewritten_")] \
        if any(not os.path.abspath(f).startswith(ROOT + os.sep) for f in function_files) else []
    try:
        if not force and _fresh(out, stamp):
            return out
        text = _run(cmd, log=out + ".log")
        if verbose:
            print(text)
        _mark(out, stamp)
        return out
    finally:
        for p in transient:
            if os.path.exists(p):
                os.remove(p)


def build_driver(source: str, model_name: str, out_name: str, force: bool = False) -> str:
    """Compile a host driver written against the reference's Face-2 names (initialise,
    singleIteration, ...) and link it to a model plugin + libfgpu_rt.so."""
    gen = os.path.join(GEN, model_name)
    out = os.path.join(LIB, "drivers", out_name)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = [NVCC] + ARCH_FLAGS + ["-std=c++17", "-O2", "-cudart", "shared", "-diag-suppress", "20012,177,550",
                                 "-I", gen, "-I", INCLUDE, "-I", CSRC, source, "-o", out,
                                 "-L", os.path.join(LIB, "models"), f"-l:{model_name}.so", "-L", LIB, "-lfgpu_rt",
                                 "-Xlinker", "-rpath=$ORIGIN/../models:$ORIGIN/..:/usr/local/cuda/lib64"]
    deps = [source, model_path(model_name), runtime_path()]
    stamp = _stamp(deps, " ".join(cmd))
    if not force and _fresh(out, stamp):
        return out
    _run(cmd, log=out + ".log")
    _mark(out, stamp)
    return out


def build_console(force: bool = False) -> str:
    """The reference's console mode for any model plugin (csrc/fgpu_console.cpp): host code only, linked to
    libfgpu_rt.so."""
    build_runtime()
    src = os.path.join(CSRC, "fgpu_console.cpp")
    out = os.path.join(LIB, "fgpu_console")
    cmd = ["g++", "-std=c++17", "-O2", "-I", INCLUDE, src, "-o", out, "-L", LIB, "-lfgpu_rt", "-Wl,-rpath,$ORIGIN", "-ldl"]
    stamp = _stamp([src, os.path.join(INCLUDE, "fgpu.h"), runtime_path()], " ".join(cmd))
    if not force and _fresh(out, stamp):
        return out
    _run(cmd, log=out + ".log")
    _mark(out, stamp)
    return out


def write_scaled_xml(src_xml: str, dst_xml: str, replacements: dict) -> str:
    """XML-only scaling of a model (bufferSize, bounds, radius, constants): plain
    text substitution of ``<tag>old</tag>`` pairs, the functions.c is untouched."""
    text = open(src_xml).read()
    for old, new in replacements.items():
        if old not in text:
            raise BuildError(f"scaling: '{old}' not found in {src_xml}")
        text = text.replace(old, new)
    os.makedirs(os.path.dirname(dst_xml), exist_ok=True)
    if not os.path.exists(dst_xml) or open(dst_xml).read() != text:
        with open(dst_xml, "w") as fh:
            fh.write(text)
    return dst_xml


if __name__ == "__main__":
    build_runtime(force="--force" in sys.argv, verbose=True)
