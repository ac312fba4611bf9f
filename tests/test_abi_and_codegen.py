"""CPU-side checks: the C-ABI library exports every symbol include/fgpu.h declares
(no compute calls without a GPU), the model plugins export fgpu_model_get with a
sane descriptor, and the host logic of the generator (XMML loader, write-set
analysis, condition emission, partition constants)."""
import ctypes
import os
import re
import subprocess

import pytest

from common import ROOT, W, build
from ev_diffusion_b200 import codegen, xmml
from ev_diffusion_b200.runtime import ModelInfo, _Model


def declared_symbols(header):
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fgpu_[a-z0-9_]+)\s*\(", text)))


def exported_symbols(so):
    out = subprocess.run(["nm", "-D", "--defined-only", so], stdout=subprocess.PIPE, text=True, check=True).stdout
    return {line.split()[-1] for line in out.splitlines() if line.strip()}


def test_runtime_exports_every_declared_symbol():
    so = build.runtime_path()
    assert os.path.exists(so), "run __graft_entry__.build() first"
    want = [s for s in declared_symbols(os.path.join(ROOT, "include", "fgpu.h"))]
    have = exported_symbols(so)
    missing = [s for s in want if s not in have]
    assert not missing, missing
    assert len(want) >= 30
    # the library loads without a GPU and reports the absence loudly instead of falling back
    lib = ctypes.CDLL(so)
    lib.fgpu_last_error.restype = ctypes.c_char_p
    lib.fgpu_sim_create.restype = ctypes.c_void_p
    assert lib.fgpu_sim_create(None, 0) is None
    assert b"null model" in lib.fgpu_last_error()


def test_plugin_exports_model_descriptor():
    so = build.model_path("Brownian")
    assert os.path.exists(so)
    assert "fgpu_model_get" in exported_symbols(so)
    lib = ctypes.CDLL(so)
    lib.fgpu_model_get.restype = ctypes.POINTER(_Model)
    m = ModelInfo(lib.fgpu_model_get().contents)
    assert m.name == "Brownian" and m.buffer_size_max == 1 << 20
    assert [a.name for a in m.agents] == ["Particle"]
    assert [v.name for v in m.agents[0].vars] == ["id", "x", "y", "z", "nbr"]
    msg = m.messages[0]
    assert msg.dims == (64, 64, 64) and msg.grid_size == 262144 and msg.radix_bits == 18 and not msg.is_2d
    assert [f.name for f in m.functions] == ["output_location", "count_neighbours", "move"]
    assert all(f.swappable for f in m.functions) and m.functions[2].rng
    assert m.layers == [[0], [1], [2]]


def test_no_cpu_fallback_in_product_path():
    """The product package never imports the oracle (a routed-through oracle would void parity)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ev_diffusion_b200")):
        if "_gen" in dirpath or "_lib" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "oracle_rt" not in text, f


def test_xmml_loader_on_brownian():
    m = xmml.parse_model(W.BROWNIAN_XML)
    assert m.buffer_size_max == 1 << 20
    msg = m.message("location")
    assert msg.partition_dims() == (64, 64, 64) and msg.grid_size() == 262144 and not msg.is_2d()
    f = m.function("move")
    assert f.rng and f.reallocate_literal == "false" and not f.reallocate
    assert [c.name for c in m.constants] == ["SIGMA", "DOMAIN_X", "DOMAIN_Y", "DOMAIN_Z"]
    assert m.layers == [["output_location"], ["count_neighbours"], ["move"]]


def test_write_set_analysis():
    m = xmml.parse_model(W.BROWNIAN_XML)
    src = [open(W.brownian_function_files()[0]).read()]
    a = m.agent("Particle")
    assert codegen.written_variables(m.function("output_location"), a, src) == set()
    assert codegen.written_variables(m.function("count_neighbours"), a, src) == {"nbr"}
    assert codegen.written_variables(m.function("move"), a, src) == {"x", "y", "z"}
    # a pointer that escapes, or a macro hiding a member access, makes everything written
    esc = ["__FLAME_GPU_FUNC__ int move(xmachine_memory_Particle* agent, RNG_rand48* r){ helper(agent); return 0; }"]
    assert codegen.written_variables(m.function("move"), a, esc) == {"id", "x", "y", "z", "nbr"}
    mac = ["#define PX agent->x\n__FLAME_GPU_FUNC__ int move(xmachine_memory_Particle* agent, RNG_rand48* r){ PX = 1; return 0; }"]
    assert codegen.written_variables(m.function("move"), a, mac) == {"id", "x", "y", "z", "nbr"}
    # ... and so does a macro standing for the member NAME (`agent->FIELD = ...` is a write of x the text does not show)
    fld = ["#define FIELD x\n__FLAME_GPU_FUNC__ int move(xmachine_memory_Particle* agent, RNG_rand48* r){ agent->FIELD = 1; return 0; }"]
    assert codegen.written_variables(m.function("move"), a, fld) == {"id", "x", "y", "z", "nbr"}
    inc = ["__FLAME_GPU_FUNC__ int move(xmachine_memory_Particle* a, RNG_rand48* r){ ++a->nbr; a->x *= 2; if (a->y == 3) return 1; return 0; }"]
    assert codegen.written_variables(m.function("move"), a, inc) == {"nbr", "x"}


def test_condition_expression_is_the_reference_text():
    """cmn:270-301 pastes <value> verbatim with full parenthesisation."""
    c = xmml.Condition("var:movement", "<", "0.25")
    assert c.c_expr("currentState->{0}[index]") == "(currentState->movement[index]<0.25)"
    n = xmml.Condition(xmml.Condition("var:a", "==", "0"), "&&", xmml.Condition("1", "!=", "var:b"))
    assert n.c_expr("s->{0}[i]") == "((s->a[i]==0)&&(1!=s->b[i]))"


def test_loader_rejects_what_the_generator_rejects(tmp_path):
    base = open(W.BROWNIAN_XML).read()
    bad_radius = tmp_path / "bad_radius.xml"
    bad_radius.write_text(base.replace("<gpu:radius>1.0</gpu:radius>", "<gpu:radius>0.7</gpu:radius>"))
    with pytest.raises(xmml.ModelError, match="factor"):      # simulation.xslt:46-76
        xmml.parse_model(str(bad_radius))
    too_few = tmp_path / "too_few.xml"
    too_few.write_text(base.replace("<gpu:radius>1.0</gpu:radius>", "<gpu:radius>32</gpu:radius>"))
    with pytest.raises(xmml.ModelError, match="at least 3"):  # simulation.xslt:79-84
        xmml.parse_model(str(too_few))
    dangling = tmp_path / "dangling.xml"
    dangling.write_text(base.replace("<name>move</name>\n      </gpu:layerFunction>", "<name>nope</name>\n      </gpu:layerFunction>"))
    with pytest.raises(xmml.ModelError, match="no agent defines"):
        xmml.parse_model(str(dangling))


def test_scaled_variants_share_the_script():
    cfg = W.BROWNIAN["c2_16m"]
    m = xmml.parse_model(W.brownian_xml(cfg))
    assert m.buffer_size_max == 1 << 24
    assert m.message("location").partition_dims() == (128, 128, 256)
    assert m.message("location").radix_bits() == 22


def test_splitmix_generator_is_stable():
    import numpy as np
    z = W.splitmix64(42, 3)
    assert [int(v) for v in z] == [13679457532755275413, 2949826092126892291, 5139283748462763858]
    u = W.uniform24(42, 4)
    assert u.dtype == np.float32 and ((u >= 0) & (u < 1)).all()
    st = W.brownian_state(W.BROWNIAN["c2_tiny"], 10)
    assert st["x"].max() < 16 and st["id"].tolist() == list(range(10))


def test_product_path_fails_loudly_without_its_cuda_parts(tmp_path, monkeypatch):
    """No CPU fallback: a missing runtime library, a missing plugin and a box without a GPU are errors with a
    message, never a silent detour through the oracle."""
    from ev_diffusion_b200 import runtime
    from ev_diffusion_b200.runtime import FgpuError
    # 1. missing model plugin
    with pytest.raises(FgpuError, match="missing|dlopen"):
        runtime.Simulation(str(tmp_path / "no_such_model.so"))
    # 2. no CUDA device in this container: creating a simulation must fail with the CUDA error, not fall back
    import torch
    if not torch.cuda.is_available():
        plugin = build.model_path("Brownian_tiny")
        if os.path.exists(plugin):
            with pytest.raises(FgpuError) as err:
                runtime.Simulation(plugin)
            assert "cuda" in str(err.value).lower() or "device" in str(err.value).lower()
    # 3. missing runtime library
    monkeypatch.setattr(runtime, "_RT", None)
    monkeypatch.setattr(build, "runtime_path", lambda: str(tmp_path / "libfgpu_rt.so"))
    with pytest.raises(FgpuError, match="no CPU fallback"):
        runtime.load_runtime()


def test_every_prebuilt_plugin_matches_the_runtime_abi():
    """The in-tree plugins travel to the GPU box as built: a plugin left over from an older plugin ABI would only
    fail there.  dlopen + descriptor check needs no GPU."""
    import glob
    from ev_diffusion_b200 import runtime
    lib = runtime.load_runtime()
    plugins = sorted(glob.glob(os.path.join(build.LIB, "models", "*.so")))
    assert plugins, "no model plugin built: run __graft_entry__.build()"
    stale = []
    for p in plugins:
        if not lib.fgpu_model_load(os.fsencode(p)):
            stale.append((os.path.basename(p), lib.fgpu_last_error().decode()))
    assert not stale, stale
