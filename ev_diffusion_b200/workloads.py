"""Synthetic workloads of BASELINE.md / SURVEY.md 8(d): model variants (XML-only
scaling of the authored models) and their seeded initial states.

Initial positions come from splitmix64 (top 24 bits / 2^24), as specified for
config C2: ``positions = extent * u``, seed 42 for C2, 46 for C2-16M.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, Tuple

import numpy as np

from . import build

MASK64 = (1 << 64) - 1


def splitmix64(seed: int, n: int) -> np.ndarray:
    """n outputs of splitmix64 started at `seed` (vectorised: the state is seed + (i+1)*gamma)."""
    gamma = np.uint64(0x9E3779B97F4A7C15)
    with np.errstate(over="ignore"):
        z = np.uint64(seed & MASK64) + gamma * (np.arange(n, dtype=np.uint64) + np.uint64(1))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def uniform24(seed: int, n: int) -> np.ndarray:
    """float32 uniforms in [0,1) with 24-bit resolution (exact in binary32)."""
    return ((splitmix64(seed, n) >> np.uint64(40)).astype(np.float32) / np.float32(16777216.0)).astype(np.float32)


@dataclass
class BrownianConfig:
    name: str
    n: int
    dims: Tuple[int, int, int]
    seed: int
    sigma: float = 0.25
    buffer: int = 0      # bufferSize of the model (per rank); 0 = n
    world: int = 1       # number of slabs the global population is cut into

    @property
    def buffer_size(self) -> int:
        return self.buffer or self.n

    @property
    def cells(self) -> int:
        return self.dims[0] * self.dims[1] * self.dims[2]


BROWNIAN = {
    # BASELINE config C2: 1M agents, 64^3 cells, 4 agents / cell
    "c2": BrownianConfig("Brownian", 1 << 20, (64, 64, 64), 42),
    # north-star target: 16M agents, 128 x 128 x 256 cells
    "c2_16m": BrownianConfig("Brownian_16M", 1 << 24, (128, 128, 256), 46),
    # strong scaling of the north-star population over z-slabs: 16M agents in total, 256 cell planes cut into
    # 128 / 64 / 32 planes per rank; per-rank buffers = the rank's share + 64K of headroom
    "c2_16m_w2": BrownianConfig("Brownian_16M_w2", 1 << 24, (128, 128, 256), 46, buffer=(1 << 23) + 65536, world=2),
    "c2_16m_w4": BrownianConfig("Brownian_16M_w4", 1 << 24, (128, 128, 256), 46, buffer=(1 << 22) + 65536, world=4),
    "c2_16m_w8": BrownianConfig("Brownian_16M_w8", 1 << 24, (128, 128, 256), 46, buffer=(1 << 21) + 65536, world=8),
    # small parity case (the oracle finishes in well under a second)
    "c2_small": BrownianConfig("Brownian_small", 1 << 14, (16, 16, 16), 7),
    # tiny ragged case: count < bufferSize, few agents per tile, some empty cells
    "c2_tiny": BrownianConfig("Brownian_tiny", 4096, (16, 16, 16), 11),
    # weak scaling over z-slabs: 1M agents and 64 cell planes per GPU
    # weak-scaling variants: 1M agents and 64 cell planes per rank; per-rank buffers = population + 64K of
    # headroom (replayed iterations launch for the buffer capacity, so slack costs empty blocks)
    "c2_w2": BrownianConfig("Brownian_w2", 2 << 20, (64, 64, 128), 42, buffer=1114112, world=2),
    "c2_w4": BrownianConfig("Brownian_w4", 4 << 20, (64, 64, 256), 42, buffer=1114112, world=4),
    "c2_w8": BrownianConfig("Brownian_w8", 8 << 20, (64, 64, 512), 42, buffer=1114112, world=8),
    # small 2-slab parity case
    "c2_s2": BrownianConfig("Brownian_s2", 1 << 15, (16, 16, 32), 13, buffer=24576, world=2),
    # ... and with distinct lower / upper neighbours (4 and 8 slabs of 16 planes)
    "c2_s4": BrownianConfig("Brownian_s4", 1 << 16, (16, 16, 64), 17, buffer=24576, world=4),
    "c2_s8": BrownianConfig("Brownian_s8", 1 << 17, (16, 16, 128), 19, buffer=24576, world=8),
}

BROWNIAN_XML = os.path.join(build.MODELS, "Brownian", "src", "model", "XMLModelFile.xml")


def brownian_xml(cfg: BrownianConfig) -> str:
    """XML-only scaling of models/Brownian: bufferSize, bounds, DOMAIN_* (functions.c untouched)."""
    if cfg.name == "Brownian":
        return BROWNIAN_XML
    dx, dy, dz = cfg.dims
    rep = {
        "<gpu:bufferSize>1048576</gpu:bufferSize>": f"<gpu:bufferSize>{cfg.buffer_size}</gpu:bufferSize>",
        "<gpu:xmax>64</gpu:xmax>": f"<gpu:xmax>{dx}</gpu:xmax>",
        "<gpu:ymax>64</gpu:ymax>": f"<gpu:ymax>{dy}</gpu:ymax>",
        "<gpu:zmax>64</gpu:zmax>": f"<gpu:zmax>{dz}</gpu:zmax>",
        "<name>DOMAIN_X</name>\n        <defaultValue>64.0</defaultValue>": f"<name>DOMAIN_X</name>\n        <defaultValue>{dx}.0</defaultValue>",
        "<name>DOMAIN_Y</name>\n        <defaultValue>64.0</defaultValue>": f"<name>DOMAIN_Y</name>\n        <defaultValue>{dy}.0</defaultValue>",
        "<name>DOMAIN_Z</name>\n        <defaultValue>64.0</defaultValue>": f"<name>DOMAIN_Z</name>\n        <defaultValue>{dz}.0</defaultValue>",
        "<name>Brownian</name>": f"<name>{cfg.name}</name>",
    }
    dst = os.path.join(build.GEN, cfg.name, "src", "model", "XMLModelFile.xml")
    build.write_scaled_xml(BROWNIAN_XML, dst, rep)
    # the scripts are shared, not copied: the build is pointed at the original functions.c
    return dst


def brownian_function_files():
    return [os.path.join(build.MODELS, "Brownian", "src", "model", "functions.c")]


def brownian_state(cfg: BrownianConfig, count: int = None) -> Dict[str, np.ndarray]:
    n = cfg.n if count is None else count
    u = uniform24(cfg.seed, 3 * n)
    ext = np.array(cfg.dims, dtype=np.float32)
    return {
        "id": np.arange(n, dtype=np.int32),
        "x": (u[0::3] * ext[0]).astype(np.float32),
        "y": (u[1::3] * ext[1]).astype(np.float32),
        "z": (u[2::3] * ext[2]).astype(np.float32),
        "nbr": np.zeros(n, dtype=np.int32),
    }


def write_states_xml(path: str, agent_name: str, columns: Dict[str, np.ndarray], env: Dict[str, float] = None) -> str:
    """0.xml in the reference's layout (io.xslt:124-200) with round-trippable `%.9g` floats."""
    names = list(columns.keys())
    n = len(columns[names[0]])
    with open(path, "w") as fh:
        fh.write("<states>\n<itno>0</itno>\n")
        if env:
            fh.write("<environment>\n")
            for k, v in env.items():
                fh.write(f"<{k}>{v!r}</{k}>\n")
            fh.write("</environment>\n")
        # one %-template per agent record, rows formatted a chunk at a time (a 16M-agent file is 2 GB)
        floating = {k: np.issubdtype(columns[k].dtype, np.floating) for k in names}
        template = (f"<xagent>\n<name>{agent_name}</name>\n" +
                    "".join(f"<{k}>%{'.9g' if floating[k] else 'd'}</{k}>\n" for k in names) + "</xagent>\n")
        for lo in range(0, n, 1 << 16):
            hi = min(n, lo + (1 << 16))
            cols = [columns[k][lo:hi].astype(np.float64 if floating[k] else np.int64).tolist() for k in names]
            fh.write("".join(template % row for row in zip(*cols)))
        fh.write("</states>\n")
    return path


# --------------------------------------------------------------------------------------------------
# BASELINE configs C3 / C4 / C5' (SURVEY.md 8d): the reference's own example scripts, XML-only scaling
# --------------------------------------------------------------------------------------------------
REFERENCE_EXAMPLES = "/root/reference/examples"


@dataclass
class ModelWorkload:
    key: str
    name: str                 # plugin / oracle name
    label: str
    xml_source: str           # XMLModelFile.xml the variant is scaled from
    replacements: Dict[str, str]
    agent: str
    state: str
    n: int
    b_alg: float              # algorithmic bytes per agent-step
    iterations: int           # iterations the BASELINE row asks for
    function_files: tuple = ()
    golden: str = ""          # shipped-input configs: tests/golden/<golden>_initial.npz (the example's own 0.xml as arrays)
    tiles: int = 0            # > 0: the shipped population replicated on a tiles x tiles grid of its own domain (x, y shifted per tile)
    tile_width: float = 0.0
    script_patch: tuple = ()  # ((regex, replacement), ...): constants of the script that XML-only scaling cannot reach; applied to a
                              # transient copy under _gen/ that is compiled and removed (never committed, SURVEY.md 8d "documented deviation")

    def xml(self) -> str:
        if not self.replacements:
            return self.xml_source
        dst = os.path.join(build.GEN, self.name, "src", "model", "XMLModelFile.xml")
        return build.write_scaled_xml(self.xml_source, dst, self.replacements)

    def functions(self):
        if self.function_files:
            return list(self.function_files)
        src = os.path.join(os.path.dirname(self.xml_source), "functions.c")
        if not self.script_patch:
            return [src]
        import re
        text = open(src, errors="replace").read()
        for pattern, repl in self.script_patch:
            text, hits = re.subn(pattern, repl, text)
            if hits != 1:
                raise build.BuildError(f"{self.key}: script patch /{pattern}/ matched {hits} times in {src}")
        dst = self.patched_script()
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with open(dst, "w") as fh:
            fh.write(text)
        return [dst]

    def patched_script(self) -> str:
        return os.path.join(build.GEN, self.name, "src", "model", "functions.c")

    def remove_patched_script(self) -> None:
        if self.script_patch and os.path.exists(self.patched_script()):
            os.remove(self.patched_script())


def _boids_state(n: int, seed: int = 43) -> Dict[str, np.ndarray]:
    u = uniform24(seed, 6 * n) - np.float32(0.5)
    return {"id": np.arange(n, dtype=np.int32), "x": u[0::6].copy(), "y": u[1::6].copy(), "z": u[2::6].copy(),
            "fx": u[3::6].copy(), "fy": u[4::6].copy(), "fz": u[5::6].copy()}


def _sph_state(side: int, seed: int = 44) -> Dict[str, np.ndarray]:
    """side^3 particles, one per partition cell of a side^3 grid over [-0.5, 0.5)^3, jittered inside the cell."""
    n = side ** 3
    i = np.arange(n, dtype=np.int64)
    u = uniform24(seed, 3 * n)
    h = np.float32(1.0 / side)
    cols = {"id": i.astype(np.int32)}
    for k, (name, idx) in enumerate((("x", i % side), ("y", (i // side) % side), ("z", i // (side * side)))):
        cols[name] = ((idx.astype(np.float32) + np.float32(0.25) + np.float32(0.5) * u[k::3]) * h - np.float32(0.5)).astype(np.float32)
    for name in ("dx", "dy", "dz", "fx", "fy", "fz", "density", "pressure"):
        cols[name] = np.zeros(n, dtype=np.float32)
    cols["isStatic"] = np.zeros(n, dtype=np.int32)
    return cols


def _birthdeath_state(n: int, seed: int = 45) -> Dict[str, np.ndarray]:
    u = uniform24(seed, 3 * n) * np.float32(32.0)
    return {"id": np.arange(n, dtype=np.int32), "x": u[0::3].copy(), "y": u[1::3].copy(), "z": u[2::3].copy(),
            "age": np.zeros(n, dtype=np.int32), "crowd": np.zeros(n, dtype=np.int32)}


_BOIDS_XML = os.path.join(REFERENCE_EXAMPLES, "Boids_Partitioning", "src", "model", "XMLModelFile.xml")
_SPH_XML = os.path.join(REFERENCE_EXAMPLES, "SmoothedParticleHydrodynamics", "src", "model", "XMLModelFile.xml")
_BD_XML = os.path.join(build.MODELS, "BirthDeath", "src", "model", "XMLModelFile.xml")

MODEL_WORKLOADS = {
    # C3: Boids_Partitioning functions.c unchanged; bufferSize 4 194 304, radius 1/128 -> 128^3 cells, ~2 agents / cell
    "c3": ModelWorkload("c3", "Boids_4M", "C3 Boids_Partitioning (reference functions.c), 4 194 304 agents, 128^3 cells, 27-cell neighbour search",
                        _BOIDS_XML, {"<gpu:bufferSize>2048</gpu:bufferSize>": "<gpu:bufferSize>4194304</gpu:bufferSize>",
                                     "<gpu:radius>0.1</gpu:radius>": "<gpu:radius>0.0078125</gpu:radius>"},
                        "Boid", "default", 1 << 22, 260.0, 100),
    # C4: SmoothedParticleHydrodynamics functions.c unchanged; 16 777 216 particles, radius 1/256 -> 256^3 cells (a
    # throughput case: the script's 1 m box cannot hold 16M particles at rest density)
    "c4": ModelWorkload("c4", "SPH_16M", "C4 SmoothedParticleHydrodynamics (reference functions.c), 16 777 216 particles, 256^3 cells, "
                        "two spatial messages, function conditions",
                        _SPH_XML, {"<gpu:bufferSize>131072</gpu:bufferSize>": "<gpu:bufferSize>16777216</gpu:bufferSize>",
                                   "<gpu:radius>0.02</gpu:radius>": "<gpu:radius>0.00390625</gpu:radius>"},
                        "FluidParticle", "default", 1 << 24, 556.0, 10),
    # a 2M-particle variant of the same (128^3 cells) that the CPU oracle finishes in seconds: parity at scale
    "c4_2m": ModelWorkload("c4_2m", "SPH_2M", "SmoothedParticleHydrodynamics, 2 097 152 particles, 128^3 cells",
                           _SPH_XML, {"<gpu:bufferSize>131072</gpu:bufferSize>": "<gpu:bufferSize>2097152</gpu:bufferSize>",
                                      "<gpu:radius>0.02</gpu:radius>": "<gpu:radius>0.0078125</gpu:radius>"},
                           "FluidParticle", "default", 1 << 21, 556.0, 10),
    # ... and 1 000 000 particles at the SHIPPED radius 0.02 (50^3 cells, 8 particles per cell): the parity case at scale
    "c4_1m": ModelWorkload("c4_1m", "SPH_1M", "SmoothedParticleHydrodynamics, 1 000 000 particles, shipped radius 0.02 (50^3 cells)",
                           _SPH_XML, {"<gpu:bufferSize>131072</gpu:bufferSize>": "<gpu:bufferSize>1048576</gpu:bufferSize>"},
                           "FluidParticle", "default", 1000000, 556.0, 10),
    # C1: CirclesPartitioning_float exactly as shipped (1024 agents, 8x8x1 cells; the reference's own CPU-runnable case --
    # launch-latency-bound at this size: reported, no roofline expectation)
    "c1": ModelWorkload("c1", "CirclesPartitioning_float", "C1 CirclesPartitioning_float (reference model, scripts and shipped 0.xml), "
                        "1024 agents, 8x8x1 cells", os.path.join(REFERENCE_EXAMPLES, "CirclesPartitioning_float", "src", "model", "XMLModelFile.xml"),
                        {}, "Circle", "default", 1024, 157.0, 100, golden="CirclesPartitioning_float"),
    # C5 at its shipped size (598 keratinocytes, growing): globalCondition, births, rand48, two spatial messages.  The 2M-cell
    # variant the BASELINE table names is degenerate on this script (its hard-coded 500-unit surface makes every cell read
    # ~40 % of all messages), so the shipped case is what is timed; no roofline claim
    "c5": ModelWorkload("c5", "Keratinocyte", "C5 Keratinocyte (reference model, scripts and shipped 0.xml), 598 cells growing, "
                        "two spatial messages, globalCondition, births", os.path.join(REFERENCE_EXAMPLES, "Keratinocyte", "src", "model", "XMLModelFile.xml"),
                        {}, "keratinocyte", "default", 598, 0.0, 200, golden="Keratinocyte"),
    # C5 at the size the BASELINE table names: the shipped colony (598 cells on a 500-unit surface) replicated on a 58 x 58 grid of
    # surfaces = 2 011 672 cells on a 29 000-unit surface.  XML-only: bufferSize, message bounds (location: 232 x 232 x 4 cells, force:
    # 1856 x 1856 x 32 = 110 M cells, 27-bit keys).  The surface width is a #define of the script, so the plugin is compiled from a
    # transient copy with that one constant changed -- the deviation SURVEY.md 8(d) documents for a physically sane 2M run.
    "c5_2m": ModelWorkload("c5_2m", "Keratinocyte_2M", "C5 Keratinocyte (reference model and scripts, SURFACE_WIDTH scaled), 2 011 672 cells = the shipped "
                           "colony tiled 58 x 58, two spatial messages (force grid 1856 x 1856 x 32 cells), globalCondition, births",
                           os.path.join(REFERENCE_EXAMPLES, "Keratinocyte", "src", "model", "XMLModelFile.xml"),
                           {"<gpu:bufferSize>8192</gpu:bufferSize>": "<gpu:bufferSize>4194304</gpu:bufferSize>",
                            "<gpu:xmax>500</gpu:xmax>": "<gpu:xmax>29000</gpu:xmax>", "<gpu:ymax>500</gpu:ymax>": "<gpu:ymax>29000</gpu:ymax>"},
                           "keratinocyte", "default", 58 * 58 * 598, 0.0, 10, golden="Keratinocyte", tiles=58, tile_width=500.0,
                           script_patch=((r"#define\s+SURFACE_WIDTH\s+500\.0f", "#define SURFACE_WIDTH 29000.0f"),)),
    # C5': BirthDeath (authored): death compaction + births + rand48 on the spatial path, 524 288 -> <= 2 097 152 agents
    "c5p": ModelWorkload("c5p", "BirthDeath", "C5' BirthDeath (authored), 524 288 agents growing in a 2 097 152 buffer, 32^3 cells, "
                         "death compaction + births + rand48", _BD_XML, {}, "Cell", "default", 1 << 19, 170.0, 200),
}


def model_state(w: ModelWorkload) -> Dict[str, np.ndarray]:
    if w.golden:
        path = os.path.join(build.ROOT, "tests", "golden", w.golden + "_initial.npz")
        st = dict(np.load(path))
        if w.tiles:
            T = w.tiles
            n0 = len(st[f"{w.agent}/id"])
            tile = np.repeat(np.arange(T * T, dtype=np.int64), n0)
            out = {k: v for k, v in st.items() if not k.startswith(w.agent + "/")}
            for k, v in st.items():
                if k.startswith(w.agent + "/"):
                    out[k] = np.tile(v, T * T)
            out[f"{w.agent}/x"] = (out[f"{w.agent}/x"] + (tile % T).astype(np.float32) * np.float32(w.tile_width)).astype(np.float32)
            out[f"{w.agent}/y"] = (out[f"{w.agent}/y"] + (tile // T).astype(np.float32) * np.float32(w.tile_width)).astype(np.float32)
            out[f"{w.agent}/id"] = (out[f"{w.agent}/id"].astype(np.int64) + tile * 8192).astype(np.int32)   # shipped ids are < 8192
            st = out
        return st
    if w.key == "c3":
        return _boids_state(w.n)
    if w.key == "c4":
        return _sph_state(256)
    if w.key == "c4_2m":
        return _sph_state(128)
    if w.key == "c4_1m":
        return _sph_state(100)
    if w.key == "c5p":
        return _birthdeath_state(w.n)
    raise KeyError(w.key)


def start_model(sim, w: ModelWorkload, st: Dict[str, np.ndarray]) -> None:
    """Initial population of a model workload on a simulation object (CUDA path or oracle)."""
    if not w.golden:
        sim.set_agents(w.agent, w.state, st)
        return
    for k, v in st.items():            # the shipped 0.xml: environment values, every agent type, then the init functions
        if k.startswith("env/"):
            sim.set_constant(k[4:], v)
    for a in sim.model.agents:
        sim.set_agents(a.name, a.states[a.initial_state], {v.name: st[f"{a.name}/{v.name}"] for v in a.vars})
    sim.run_init_functions()
