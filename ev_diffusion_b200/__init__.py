"""The FLAME GPU 1.5 simulation step for B200 (sm_100a): model loader, plugin generator and the ctypes mirror of the C ABI.

    from ev_diffusion_b200 import build_model, Simulation
    sim = Simulation(build_model("model/XMLModelFile.xml", "MyModel"))      # nvcc: functions.c + generated glue -> plugin .so
    sim.load_states("iterations/0.xml"); sim.step(100); sim.agents("Particle")

Everything that runs per iteration lives in csrc/ (libfgpu_rt.so + one plugin per model); there is no CPU fallback.
Submodules load lazily: importing the package touches neither nvcc nor the CUDA runtime.
"""
__all__ = ["Simulation", "FgpuError", "build_model", "build_runtime", "parse_model"]


def __getattr__(name):
    if name in ("Simulation", "FgpuError"):
        from . import runtime
        return getattr(runtime, name)
    if name in ("build_model", "build_runtime"):
        from . import build
        return getattr(build, name)
    if name == "parse_model":
        from . import xmml
        return xmml.parse_model
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
