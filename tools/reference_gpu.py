#!/usr/bin/env python
"""The reference's own CUDA beside ours on one B200: same model, same 0.xml, same iteration count (SURVEY.md 8d,
"reference side-by-side").

    # in the build container (needs /root/reference: templates + glm): nvcc cross-compiles without a GPU
    python tools/reference_gpu.py --build --workload c2
    # on the GPU box (the binaries travel in oracle/_ref/, git-ignored)
    gpurun --timeout 900 -- 'python tools/reference_gpu.py --run --workload c2 --iterations 100 > gpurun_out/reference_gpu.json'

--build   oracle/build_ref.build_ref_device: the reference's templates expanded (oracle/xslt_subset.py), texture references
          rewritten to plain cached loads (the only edit; CUDA 12 has no texture references), nvcc -O3 for sm_100a with the
          reference's default options (FAST_ATOMIC_SORTING on) -> oracle/_ref/<name>/ref_gpu_<name>, the reference's
          console program.
--run     writes the workload's synthetic 0.xml (%.9g floats), runs `ref_gpu_<name> 0.xml N 0 0` (no XML output) for the
          reference's own "Total Processing time", runs it again with one XML dump at iteration N, runs this repository's
          console program (ev_diffusion_b200/_lib/fgpu_console) on the same plugin-independent input the same way, and
          prints one JSON line: both processing times, agent-steps/s, and whether the two final iteration files are
          byte-identical (integers and %f-formatted floats; the Brownian model's results do not depend on message order,
          so the reference's atomic sort does not matter here).

Nothing under tests/ or bench.py depends on this tool; it is a measurement aid for the next round.  It has only been
exercised up to the launch of the binaries (the build container has no GPU): treat its first GPU run as its test.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ev_diffusion_b200 import build, workloads as W  # noqa: E402


def _paths(cfg):
    from oracle import build_ref
    return build_ref.ref_device_path(cfg.name), build.model_path(cfg.name), os.path.join(build.LIB, "fgpu_console")


def do_build(cfg) -> None:
    from oracle import build_ref
    t = time.time()
    out = build_ref.build_ref_device(W.brownian_xml(cfg), cfg.name, function_dir=os.path.dirname(W.brownian_function_files()[0]))
    print("built %s in %.0f s" % (out, time.time() - t))


def _processing_ms(cmd, cwd) -> float:
    proc = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    m = re.search(r"Total Processing time: ([0-9.eE+-]+) \(ms\)", proc.stdout)
    if proc.returncode != 0 or not m:
        raise RuntimeError("%s failed (exit %d):\n%s" % (" ".join(cmd), proc.returncode, proc.stdout[-3000:]))
    return float(m.group(1))


def do_run(cfg, iterations: int, repeats: int) -> None:
    ref_bin, plugin, console = _paths(cfg)
    for p in (ref_bin, plugin, console):
        if not os.path.exists(p):
            sys.exit("missing %s: run --build (and __graft_entry__.build()) in the build container first" % p)
    n = cfg.n
    out = {"workload": cfg.name, "agents": n, "iterations": iterations}
    with tempfile.TemporaryDirectory() as tmp:
        sides = {"reference": [ref_bin], "ours": [console, plugin]}
        shared = W.write_states_xml(os.path.join(tmp, "input.xml"), "Particle", W.brownian_state(cfg))   # 2 GB at 16M agents
        for side, head in sides.items():
            d = os.path.join(tmp, side)
            os.makedirs(d)
            inp = os.path.join(d, "0.xml")           # both programs write their iteration files beside their input
            os.symlink(shared, inp)
            times = [_processing_ms(head + [inp, str(iterations), "0", "0"], d) for _ in range(repeats)]
            out[side] = {"processing_ms": times, "best_ms_per_iteration": min(times) / iterations,
                         "agent_steps_per_s": n * iterations / (min(times) * 1e-3)}
            _processing_ms(head + [inp, str(iterations), "0", str(iterations)], d)          # one dump at the end
            final = os.path.join(d, "%d.xml" % iterations)
            out[side]["final_file_bytes"] = os.path.getsize(final) if os.path.exists(final) else None
        a, b = (os.path.join(tmp, s, "%d.xml" % iterations) for s in ("reference", "ours"))
        if os.path.exists(a) and os.path.exists(b):
            with open(a, "rb") as fa, open(b, "rb") as fb:
                out["final_files_identical"] = fa.read() == fb.read()
    out["ours_over_reference"] = out["ours"]["agent_steps_per_s"] / out["reference"]["agent_steps_per_s"]
    print(json.dumps(out))


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--workload", default="c2", choices=sorted(k for k, c in W.BROWNIAN.items() if c.world == 1))
    ap.add_argument("--build", action="store_true")
    ap.add_argument("--run", action="store_true")
    ap.add_argument("--iterations", type=int, default=100)
    ap.add_argument("--repeats", type=int, default=3)
    a = ap.parse_args()
    cfg = W.BROWNIAN[a.workload]
    if a.build:
        do_build(cfg)
    if a.run:
        do_run(cfg, a.iterations, a.repeats)
    if not (a.build or a.run):
        ap.print_help()


if __name__ == "__main__":
    main()
