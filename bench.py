#!/usr/bin/env python
"""bench.py -- agent-steps/s of the FLAME GPU 1.5 simulation step (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2_16m|c2] [--impl ours|reference]

One "step" = one simulation iteration (singleIteration(), simulation.xslt:878-975) of
the Brownian random-walk model over its whole population.  The default workload is the
north-star configuration C2-16M (16 777 216 float agents, 128 x 128 x 256 spatially
partitioned cells, rand48, wall reflection); with N GPUs the SAME population is cut into
N z-slabs (strong scaling).  `--workload c2` is BASELINE config C2 (1 048 576 agents,
64^3 cells; N GPUs: 1M agents per GPU, weak scaling).  Prints ONE JSON line (rank 0).

 value      whole-job agent-steps/s, inputs resident in HBM, device time from CUDA
            events on the simulation stream, max over ranks; every timed iteration
            starts from a flushed L2 (a 256 MiB scratch buffer is rewritten between
            iterations, outside the timed span)
 e2e        the same metric through the public API with HOST buffers: per step the
            population is uploaded from pinned host memory, one iteration runs, and
            the resulting x,y,z,nbr columns are read back
 roofline   dominant kernel, algorithmic bytes / CUDA-event time
 cpu_baseline  the CPU oracle (host-C execution of the same functions.c) on a bounded
            sample of the same workload: OpenMP on all cores (`value`) and one thread
 reference_cuda  (N=1) the reference's own generated CUDA, built for sm_100a by
            oracle/build_ref.py --device, run on the same workload in its own process
 parity     untimed, after the measurement: N=1 -- one iteration of the bit-exact build of the same
            model against one oracle iteration on the same input (message order, every variable, seeds),
            and the timed (FMA-contracted) build against the same oracle iteration; N>1 -- id multiset
            conserved, every agent on its owner, brute-force neighbour count on a 64K sample
 --impl reference  the reference has no CPU path; the arm times the host-C oracle on
            all host cores, as the task statement prescribes for this tier.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from ev_diffusion_b200 import build as B, workloads as W  # noqa: E402

METRIC = "agent-steps/sec"
FLUSH_BYTES = 256 << 20

# algorithmic bytes per agent-step, SURVEY.md 8(d) row C2 (the rule is applied per function)
B_ALG = {
    "func:output_location": 32.0,   # 16 r (id,x,y,z) + 16 w (message)
    "build:location": 94.0,         # 12 (key source) + 3 passes x 16 + 16 r + 16 w (payload) + p=2
    "func:count_neighbours": 38.0,  # 16 r + 16 (messages, perfect reuse) + 4 w + p=2
    "func:move": 40.0,              # 12 r + 8 + 8 (seed) + 12 w
}
B_ALG_ITERATION = 204.0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons during the timed region (NVML)."""

    def __init__(self, index=0, period=0.1):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.period, self.index, self._halt = period, index, threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            # first-call costs of the two queries are paid here, outside the timed region: an NVML query takes a
            # driver lock and was seen to stall one timed iteration by ~2 ms now and then, so the thread samples
            # sparsely (first sample 5 ms into the region, then every `period`)
            pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons"):
                pynvml.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            else:
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake": 0x80, "sw_power_cap": 0x4}
        if self._halt.wait(0.005):
            return
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._halt.wait(self.period)

    def stop(self):
        self._halt.set()
        if self.ok:
            self.join(timeout=1.0)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def slab_key(workload, world):
    return workload if world == 1 else f"{workload}_w{world}"


def scaling_of(workload):
    return "strong" if workload == "c2_16m" else "weak"


def run_reference(args, cfg):
    """CPU arm: the host-C oracle (OpenMP, all host cores) on the same workload (whole-job population)."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from oracle import gen_oracle
    from oracle.oracle import OracleSimulation
    cores = os.cpu_count() or 1
    so = gen_oracle.oracle_path(cfg.name)
    if not os.path.exists(so):
        so = gen_oracle.build_oracle(W.brownian_xml(cfg), cfg.name, function_files=W.brownian_function_files())
    sim = OracleSimulation(so, threads=cores)
    st = W.brownian_state(cfg)
    sim.set_agents("Particle", "default", st)
    sim.step(args.warmup)
    t0 = time.perf_counter()
    sim.step(args.steps)
    dt = time.perf_counter() - t0
    value = cfg.n * args.steps / dt
    same = scaling_of(args.workload) == "strong" or args.gpus == 1
    line = {
        "metric": METRIC, "value": value, "unit": "agent-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": scaling_of(args.workload),
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": {"workload": workload_name(args, cfg), "agents": cfg.n, "cells": list(cfg.dims),
                   "population_of_gpu_arm": cfg.n if same else cfg.n * args.gpus,
                   "same_population_as_gpu_arm": same},
        "cpu_baseline": {"value": value, "unit": "agent-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} full iterations of the {cfg.n}-agent workload, OpenMP {cores} threads" +
                                   ("" if same else f" (one slab's worth of the {args.gpus}-GPU weak-scaling job: same density, "
                                                    "same cells per slab; agent-steps/s normalises)")},
        "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference has no CPU path: this arm is the host-C oracle (port) executing the same functions.c with OpenMP. "
                "The reference's own generated code also runs on the CPU here (oracle/build_ref.py, one host thread, "
                "0.16 M agent-steps/s on C2) and computes the same bytes (tests/test_oracle_vs_reference.py); "
                "the port is timed because it is the faster CPU arm.  The reference's CUDA build for sm_100a is the "
                "reference_cuda object of the GPU arm's line.",
    }
    print(json.dumps(line))


def workload_name(args, cfg):
    return {"c2": "C2 Brownian random walk, 1M float agents, 64^3 spatially partitioned cells, rand48, wall reflection",
            "c2_16m": "C2-16M Brownian random walk (north-star config), 16 777 216 float agents, 128x128x256 spatially "
                      "partitioned cells, rand48, wall reflection"}.get(args.workload, cfg.name)


def cpu_baseline(cfg, budget_s=12.0):
    """north_star: "a straightforward single-threaded and OpenMP host-C execution of the same functions.c on the box's
    host cores, with the core count stated".  `value` is the OpenMP figure (the stronger baseline).  The one-thread
    figure is taken on the 1M-agent C2 population (same model, same density) when the workload is larger, so that the
    leg stays bounded: one thread needs ~28 s per 16M-agent iteration."""
    from oracle import gen_oracle
    from oracle.oracle import OracleSimulation
    cores = os.cpu_count() or 1

    def timed(c, threads, budget, most, least, warm):
        sim = OracleSimulation(gen_oracle.oracle_path(c.name), threads=threads)
        sim.set_agents("Particle", "default", W.brownian_state(c))
        if warm:
            sim.step(1)
        t0 = time.perf_counter()
        iters = 0
        while iters < most and (iters < least or time.perf_counter() - t0 < budget):
            sim.step(1)
            iters += 1
        dt = time.perf_counter() - t0
        sim.close()
        return iters, dt

    iters, dt = timed(cfg, cores, budget_s, 20, 2, True)
    c1 = cfg if cfg.n <= (1 << 20) else W.BROWNIAN["c2"]
    it1, dt1 = timed(c1, 1, budget_s / 2, 4, 1, False)
    return {"value": cfg.n * iters / dt, "unit": "agent-steps/s", "cores": cores, "kind": "port",
            "sample": f"{iters} full iterations of the {cfg.n}-agent workload ({dt:.1f} s), OpenMP {cores} threads",
            "single_thread": {"value": c1.n * it1 / dt1, "unit": "agent-steps/s", "cores": 1,
                              "sample": f"{it1} full iterations of the {c1.n}-agent {c1.name} population ({dt1:.1f} s), one thread"}}


def reference_cuda(cfg, sim_for_xml=None, iterations=None, timeout_s=420.0):
    """north_star: "the reference's own CUDA build on one B200 ... timed in the same run".  oracle/build_ref.py --device
    expands the reference's templates and compiles its console program for sm_100a (texture references rewritten to
    cached loads, nothing else; built in the build container, it travels in oracle/_ref/).  Here it is run as its own
    process on the workload's 0.xml (written by this repository's own state writer with round-trippable %.9g floats),
    XML output off, and its own "Total Processing time" is reported.  A baseline, never part of the product path; any
    failure is recorded in the JSON and changes nothing else."""
    import re
    import subprocess
    import tempfile
    exe = os.path.join(ROOT, "oracle", "_ref", cfg.name, f"ref_gpu_{cfg.name}")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/%s/ref_gpu_%s is not built (python oracle/build_ref.py --device, needs /root/reference)" % (cfg.name, cfg.name)}
    big = cfg.n > (1 << 21)
    iterations = iterations or (30 if big else 100)
    runs = 1 if big else 2          # the reference's reader takes most of a minute on a 2 GB 0.xml
    try:
        with tempfile.TemporaryDirectory() as tmp:
            t0 = time.perf_counter()
            if sim_for_xml is not None:                          # the runtime's C writer: seconds instead of minutes at 16M agents
                sim_for_xml.set_option("xml_float_digits", 9)
                sim_for_xml.set_agents("Particle", "default", W.brownian_state(cfg))
                inp = sim_for_xml.save_states(tmp + os.sep, 0)
            else:
                inp = W.write_states_xml(os.path.join(tmp, "0.xml"), "Particle", W.brownian_state(cfg))
            t_xml = time.perf_counter() - t0
            best = None
            t0 = time.perf_counter()
            for _ in range(runs):                                # with two runs the first also warms the driver's caches
                proc = subprocess.run([exe, inp, str(iterations), "0", "0"], cwd=tmp, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True, timeout=timeout_s)
                m = re.search(r"Total Processing time: ([0-9.eE+-]+) \(ms\)", proc.stdout)
                if proc.returncode != 0 or not m:
                    lines = [x for x in proc.stdout.strip().splitlines() if x.strip()]
                    why = next((x for x in lines if "rror" in x), lines[-1] if lines else "")
                    return {"unavailable": "exit %d: %s" % (proc.returncode, why[:200])}
                ms = float(m.group(1))
                best = ms if best is None else min(best, ms)
            t_run = time.perf_counter() - t0
        return {"value": cfg.n * iterations / (best * 1e-3), "unit": "agent-steps/s", "ms_per_step": best / iterations,
                "steps": iterations, "agents": cfg.n,
                "kind": "reference templates expanded, nvcc sm_100a, FAST_ATOMIC_SORTING (its default)",
                "timing": "the program's own cudaEvent pair around its iteration loop (main.xslt:415-446), best of %d run(s), "
                          "warm L2, per-iteration printf to a pipe" % runs,
                "wall_s": {"write_0xml": round(t_xml, 1), "reference_process": round(t_run, 1)}}
    except Exception as e:                                       # a baseline must never take the bench down
        return {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}


def parity_single_gpu(cfg, timed_plugin, device, oracle_threads):
    """One iteration on the same seeded input: the bit-exact build (-fmad=false) against the CPU oracle -- message
    order, every agent variable bit for bit, every rand48 seed -- and the timed build (-fmad=true, the reference's
    own setting) against the same oracle iteration: integer outputs that do not pass through float rounding
    (message order, seeds) equal, nbr equal except where a distance is within rounding of the radius, positions
    within 1 ulp-scale of the FMA contraction."""
    from ev_diffusion_b200 import check
    from ev_diffusion_b200.runtime import Simulation
    from oracle import gen_oracle
    from oracle.oracle import OracleSimulation
    out = {}
    st = W.brownian_state(cfg)
    orc = OracleSimulation(gen_oracle.oracle_path(cfg.name), threads=oracle_threads)
    orc.set_agents("Particle", "default", st)
    orc.step(1)
    want = orc.agents("Particle")
    want_seeds = orc.seeds()
    want_order = orc.message_order("location")
    orc.close()
    exact = B.model_path(cfg.name + "_exact")
    ok = True
    if os.path.exists(exact):
        g = Simulation(exact, device=device)
        g.set_agents("Particle", "default", st)
        g.step(1)
        got = g.agents("Particle")
        e = {"message_order": bool(np.array_equal(g.message_order("location"), want_order)),
             "seeds": bool(np.array_equal(g.seeds(), want_seeds))}
        for k in want:
            e[k] = bool(np.array_equal(got[k].view(np.uint32), want[k].view(np.uint32)))
        g.close()
        out["exact_build_vs_oracle"] = e
        ok = ok and all(e.values())
    else:
        out["exact_build_vs_oracle"] = {"unavailable": exact + " not built"}
        ok = False
    g = Simulation(timed_plugin, device=device)
    g.set_agents("Particle", "default", st)
    g.step(1)
    got = g.agents("Particle")
    t = {"message_order": bool(np.array_equal(g.message_order("location"), want_order)),
         "seeds": bool(np.array_equal(g.seeds(), want_seeds)),
         "id": bool(np.array_equal(got["id"], want["id"])),
         "nbr_mismatch_fraction": float((got["nbr"] != want["nbr"]).mean())}
    g.close()
    err = max(float(np.abs(got[k].astype(np.float64) - want[k].astype(np.float64)).max()) for k in ("x", "y", "z"))
    t["max_abs_position_error"] = err
    t["position_tolerance"] = 2.0 ** -16      # one rounding of fma(SIGMA, 2u-1, p) at |p| < 256: half an ulp = 2^-17
    out["timed_build_vs_oracle"] = t
    ok = ok and t["message_order"] and t["seeds"] and t["id"] and t["nbr_mismatch_fraction"] < 1e-5 and err <= t["position_tolerance"]
    # the neighbour counts of the timed build against plain geometry on a sample (the check used at N > 1)
    smp = check.neighbour_sample_check(st, got, cfg.dims)
    out["neighbour_sample"] = smp
    ok = ok and smp["mismatches"] == 0
    out["iterations"] = 1
    out["agents"] = cfg.n
    return ok, out


def gather_population(sim, names, dist, device):
    """All ranks' state lists on rank 0 (padded all_gather over NCCL; check plumbing, untimed)."""
    import torch
    a = sim.agents("Particle")
    n = len(a["id"])
    world = dist.get_world_size()
    sizes = torch.zeros(world, dtype=torch.int64, device=device)
    sizes[dist.get_rank()] = n
    dist.all_reduce(sizes)
    cap = int(sizes.max().item())
    out = {}
    for k in names:
        src = a[k]
        buf = torch.zeros(cap, dtype=torch.int32, device=device)
        buf[:n] = torch.from_numpy(src.view(np.int32).copy()).to(device)
        parts = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf)
        if dist.get_rank() == 0:
            out[k] = np.concatenate([p[:int(sizes[r])].cpu().numpy() for r, p in enumerate(parts)]).view(src.dtype)
    owner = np.concatenate([np.full(int(sizes[r]), r, dtype=np.int32) for r in range(world)])
    return out, owner


def parity_slabs(cfg, sim, dist, device, world):
    """N > 1 (no oracle at this size, and a P-rank run is not bit-comparable to a 1-rank run: list order defines
    rand48 stream ownership): one more iteration, then on rank 0 -- id multiset conserved, every agent on the rank
    that owns its cell plane, and nbr of a 64K sample equal to a brute-force float32 count over the GLOBAL positions
    at the start of that iteration (needs halo planes incl. the wrap pair, migration and ownership to be right)."""
    from ev_diffusion_b200 import check
    before, _ = gather_population(sim, ("id", "x", "y", "z"), dist, device)
    sim.step(1)
    after, owner = gather_population(sim, ("id", "x", "y", "z", "nbr"), dist, device)
    if dist.get_rank() != 0:
        return True, {}
    out = {"ids_conserved": check.ids_conserved(after["id"], cfg.n),
           "agents_on_owner": check.owners_ok(after["z"], owner, cfg.dims, world),
           "agents_per_rank": [int((owner == r).sum()) for r in range(world)]}
    if out["ids_conserved"]:
        out["neighbour_sample"] = check.neighbour_sample_check(before, after, cfg.dims)
    ok = out["ids_conserved"] and out["agents_on_owner"] and out["neighbour_sample"]["mismatches"] == 0
    return bool(ok), out


def load_ncu_record(workload, kernel):
    """Per-launch DRAM traffic and issue statistics of the dominant kernel from this round's committed ncu capture
    (profiles/r02_traffic.json, written by tools/ncu_traffic.py from an `ncu --set full` raw page; it records the
    commit and command it came from).  Not measurable inside bench.py without a profiler."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as fh:
                rec = json.load(fh).get(workload, {}).get(kernel)
            if rec:
                rec = dict(rec)
                rec["source"] = "profiles/" + name
                return rec
        except (OSError, ValueError, KeyError):
            pass
    return None


# --------------------------------------------------------------------------------------------------
# BASELINE configs C3 / C4 / C5' (`--workload c3|c4|c4_2m|c5p`): the dynamic path (function conditions, two
# spatial messages, death compaction, births) at its BASELINE size.  Single GPU.
# --------------------------------------------------------------------------------------------------
def _population(sim):
    return sum(sim.agent_count(a.name, st) for a in sim.model.agents for st in range(len(a.states)))


def _all_states(sim):
    out = {}
    for a in sim.model.agents:
        for si, st in enumerate(a.states):
            if sim.agent_count(a.name, si) > 0:
                out[(a.name, st)] = sim.agents(a.name, si)
    return out


def run_model_workload_slabs(args, w, st, base, rank, world, local):
    """The same model and population cut into `world` z-slabs (strong scaling), one process per GPU: BASELINE config C4 is
    SPH "with slab domain decomposition at 2/4/8 B200".  Models with function conditions run the synchronous slab mode
    (populations read back after every compaction, like the reference's own scans)."""
    import torch
    import torch.distributed as dist
    from ev_diffusion_b200 import check, slabs
    from ev_diffusion_b200.runtime import Simulation, nccl_unique_id
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    sim = Simulation(B.model_path(w.name), device=local)
    for kv in args.option:
        k, v = kv.split("=")
        sim.set_option(k, int(v))
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(uid, 0)
    sim.slab_init(rank, world, uid.cpu().numpy().tobytes())
    m = next(x for x in sim.model.messages if x.partitioning == 1)
    plan = slabs.make_plan(m.dims, m.min_bounds, m.max_bounds, m.is_2d, rank, world)
    sim.set_agents(w.agent, w.state, slabs.partition_state(st, plan))

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()
        sim.sync()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        t = torch.tensor([x], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        return int(t.item())

    sim.step(args.warmup)
    barrier()
    n0 = allsum(_population(sim))
    launches0 = sim.launch_count
    sampler = ClockSampler(local)
    sampler.start()
    ms_each = sim.step_timed(args.steps, FLUSH_BYTES)
    barrier()
    clocks = sampler.stop()
    total_ms = allmax(float(ms_each.sum()))
    n1 = allsum(_population(sim))
    value = 0.5 * (n0 + n1) * args.steps / (total_ms * 1e-3)
    reps = min(args.steps, 5)
    sim.profile_iterations(reps)
    stage = dict(sim.profile_iterations(reps))
    peak, peak_src = peaks()
    ach = w.b_alg * value / world / 1e9
    # correctness of the final state: ids conserved over the ranks, every particle on the rank that owns its plane
    ids, zs = [], []
    for a in sim.model.agents:
        for si in range(len(a.states)):
            if sim.agent_count(a.name, si) > 0:
                cols = sim.agents(a.name, si)
                ids.append(cols["id"])
                zs.append(cols["z" if plan.axis == 2 else "y"])
    ids = np.concatenate(ids) if ids else np.zeros(0, np.int32)
    zs = np.concatenate(zs) if zs else np.zeros(0, np.float32)
    mine_ok = bool((slabs.owner_of(zs, plan) == rank).all())
    seen = torch.zeros(w.n, dtype=torch.int32, device="cuda")
    seen[torch.from_numpy(ids.astype(np.int64)).cuda()] += 1
    dist.all_reduce(seen)
    ids_ok = bool((seen == 1).all().item())
    own = torch.tensor([1 if mine_ok else 0], device="cuda")
    dist.all_reduce(own, op=dist.ReduceOp.MIN)
    parity = {"ids_conserved": ids_ok, "agents_on_owner": bool(own.item() == 1)}
    line = dict(base, value=value, n_gpus=world, scaling="strong", steps=args.steps, warmup=args.warmup, ms_per_step=total_ms / args.steps,
                clocks=clocks, gpu_launches=int(sim.launch_count - launches0),
                roofline={"bound": "hbm", "kernel": "whole iteration", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                          "traffic": None, "peak_source": peak_src, "alg_bytes_per_agent_step": w.b_alg,
                          "stage_ms": {k: round(v, 5) for k, v in stage.items()}},
                parity_checked=bool(ids_ok and own.item() == 1), parity=parity, population={"start": n0, "end": n1},
                e2e=None)
    line["config"] = dict(line["config"], parallelism=f"{world} z-slabs of {m.dims[plan.axis] // world} cell planes (strong scaling), "
                          + ("peer-memory transport" if sim.slab_info().get("p2p") else "NCCL send/recv"),
                          l2="flushed between timed iterations (256 MiB scratch rewrite, untimed)")
    if rank == 0:
        print(json.dumps(line))
    sim.close()
    dist.destroy_process_group()


def run_model_workload(args):
    from oracle import gen_oracle
    from oracle.oracle import OracleSimulation
    w = W.MODEL_WORKLOADS[args.workload]
    rank, world, local = dist_env()
    if rank != 0 and (args.impl == "reference" or world == 1):
        return
    cores = os.cpu_count() or 1
    if w.n < 100000:
        cores = 1          # a 1024-agent iteration is microseconds of work: OpenMP fork/join over many threads only slows it down
    st = W.model_state(w)
    base = {"metric": METRIC, "unit": "agent-steps/s", "n_gpus": 1, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": w.label, "agents": w.n, "parallelism": "single GPU", "baseline_iterations": w.iterations}}

    def cpu_run(steps, warm):
        orc = OracleSimulation(gen_oracle.oracle_path(w.name), threads=cores)
        W.start_model(orc, w, st)
        if warm:
            orc.step(warm)
        n0 = _population(orc)
        t0 = time.perf_counter()
        orc.step(steps)
        dt = time.perf_counter() - t0
        n1 = _population(orc)
        orc.close()
        return 0.5 * (n0 + n1) * steps / dt, dt, steps

    if args.impl == "reference":
        steps = max(1, min(args.steps, 5))
        v, dt, k = cpu_run(steps, 1)
        line = dict(base, value=v, steps=k, warmup=1, ms_per_step=dt / k * 1e3, impl="reference", gpu_launches=0,
                    cpu_baseline={"value": v, "unit": "agent-steps/s", "cores": cores, "kind": "port",
                                  "sample": f"{k} full iterations of the workload, OpenMP {cores} threads"},
                    e2e={"value": v, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(line))
        return

    from ev_diffusion_b200.runtime import Simulation
    plugin = B.model_path(w.name)
    if not os.path.exists(plugin):
        raise SystemExit(f"{plugin} missing: run `python -c 'import __graft_entry__ as g; g.build()'` first (needs /root/reference)")
    if world > 1:
        if w.golden:
            raise SystemExit(f"--workload {w.key} is a shipped-size single-GPU config")
        run_model_workload_slabs(args, w, st, base, rank, world, local)
        return
    sim = Simulation(plugin, device=local)
    for kv in args.option:
        k, v = kv.split("=")
        sim.set_option(k, int(v))
    W.start_model(sim, w, st)
    sim.step(args.warmup)
    n0 = _population(sim)
    launches0 = sim.launch_count
    sampler = ClockSampler(local)
    sampler.start()
    ms_each = sim.step_timed(args.steps, FLUSH_BYTES)
    clocks = sampler.stop()
    launches = sim.launch_count - launches0
    n1 = _population(sim)
    total_ms = float(ms_each.sum())
    agent_steps = 0.5 * (n0 + n1) * args.steps            # populations change by a few percent per iteration at most
    value = agent_steps / (total_ms * 1e-3)
    reps = min(args.steps, 10)
    sim.profile_iterations(reps)
    stage = dict(sim.profile_iterations(reps))
    dom = max(stage, key=lambda k: stage[k])
    peak, peak_src = peaks()
    ach = w.b_alg * value / 1e9      # b_alg 0: a config SURVEY.md 8(d) lists without a byte count (C5), no roofline claim
    roofline = {"bound": "hbm", "kernel": "whole iteration (sum of its kernels; dominant stage: %s)" % dom, "achieved": ach, "peak": peak,
                "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src, "alg_bytes_per_agent_step": w.b_alg,
                "stage_ms": {k: round(v, 5) for k, v in stage.items()}, "dominant_stage": dom, "dominant_stage_ms": stage[dom]}
    # end to end: upload of the initial population, one iteration, read-back of every variable (host buffers)
    import torch
    pin = {k: torch.from_numpy(np.ascontiguousarray(v).copy()).pin_memory().numpy() for k, v in st.items() if not k.startswith("env/")}
    h2d = sum(v.nbytes for v in pin.values())
    e2e_steps = min(args.steps, 5)
    sim2 = Simulation(plugin, device=local)

    def upload(s_):
        if w.golden:      # the shipped input: every agent type into its initial state
            for a in s_.model.agents:
                s_.set_agents(a.name, a.states[a.initial_state], {v.name: pin[f"{a.name}/{v.name}"] for v in a.vars})
        else:
            s_.set_agents(w.agent, w.state, pin)
    W.start_model(sim2, w, st)
    sim2.step(1)
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(e2e_steps):
        upload(sim2)
        sim2.step(1)
        got = _all_states(sim2)
        d2h = sum(v.nbytes for cols in got.values() for v in cols.values())
    e2e_dt = time.perf_counter() - t0
    sim2.close()
    e2e = {"value": w.n * e2e_steps / e2e_dt, "unit": "agent-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps}
    # parity: the bit-exact build against the oracle on the same input, where the oracle finishes in seconds
    parity_ok, parity = None, None
    exact = B.model_path(w.name + "_exact")
    if not args.no_parity and os.path.exists(exact) and os.path.exists(gen_oracle.oracle_path(w.name)):
        iters = 1 if w.key.startswith("c4") else 2     # SPH at this density is stiff: rounding differences of powf grow within a step or two
        orc = OracleSimulation(gen_oracle.oracle_path(w.name), threads=cores)
        g = Simulation(exact, device=local)
        for s_ in (orc, g):
            W.start_model(s_, w, st)
            s_.step(iters)
        a, b = _all_states(g), _all_states(orc)
        parity = {"iterations": iters, "lists": {}}
        parity_ok = set(a) == set(b)
        for key in a:
            if key not in b:
                continue
            same = {k: bool(a[key][k].shape == b[key][k].shape and np.array_equal(a[key][k].view(np.uint32), b[key][k].view(np.uint32)))
                    for k in a[key]}
            parity["lists"]["%s/%s" % key] = {"agents": int(len(next(iter(a[key].values())))), "bit_identical": all(same.values()),
                                              "differing": [k for k, v in same.items() if not v]}
            intcols = [k for k in a[key] if a[key][k].dtype.kind in "iu"]
            parity_ok = parity_ok and all(same[k] for k in intcols)
            if not all(same.values()):               # libm functions (powf, sinf ...) differ between libdevice and glibc
                worst = 0.0
                for k, v in same.items():
                    if not v and a[key][k].dtype.kind == "f" and a[key][k].shape == b[key][k].shape:
                        bb = b[key][k].astype(np.float64)
                        den = max(float(np.nanmax(np.abs(bb))), 1e-30)      # column-wise scale: error relative to the largest value
                        worst = max(worst, float(np.nanmax(np.abs(a[key][k].astype(np.float64) - bb)) / den))
                parity["lists"]["%s/%s" % key]["max_float_error_over_column_max"] = worst
                parity["float_rtol"] = 1e-3
                parity_ok = parity_ok and worst <= 1e-3
        parity["seeds_equal"] = bool(np.array_equal(g.seeds(), orc.seeds()))
        parity_ok = bool(parity_ok and parity["seeds_equal"])
        g.close()
        orc.close()
    line = dict(base, value=value, steps=args.steps, warmup=args.warmup, ms_per_step=total_ms / args.steps, clocks=clocks, e2e=e2e,
                gpu_launches=int(launches), roofline=roofline, parity_checked=parity_ok, parity=parity,
                population={"start": n0, "end": n1})
    line["config"]["l2"] = "flushed between timed iterations (256 MiB scratch rewrite, untimed)"
    if not args.no_cpu_baseline:
        v, dt, k = cpu_run(1 if w.n > (1 << 22) else 2 if w.n > 100000 else w.iterations, 1)
        line["cpu_baseline"] = {"value": v, "unit": "agent-steps/s", "cores": cores, "kind": "port",
                                "sample": f"{k} full iteration(s) of the workload ({dt:.1f} s), OpenMP {cores} threads"}
    print(json.dumps(line))
    sim.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c2_16m",
                    choices=sorted(k for k, c in W.BROWNIAN.items() if c.world == 1) + sorted(W.MODEL_WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-cuda", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--option", action="append", default=[], help="runtime option key=value")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.workload in W.MODEL_WORKLOADS:
        run_model_workload(args)
        return
    cfg = W.BROWNIAN[args.workload]

    if args.impl == "reference":
        run_reference(args, cfg)
        return

    rank, world, local = dist_env()
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod

    from ev_diffusion_b200 import slabs
    from ev_diffusion_b200.runtime import Simulation, nccl_unique_id
    cfg1 = cfg
    if world > 1:
        key = slab_key(args.workload, world)
        if key not in W.BROWNIAN:
            raise SystemExit(f"no slab variant {key} of workload {args.workload}")
        cfg = W.BROWNIAN[key]
    scaling = scaling_of(args.workload)
    plugin = B.model_path(cfg.name)
    if not os.path.exists(plugin):
        raise SystemExit(f"{plugin} missing: run `python -c 'import __graft_entry__ as g; g.build()'` first")
    sim = Simulation(plugin, device=local)
    for kv in args.option:
        k, v = kv.split("=")
        sim.set_option(k, int(v))
    state = W.brownian_state(cfg)
    transport = None
    if world > 1:
        import torch
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(uid, 0)
        sim.slab_init(rank, world, uid.cpu().numpy().tobytes())
        m = sim.model.messages[0]
        plan = slabs.make_plan(m.dims, m.min_bounds, m.max_bounds, m.is_2d, rank, world)
        state = slabs.partition_state(state, plan)
        transport = ("peer-memory stores over NVLink (CUDA IPC; k_push_planes / k_push) with release/acquire flags; NCCL only "
                     "exchanges the IPC handles" if sim.slab_info().get("p2p") else "NCCL grouped send/recv of fixed-capacity buffers")
    sim.set_agents("Particle", "default", state)

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()
        sim.sync()

    # ---- device-resident metric ------------------------------------------------
    sim.step(args.warmup)
    barrier()
    launches0 = sim.launch_count
    sampler = ClockSampler(local)
    if os.environ.get("BENCH_NO_SAMPLER"):
        sampler.ok = False
    sampler.start()
    ms_each = sim.step_timed(args.steps, FLUSH_BYTES)
    barrier()
    clocks = sampler.stop()
    launches = sim.launch_count - launches0
    total_ms = float(ms_each.sum())
    if os.environ.get("BENCH_DEBUG_SPANS"):
        q = np.percentile(ms_each, [0, 10, 50, 90, 99, 100])
        print(f"rank {rank}: spans ms min/p10/p50/p90/p99/max = " + " ".join("%.3f" % v for v in q) + f" mean {ms_each.mean():.3f}", file=sys.stderr, flush=True)
    if dist is not None:
        import torch
        t = torch.tensor([total_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    n_global = cfg.n                                   # cfg.n is the whole-job population at every N
    value = n_global * args.steps / (total_ms * 1e-3)

    # warm-L2 variant: K back-to-back iterations (graph replays), one event pair
    sim.step(args.steps)
    warm_ms = sim.last_step_ms
    if dist is not None:
        import torch
        t = torch.tensor([warm_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        warm_ms = float(t.item())

    # ---- per-stage times (stream mode, one cudaEvent pair per stage) ------------
    reps = min(args.steps, 20)
    sim.profile_iterations(reps)                      # warms the event pool
    stage = dict(sim.profile_iterations(reps))        # steady state, no host sync between iterations
    classes = {"func:output_location": stage.get("func:output_location", 0.0),
               "build:location": stage.get("sort:location", 0.0) + stage.get("pbm:location", 0.0),  # single-GPU stages
               "func:count_neighbours": stage.get("func:count_neighbours", 0.0),
               "func:move": stage.get("func:move", 0.0)}
    dom = max(classes, key=lambda k: classes[k])
    peak, peak_src = peaks()
    dom_ms = classes[dom]
    n_kernel = sim.agent_count("Particle", "default")   # units one launch of the dominant kernel processes
    achieved = B_ALG[dom] * n_kernel / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    rec = load_ncu_record(args.workload, dom) if world == 1 else None
    traffic = (rec["dram_read_bytes"] + rec["dram_write_bytes"]) if rec else None
    issue = None
    if rec and "warp_instructions" in rec and dom_ms > 0:
        sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
        slots = 4 * 148 * sm_hz                          # warp instructions per second a B200 can issue
        issue = {"bound": "issue slots (4 warp instructions / clock / SM x 148 SMs)", "warp_instructions": rec["warp_instructions"],
                 "achieved": rec["warp_instructions"] / (dom_ms * 1e-3) / 1e9, "peak": slots / 1e9, "unit": "G warp-instr/s",
                 "frac": rec["warp_instructions"] / (dom_ms * 1e-3) / slots,
                 "ncu": {k: rec[k] for k in ("issue_active_frac", "active_lanes_per_instruction", "l1_sector_hit_rate",
                                             "l2_sector_hit_rate", "source", "commit") if k in rec}}
    l1_pipe = None
    if rec and "l1_wavefronts_per_sm" in rec and dom_ms > 0:
        sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
        # every LDG.128 of a message record delivers 32 lanes x 16 bytes to the register file = 4 wavefronts of the
        # SM's 128-byte/clock L1 data pipe, whatever the hit rate: the second bound of the neighbourhood walk
        l1_pipe = {"bound": "L1 data pipe (1 wavefront of 128 bytes / clock / SM)", "wavefronts_per_sm": rec["l1_wavefronts_per_sm"],
                   "achieved": rec["l1_wavefronts_per_sm"] / (dom_ms * 1e-3) / 1e9, "peak": sm_hz / 1e9, "unit": "G wavefronts/s/SM",
                   "frac": rec["l1_wavefronts_per_sm"] / (dom_ms * 1e-3) / sm_hz,
                   "ncu": {k: rec[k] for k in ("l1_data_pipe_frac", "lsu_writeback_frac", "global_load_requests", "source", "commit") if k in rec}}
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "alg_bytes_per_agent": B_ALG[dom], "kernel_ms": dom_ms, "agents_per_launch": n_kernel,
                # the bound that actually limits this kernel (user script over ~115 messages per agent): issue slots
                "issue": issue,
                "l1_data_pipe": l1_pipe,
                "stage_ms": {k: round(v, 5) for k, v in stage.items()},
                "iteration": {"alg_bytes_per_agent_step": B_ALG_ITERATION,
                              "achieved_per_gpu": B_ALG_ITERATION * n_global / world * args.steps / (total_ms * 1e-3) / 1e9,
                              "frac": B_ALG_ITERATION * n_global / world * args.steps / (total_ms * 1e-3) / 1e9 / peak}}

    # ---- end to end through the public API with host buffers --------------------
    e2e_steps = min(max(args.steps, 12), 24)
    import torch
    pin = {k: torch.from_numpy(v.copy()).pin_memory().numpy() for k, v in state.items()}
    out = {k: torch.empty(len(state[k]) + 65536, dtype=torch.from_numpy(state[k]).dtype).pin_memory().numpy() for k in ("x", "y", "z", "nbr")}
    if world > 1:     # the step's inputs are id, x, y, z; nbr is an output the iteration overwrites (filled with its default on the device)
        pin = {k: v for k, v in pin.items() if k != "nbr"}
    h2d = sum(v.nbytes for v in pin.values())
    d2h = 0
    for _ in range(2 if world == 1 else 4):   # untimed: first-touch of the pinned pages, graph capture (slab mode replays up to 4 graphs)
        sim.set_agents("Particle", "default", pin)
        sim.step(1)
        sim.read_agents("Particle", "default", out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        sim.set_agents("Particle", "default", pin)                 # H2D: the whole population (N = 1: 5 columns; N > 1: id, x, y, z)
        sim.step(1)                                                # one iteration
        got = sim.read_agents("Particle", "default", out)          # D2H: x, y, z, nbr
    barrier()
    e2e_dt = time.perf_counter() - t0
    d2h = sum(got * o.itemsize for o in out.values())
    if dist is not None:
        t = torch.tensor([e2e_dt], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
        b = torch.tensor([h2d, d2h], dtype=torch.int64, device="cuda")   # whole-job bytes: every rank moves its own slab
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
        h2d, d2h = int(b[0].item()), int(b[1].item())
    e2e_serial = {"value": n_global * e2e_steps / e2e_dt, "unit": "agent-steps/s", "h2d_bytes_per_step": h2d,
                  "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                  "mode": "serial: set_agents (H2D) -> step -> read_agents (D2H), one simulation, each call synchronous"}
    e2e = e2e_serial
    if world == 1 and not os.environ.get("BENCH_NO_PIPELINE"):
        # the same per-step work -- upload of the step's population from pinned host memory, one iteration, read-back of
        # x, y, z, nbr -- with three simulations taking turns through the asynchronous calls of the C ABI, so that one's
        # upload, another's iteration and the third's read-back run at the same time
        depth = 3
        extra = [Simulation(plugin, device=local) for _ in range(depth - 1)]
        for s_ in extra:
            for kv in args.option:
                k, v = kv.split("=")
                s_.set_option(k, int(v))
        n_agents = len(state["id"])
        ring = [(sim, out)] + [(s_, {k: torch.empty_like(torch.from_numpy(v)).pin_memory().numpy() for k, v in out.items()}) for s_ in extra]
        # the step's inputs are the variables the model reads (id, x, y, z); nbr is an output the iteration overwrites, so
        # it is left to its XMML default (filled on the device) instead of being sent over PCIe
        up = {k: v for k, v in pin.items() if k != "nbr"}
        h2d_p = sum(v.nbytes for v in up.values())
        for s_, o_ in ring:                                          # untimed: graph capture, first touch
            s_.set_agents("Particle", "default", up)
            s_.step(1)
            s_.read_agents("Particle", "default", o_)
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            s_, o_ = ring[k % depth]
            s_.sync()                                                # its previous step's results have landed in o_
            s_.set_agents_async("Particle", "default", up)           # H2D
            s_.step_async(1)                                         # one iteration
            s_.read_agents_async("Particle", "default", o_, n_agents)   # D2H
        for s_, _ in ring:
            s_.sync()
        dt2 = time.perf_counter() - t0
        for s_ in extra:
            s_.close()
        e2e = {"value": n_global * e2e_steps / dt2, "unit": "agent-steps/s", "h2d_bytes_per_step": h2d_p, "d2h_bytes_per_step": d2h,
               "steps": e2e_steps,
               "mode": "pipelined: %d simulations take turns, fgpu_sim_set_agents_async / step_async / read_agents_async per step "
                       "(every step uploads its whole input population -- id, x, y, z -- from pinned host memory and reads x, y, z, nbr "
                       "back; both PCIe directions and the SMs are busy at once)" % depth,
               "serial": e2e_serial}
    # ---- parity (untimed) -----------------------------------------------------------
    parity_ok, parity = None, None
    if not args.no_parity:
        try:
            if world == 1:
                sim.close()
                sim = None
                parity_ok, parity = parity_single_gpu(cfg, plugin, local, os.cpu_count() or 1)
            else:
                sim.set_agents("Particle", "default", state)
                parity_ok, parity = parity_slabs(cfg, sim, dist, torch.device("cuda", local), world)
        except Exception as e:                                   # reported, never hidden
            parity_ok, parity = False, {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}

    planes = cfg.dims[2] // world
    line = {
        "metric": METRIC, "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, cfg1), "agents": n_global, "agents_per_gpu": n_global // world,
                   "cells": list(cfg.dims),
                   "parallelism": "single GPU" if world == 1 else
                   f"{world} z-slabs of {planes} cell planes ({scaling} scaling); halo planes + migrating agents by {transport}",
                   "l2": "flushed between timed iterations (256 MiB scratch rewrite, untimed)" +
                         ("; ranks re-aligned by an untimed neighbour flag exchange after each flush" if world > 1 else ""),
                   "options": args.option},
        "ms_per_step_warm_l2": warm_ms / args.steps,
        "value_warm_l2": n_global * args.steps / (warm_ms * 1e-3),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
        "parity_checked": parity_ok, "parity": parity,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(cfg)
    if rank == 0 and world == 1 and not args.no_reference_cuda:
        xsim = Simulation(plugin, device=local)
        line["reference_cuda"] = reference_cuda(cfg, xsim)       # its own process; this one idles meanwhile
        xsim.close()
        if "value" in line["reference_cuda"]:
            line["reference_cuda"]["ours_over_reference_cuda"] = value / line["reference_cuda"]["value"]
    if rank == 0:
        if parity_ok is False:
            print("bench.py: PARITY CHECK FAILED: %s" % json.dumps(parity), file=sys.stderr, flush=True)
        print(json.dumps(line))
    if sim is not None:
        sim.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
