#!/usr/bin/env python
"""bench.py — particle-steps/s of the advance() hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" = one advance() (P2G + grid update + G2P, plus the per-step re-binning) over the whole
synthetic scene.  `value` is device-timed with inputs resident in HBM; `e2e` goes through the C-ABI with
HOST buffers: every step uploads the particle state from pinned host memory, advances once and downloads
the new state (the reference's per-step snapshot pattern, src/solver.cpp:50-59).

Workloads (BASELINE.json configs; scenes from BASELINE.md §4):
  cfg4      3D snow, cube<3>(256,0.25,0.5) = 16 777 216 particles, 512^3 grid   [default, all N; strong scaling]
  snow128   3D snow, cube<3>(154,0.2,0.8)  =  3 652 264 particles, 128^3 grid   (the 1e9 single-GPU target scene)
  cfg2      3D jelly, cube<3>(64,0.375,0.625) = 262 144 particles, 128^3 grid
  cfg3      3D liquid, 126^3 block = 2 000 376 particles, 256^3 grid
  cfg1      2D snow, two 50x50 squares = 5 000 particles, 64^2 grid (README scene)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SNOW, JELLY, LIQUID = 0, 1, 2


def scene(name: str):
    """-> (positions (n,dim) float32, model, res, description)"""
    import nuclearmpm_b200 as nm
    if name == "cfg4":
        return nm.cube(3, 256, 0.25, 0.5), SNOW, 512, "cfg4: 3D snow cube<3>(256,0.25,0.5) 16777216 p, 512^3 grid"
    if name == "snow128":
        return nm.cube(3, 154, 0.2, 0.8), SNOW, 128, "3D snow cube<3>(154,0.2,0.8) 3652264 p, 128^3 grid"
    if name == "cfg2":
        return nm.cube(3, 64, 0.375, 0.625), JELLY, 128, "cfg2: 3D jelly cube<3>(64,0.375,0.625) 262144 p, 128^3 grid"
    if name == "cfg3":
        return (nm.cube(3, 126, 0.05, 0.05 + 62.5 / 256), LIQUID, 256,
                "cfg3: 3D liquid 126^3 block 2000376 p, 256^3 grid")
    if name == "cfg1":
        a = nm.cube(2, 50, 0.4, 0.6)
        b = a - np.array([0, 0.35], np.float32)
        return np.concatenate([a, b]).astype(np.float32), SNOW, 64, "cfg1: 2D snow two 50x50 squares 5000 p, 64^2 grid"
    raise SystemExit(f"unknown workload {name}")


# algorithmic bytes per particle-step: P2G reads S+2 floats, G2P reads x,F,Jp and writes S
ALGO_BYTES = {3: dict(p2g=108, g2p=152), 2: dict(p2g=60, g2p=80)}
# Fused G2P+P2G (nmpm_fused.cuh): ONE launch does the G2P of step n and the P2G of step n+1.  Its algorithmic bytes by
# SURVEY.md §8(d)'s per-unit figures are the sum (260 B: what the two phases move when they are separate passes);
# the bytes the fused pass itself has to move are 160 B per particle (read x,F,Jp,mass,volume = 60, write the state = 100).
FUSED_MIN_BYTES = {3: 160}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.proc, self.lines = device, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self) -> dict:
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel: str, n_particles: float):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture (profiles/ncu_traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum per particle on the 3D snow scenes), scaled to this launch."""
    p = ROOT / "profiles" / "ncu_traffic.json"
    try:
        per_particle = json.loads(p.read_text())["bytes_per_particle"][kernel]
        return float(per_particle) * float(n_particles)
    except Exception:
        return None


# ----------------------------------------------------------------------------------------------
# CPU baseline: the reference's own advance() (oracle/_ref, -Ofast build) or the oracle port
# ----------------------------------------------------------------------------------------------
def cpu_sample(workload: str):
    """What the CPU arm times for `workload`.

    Every workload except cfg4 is run in full.  cfg4 (16.8 M particles on 513^3 nodes) costs the single-threaded
    reference ~25 s per step, so K steps of it do not fit a bench run; its sample is the SAME scene at half the
    linear resolution — cube<3>(128, 0.25, 0.5) on a 256^3 grid: same material, dt, particles per cell (8) and
    grid nodes per particle (8.1 vs 8.05), i.e. the same mix of per-particle work (p2g/g2p + SVD) and per-node work
    (the reference re-allocates and walks the dense grid every step, src/nclr.h:105-109,263-282).  A sub-block on
    the full 513^3 grid (round 1) over-weights the per-node work and under-states the reference several times.
    `--impl reference` additionally times real full-scene steps (`full_scene`)."""
    import nuclearmpm_b200 as nm
    x, model, res, desc = scene(workload)
    dim = x.shape[1]
    if workload == "cfg4":
        xs = nm.cube(3, 128, 0.25, 0.5)
        return xs, model, 256, dim, ("half-scale replica of cfg4: cube<3>(128,0.25,0.5) = 2097152 p on a 256^3 grid "
                                     "(same material/dt, 8 particles per cell, 8.1 grid nodes per particle as cfg4's 8.05)")
    return x, model, res, dim, "the full scene"


def host_cpu():
    model = "unknown"
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                model = ln.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return {"model": model, "cores_total": os.cpu_count()}


def cpu_reference_rate(workload: str, steps: int, warmup: int):
    from oracle import cpu_oracle as co
    kind = "ref_fast" if co.available("ref_fast") else "port"
    xs, model, res, dim, what = cpu_sample(workload)
    sim = co.CpuSim(xs, model, res, kind=kind)
    if warmup:
        sim.time_advance(warmup)
    t = sim.time_advance(steps)
    rate = len(xs) * steps / t
    return dict(value=rate, unit="particle-steps/s", cores=1, kind="reference" if kind == "ref_fast" else "port",
                host=host_cpu(), sample_particles=int(len(xs)), sample_grid_res=int(res),
                sample=f"{what}; {steps} advance() steps (after {warmup} warm-up) in {t:.2f} s on 1 host core "
                       f"({'reference nclr.h, -Ofast, Eigen stand-in' if kind == 'ref_fast' else 'oracle port, -O2'}; "
                       "the reference is single-threaded: its omp pragmas are inert, CMakeLists.txt:9-10)"), t


def cpu_full_scene(workload: str, steps: int = 2):
    """Real full-scene reference steps (cfg4: ~4 GB of host memory, tens of seconds per step)."""
    from oracle import cpu_oracle as co
    kind = "ref_fast" if co.available("ref_fast") else "port"
    x, model, res, desc = scene(workload)
    sim = co.CpuSim(x, model, res, kind=kind)
    per = [sim.time_advance(1) for _ in range(steps)]
    return {"steps": steps, "seconds_per_step": per, "value": len(x) * steps / sum(per), "unit": "particle-steps/s",
            "what": f"{desc}: the first {steps} advance() steps of the FULL scene on 1 host core"}


def bench_config(args, desc, n_total, res, dim, model):
    """`config` of the JSON line — the same dict for the GPU arm and for `--impl reference` (same workload)."""
    return {"workload": desc, "particles": int(n_total), "grid_res": int(res), "dim": int(dim),
            "material": ["snow", "jelly", "liquid"][model],
            "l2_policy": "inputs larger than L2 (particle store + grid >> 126 MB)" if n_total * 116 > 2e8
            else "working set smaller than L2 (scene is small); no flush"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    x, model, res, desc = scene(args.workload)
    cb, t = cpu_reference_rate(args.workload, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "particle-steps/s", "value": cb["value"], "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args, desc, len(x), res, x.shape[1], model),
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if cb["sample_particles"] != len(x) and args.gpus == 1 and not args.no_full_check:
        line["full_scene"] = cpu_full_scene(args.workload, 2)   # cross-check of the replica's rate on the real scene
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import nuclearmpm_b200 as nm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL writes its version banner / debug lines to stdout by default: keep stdout for the ONE JSON line.
        # (NCCL_DEBUG_FILE is honoured only above the VERSION level, so VERSION is raised to WARN.)
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    x, model, res, desc = scene(args.workload)
    dim = x.shape[1]
    n_total = len(x)

    if world > 1:
        from nuclearmpm_b200 import slab
        result = slab.bench_slabs(args, x, model, res, desc, rank, world, local)
        if rank == 0:
            print(json.dumps(result))
        dist.destroy_process_group()
        return

    stream = torch.cuda.Stream()
    sim = nm.MPMSimulation(x, model, res, device=local, sort_every=args.sort_every, p2g_variant=args.p2g_variant,
                           fuse=args.fuse)
    sim.set_stream(stream.cuda_stream)
    fused = bool(sim.fused)

    def sync():
        sim.synchronize()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------------
    # advance() replays one CUDA graph per host-side step state (2 * sort_every of them): run through the cycle once
    # so that no capture / instantiation lands in the warm-up or in the timed region
    # (and the 8-step cycle graph nmpm_advance uses for runs of steps: 4 cycles in one call capture it)
    # advance() replays CUDA graphs: one per host-side step state for single steps (8 states at the default cadence) and one
    # for a whole 8-step cycle.  Capture them all before anything is timed: W warm-up steps, single steps up to the next
    # cycle boundary, two whole cycles (capture + replay), then one cycle of single steps.
    sim.advance(args.warmup)
    cyc = step = 2 * args.sort_every if args.sort_every > 0 else 1   # nmpm_api.cu: graph_cycle()
    while cyc % 4:
        cyc += step
    prime = 0
    while (args.warmup + prime) % cyc:
        sim.advance(1)
        prime += 1
    sim.advance(cyc), sim.advance(cyc)
    for _ in range(cyc):
        sim.advance(1)
    prime += 3 * cyc
    sync()
    peak, peak_src = measured_peak_gbs()
    ab = ALGO_BYTES[dim]
    steps_done = prime + args.warmup

    def timed_window(clk_device=None):
        """K steps between CUDA events on the sim's stream, then a separate per-phase pass (events per phase
        serialise the step, so it is not part of the timed region)."""
        nonlocal steps_done
        first = steps_done
        l0 = sim.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync()
        with torch.cuda.stream(stream):
            e0.record(stream)
            sim.advance(args.steps)
            e1.record(stream)
        sync()
        ms = e0.elapsed_time(e1)
        launches = sim.launch_count() - l0
        sim.timing_enable(True)
        sim.timing_read(reset=True)
        prof_steps = min(args.steps, 20)
        sim.advance(prof_steps)
        sync()
        tm = sim.timing_read(reset=True)
        sim.timing_enable(False)
        steps_done += args.steps + prof_steps
        phases = {k + "_ms": tm[k] / max(1, tm["steps"]) for k in ("sort", "p2g", "grid", "g2p")}
        if fused:
            # g2p_ms is the fused launch (G2P of step n + P2G of step n+1); p2g_ms is what is left of the P2G phase
            # (buffer swap + clear of the other grid)
            t = phases["g2p_ms"] * 1e-3
            kern = {"g2p_p2g": {"ms": phases["g2p_ms"],
                                "achieved_GBps": (ab["p2g"] + ab["g2p"]) * n_total / t / 1e9 if t > 0 else 0.0,
                                "min_traffic_GBps": FUSED_MIN_BYTES[dim] * n_total / t / 1e9 if t > 0 else 0.0}}
        else:
            kern = {k: {"ms": phases[k + "_ms"],
                        "achieved_GBps": ab[k] * n_total / (phases[k + "_ms"] * 1e-3) / 1e9 if phases[k + "_ms"] > 0 else 0.0}
                    for k in ("p2g", "g2p")}
        for k in kern:
            kern[k]["frac"] = kern[k]["achieved_GBps"] / peak
        return {"steps": [first, first + args.steps], "value": n_total * args.steps / (ms * 1e-3),
                "ms_per_step": ms / args.steps, "phase_ms": phases, "kernels": kern, "launches": int(launches)}

    with ClockSampler(local) as clk:
        early = timed_window()
    clocks = clk.summary()
    value, ms, launches = early["value"], early["ms_per_step"] * args.steps, early["launches"]
    phases = early["phase_ms"]
    dom = "g2p_p2g" if fused else ("g2p" if phases["g2p_ms"] >= phases["p2g_ms"] else "p2g")
    achieved = early["kernels"][dom]["achieved_GBps"]
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "frac_of_nominal_8TBs": achieved / 8000.0,
                "traffic": ncu_traffic(dom, n_total) if dim == 3 else None, "peak_source": peak_src,
                "algorithmic_bytes_per_particle": (ab["p2g"] + ab["g2p"]) if fused else ab[dom],
                "other": early["kernels"], "phase_ms": phases}
    if fused:
        roofline["what"] = ("one launch = G2P of step n + P2G of step n+1 (nmpm_fused.cuh): algorithmic bytes = SURVEY 8(d)'s "
                            "108 (P2G) + 152 (G2P) per particle; frac_min_traffic counts only the 160 B per particle the fused "
                            "pass itself must move")
        roofline["frac_min_traffic"] = early["kernels"][dom]["min_traffic_GBps"] / peak

    # ---- late phase (under Q1 the snow turns into a particle gas that fills the box) ----
    late = None
    if args.late_step > steps_done:
        sim.advance(args.late_step - steps_done)
        steps_done = args.late_step
        late = timed_window()

    # ---- end to end through the C-ABI with host buffers ---------------------------------------
    # Per step: the particle state comes from pinned host memory (H2D), one advance(), the new state goes back to pinned
    # host memory (D2H) — the reference's per-step snapshot loop (src/solver.cpp:50-59) with a teacher-forced input.
    # The async C-ABI calls only enqueue: copy-in of step k+1 and copy-out of step k-1 overlap step k (two copy
    # streams, double-buffered staging on the device, two host buffers per direction).
    st = sim.particles()
    keys = ("x", "v", "F", "C", "Jp")
    h_in = [{k: torch.from_numpy(v.copy()).pin_memory() for k, v in st.items()} for _ in range(2)]
    h_out = [{k: torch.empty_like(h_in[0][k]).pin_memory() for k in keys} for _ in range(2)]
    np_in = [{k: t.numpy() for k, t in d.items()} for d in h_in]
    np_out = [{k: t.numpy() for k, t in d.items()} for d in h_out]
    e2e_steps = max(4, min(args.steps, 12))
    state_bytes = sum(v.nbytes for v in np_in[0].values())

    def e2e_loop(upload):
        sync()
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            if upload:
                sim.upload_async(*[np_in[k & 1][f] for f in keys])
            sim.advance(1)
            sim.download_async(np_out[k & 1])
        sync()
        return time.perf_counter() - t0

    e2e_loop(True)  # warm-up: staging buffers, streams
    t_e2e = e2e_loop(True)
    # the answer that came back is the step of the uploaded state (checked against a blocking round trip)
    sim.upload(*[np_in[0][f] for f in keys])
    sim.advance(1)
    chk = sim.particles()
    e2e_ok = bool(np.abs(chk["x"] - np_out[(e2e_steps - 1) & 1]["x"]).max() <= 1e-5)
    t_dl = e2e_loop(False)
    e2e = {"value": n_total * e2e_steps / t_e2e, "unit": "particle-steps/s", "h2d_bytes_per_step": state_bytes,
           "d2h_bytes_per_step": state_bytes, "steps": e2e_steps, "result_checked": e2e_ok,
           "what": "per step: nmpm_upload_particles_async (pinned host -> device) + nmpm_advance(1) + "
                   "nmpm_download_particles_async (device -> pinned host, input order); copies overlap the steps",
           "download_only": {"value": n_total * e2e_steps / t_dl, "unit": "particle-steps/s", "h2d_bytes_per_step": 0,
                             "d2h_bytes_per_step": state_bytes,
                             "what": "the reference's --dump pattern: advance(1) + snapshot of particles() per step"}}

    # ---- CPU baseline (rank 0, bounded sample) -------------------------------------------------
    if args.no_cpu:
        cpu = None
    else:
        cpu, _ = cpu_reference_rate(args.workload, 6 if n_total > 1_000_000 else 20, 0)

    line = {
        "metric": "particle-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args, desc, n_total, res, dim, model),
        "run": {"sort_every": args.sort_every, "graph_priming_steps": prime, "p2g_variant": args.p2g_variant,
                "fused_g2p_p2g": fused, "timed_steps": early["steps"]},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        "phases": {"early": early, "late": late},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="cfg4")
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--sort-every", type=int, default=4)
    ap.add_argument("--p2g-variant", type=int, default=0)
    ap.add_argument("--fuse", type=int, default=0, help="nmpm_options.fuse: 0 auto (on in 3D), 1 off, 2 on, 3 always")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-full-check", action="store_true", help="--impl reference: skip the 2 real full-scene steps")
    ap.add_argument("--late-step", type=int, default=-1,
                    help="also time K steps starting at this step (late, dispersed phase); -1 = 300 for the 3D snow "
                         "scenes, off otherwise; 0 = off")
    ap.add_argument("--rebalance-every", type=int, default=50, help="multi-GPU: re-balance slab boundaries every k steps")
    ap.add_argument("--no-verify", action="store_true", help="multi-GPU: skip the 3-step comparison with a single-GPU twin")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "graft" else args.warmup
    if args.late_step < 0:
        args.late_step = 300 if args.workload in ("cfg4", "snow128") else 0
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
