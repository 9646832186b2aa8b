#!/usr/bin/env python
"""bench.py — TRAPPIST-1 ensemble throughput of the B200 WHFast path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

One bench "step" = one pass of the hot path over the batch: `--steps-per-call` WHFast steps of every system of the
ensemble (cases/trappist1.py, 8 bodies, tides + flattening + GR Kidder1995, democratic-heliocentric WHFast, dt = 0.08 d)
in ONE kernel launch, state resident in registers. Workload per GPU is fixed (65536 systems): weak scaling, no
collective on the hot path; one tiny NCCL all-reduce of the per-rank status counts after the timed region.

Prints ONE JSON line (rank 0). Keys follow the driver contract:
  value     system-steps/s, all GPUs, inputs resident in HBM, timed with CUDA events on the launching stream, max over ranks
  e2e       the same metric through the C-ABI call with HOST (pinned) buffers: H2D state upload + steps + D2H download
  roofline  FP64 vector pipe: algorithmic flops (exact count from the oracle's counting build) / kernel time / peak
  cpu_baseline  the CPU restatement of the reference (oracle/, validated bit-exact against the reference's goldens)
                on the box's host cores, bounded sample — the reference Rust binary cannot be built in this image.
"""
import argparse
import gzip
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOAD = "c4_trappist1"
CONFIG_INDEX = 4
# exact operation count of one system-step of this workload (oracle counting build, every + - * / sqrt = 1 flop,
# SURVEY.md §8d counting rule; regenerate with oracle.binding.count_flops) — see DESIGN.md
FLOPS_PER_SYSTEM_STEP = 38471.0
FP64_THEORETICAL_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12  # 37.2


def load_case():
    from posidonius_b200.case import case_from_dict
    with gzip.open(os.path.join(ROOT, "tests", "golden", "configs", WORKLOAD + ".json.gz"), "rt") as f:
        return case_from_dict(json.load(f))


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons of one GPU during the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                      "--format=csv,noheader,nounits"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm, mx, reasons = [], 0.0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_baseline(case, tables, steps_per_system=1000, target_seconds=12.0):
    """Times the CPU restatement (oracle) on all host cores over a bounded sample of the same workload."""
    from oracle.binding import run_ensemble
    from posidonius_b200.perturb import make_ensemble_cases
    cores = os.cpu_count() or 1
    # calibrate: ~12 us per system-step per core for this workload
    cal_sys = 4 * cores
    cases = make_ensemble_cases(case, cal_sys, 20261017 + CONFIG_INDEX)
    _, _, secs = run_ensemble(cases, cal_sys, tables, 200, True, cores)
    rate = cal_sys * 200 / max(secs, 1e-6)
    n_sys = int(max(cores, min(65536, rate * target_seconds / steps_per_system)))
    n_sys = (n_sys // cores) * cores or cores
    cases = make_ensemble_cases(case, n_sys, 20261017 + CONFIG_INDEX)
    _, status, secs = run_ensemble(cases, n_sys, tables, steps_per_system, True, cores)
    import shutil
    rust = [x for x in ("cargo", "rustc", "posidonius") if shutil.which(x)]
    return {"value": n_sys * steps_per_system / secs, "unit": "system-steps/s", "cores": cores, "kind": "port",
            "sample": "%d perturbed TRAPPIST-1 systems x %d steps on %d host threads (%.1f s); CPU restatement of the "
                      "reference (oracle/, bit-exact vs the reference goldens), not the Rust binary (%s)"
                      % (n_sys, steps_per_system, cores, secs,
                         "found on this box but not used: " + ", ".join(rust) if rust else "no cargo / rustc / posidonius binary on this box")}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port) on the host cores, same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.binding import run_ensemble
    from posidonius_b200.perturb import make_ensemble_cases
    case, tables = load_case()
    cores = os.cpu_count() or 1
    n_sys = 128 * cores   # about a second of work per step for all host threads
    spc = 1000
    cases = make_ensemble_cases(case, n_sys, 20261017 + CONFIG_INDEX)
    for _ in range(args.warmup):
        run_ensemble(cases, n_sys, tables, 100, True, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run_ensemble(cases, n_sys, tables, spc, True, cores)
    dt = time.perf_counter() - t0
    value = n_sys * spc * args.steps / dt
    line = {
        "impl": "reference", "metric": "TRAPPIST-1 ensemble system-steps/s", "value": value, "unit": "system-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "systems_per_step": n_sys, "whfast_steps_per_step": spc, "bodies": case.n_particles,
                   "note": "CPU restatement of the reference algorithm on all host threads; the Rust binary cannot be built here (no cargo/rustc)"},
        "cpu_baseline": {"value": value, "unit": "system-steps/s", "cores": cores, "kind": "port",
                         "sample": "%d systems x %d steps x %d repeats" % (n_sys, spc, args.steps)},
        "e2e": {"value": value, "unit": "system-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--systems", type=int, default=65536, help="systems per GPU")
    ap.add_argument("--steps-per-call", type=int, default=2000,
                    help="WHFast steps per launch (one bench step); the default times 5 x 2000 = 10^4 steps (SURVEY §8d horizon)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--arithmetic", default="hybrid", choices=["hybrid", "fast", "strict"],
                    help="hybrid (default): fast midpoint iterates, exact committed evaluation; strict: every evaluation exact; fast: none")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # The contract is ONE JSON line on stdout. Libraries write there too (NCCL prints its version banner on fd 1 when
    # NCCL_DEBUG is set): send fd 1 to stderr for the run and keep the real stdout for the result line.
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from posidonius_b200.ensemble import Ensemble, measure_fp64_peak
    from posidonius_b200.perturb import make_ensemble_cases
    case, tables = load_case()
    n_sys, spc = args.systems, args.steps_per_call
    # keep the whole run inside the case's time limit
    cases = make_ensemble_cases(case, n_sys, 20261017 + CONFIG_INDEX + 1000 * rank)
    ens = Ensemble(cases, tables, device=local, arithmetic={"fast": 0, "strict": 1, "hybrid": 2}[args.arithmetic])
    ens.initialize_physical_values()
    ens.synchronize()
    # pinned host buffers of the boundary call
    e2e_fields = ("position", "velocity", "acceleration", "angular_momentum", "spin", "velocity_errors", "angular_momentum_errors",
                  "radius", "radius_of_gyration_2", "moment_of_inertia", "current_time")
    host = ens.make_state_buffers(e2e_fields, pinned=True)
    ens.download(out=host)
    io_bytes = int(sum(a.nbytes for a in host.values()))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    e_start, l_start = ens.summary()   # Universe::compute_total_energy / angular momentum of every member (untimed)
    for _ in range(args.warmup):
        ens.iterate(spc, synchronize=True)
    fp64_peak = measure_fp64_peak(local, 30.0) if rank == 0 else 0.0

    # ---- device-resident throughput
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ens.launch_count()
    barrier()
    t0 = time.perf_counter()
    kernel_ms = 0.0
    for _ in range(args.steps):
        ens.iterate(spc, synchronize=False)
        kernel_ms += ens.last_step_ms()   # CUDA events on the launching stream (synchronizes on the end event)
    barrier()
    wall = time.perf_counter() - t0
    launches = ens.launch_count() - launches0
    clocks = sampler.summary()
    st, warn, _ = ens.status()
    alive = int(np.sum(st == 0))

    # ---- end to end through the boundary with host buffers
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        ens.run_host(host, spc)
    barrier()
    e2e_wall = time.perf_counter() - t1

    # ---- after the timed regions: per-system summaries {status, t, dE/E, dL/L}, gathered on rank 0 (the one gather of
    # ensemble data; NCCL when there are several ranks)
    from posidonius_b200.shard import gather_summaries
    e_end, l_end = ens.summary()
    t_end = ens.get_current_time()
    rows = torch.from_numpy(np.stack([st.astype(np.float64), t_end, (e_end - e_start) / np.abs(e_start),
                                      (l_end - l_start) / np.abs(l_start)], axis=1)).to("cuda")
    rows = gather_summaries(rows, dist if world > 1 else None)

    from posidonius_b200.shard import reduce_timing
    elapsed = torch.tensor([kernel_ms * 1e-3, wall, e2e_wall], dtype=torch.float64, device="cuda")
    counts = torch.tensor([alive, n_sys], dtype=torch.int64, device="cuda")
    # the only collective: max of the timed regions, per-rank summaries — after the timed region
    elapsed, counts = reduce_timing(elapsed, counts, dist if world > 1 else None)
    kern_s, wall_s, e2e_s = [float(x) for x in elapsed.tolist()]
    total_sys = world * n_sys
    units = total_sys * spc * args.steps

    if rank == 0:
        value = units / kern_s
        achieved = n_sys * spc * args.steps * FLOPS_PER_SYSTEM_STEP / (kernel_ms * 1e-3) / 1e12  # this GPU's dominant kernel
        line = {
            "metric": "TRAPPIST-1 ensemble system-steps/s", "value": value, "unit": "system-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * kern_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "systems_per_gpu": n_sys, "bodies": case.n_particles, "whfast_steps_per_step": spc,
                       "effects": "tides(CTL)+rotational_flattening(oblate)+GR(Kidder1995)", "coordinates": "DemocraticHeliocentric",
                       "time_step_days": case.time_step, "arithmetic": args.arithmetic, "parallelism": "ensemble-sharded x%d, no collective" % world,
                       "l2": "state (%.0f MB/GPU) larger than L2; registers hold it between launch start and end" % (io_bytes / 1e6),
                       "wall_clock_value": units / wall_s, "systems_alive": int(counts[0]), "systems_total": int(counts[1]),
                       "ensemble_summary": {"gathered_systems": int(rows.shape[0]), "max_abs_dE_over_E": float(rows[:, 2].abs().max()),
                                            "max_abs_dL_over_L": float(rows[:, 3].abs().max()), "t_days_min": float(rows[:, 1].min()),
                                            "t_days_max": float(rows[:, 1].max())}},
            "clocks": clocks,
            "e2e": {"value": units / e2e_s, "unit": "system-steps/s", "h2d_bytes_per_step": io_bytes, "d2h_bytes_per_step": io_bytes},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
                         "frac": achieved / (fp64_peak / 1e12) if fp64_peak else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one bench launch (65536 systems, its steps cut into
                         # 4 time slices: the state crosses HBM once per slice), ncu, profiles/r1_traffic_bench_launch.csv
                         "traffic": 1230.1e6 * n_sys / 65536.0,
                         "peak_source": "measured here: DFMA-chain microbenchmark (pb200_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 figure",
                         "frac_of_theoretical_37.2": achieved / FP64_THEORETICAL_TFLOPS,
                         "flops_per_system_step": FLOPS_PER_SYSTEM_STEP},
        }
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(case, tables)
        sys.stdout.flush()
        os.write(result_fd, (json.dumps(line) + "\n").encode())
    ens.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
