#!/usr/bin/env python
"""bench.py — ensemble throughput of the B200 WHFast path (BASELINE.json metric: TRAPPIST-1 ensemble system-steps/s).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W
    python bench.py --workload c5_circumbinary --history     (other BASELINE configurations; historic snapshots drained)

One bench "step" = one pass of the hot path over the batch: `--steps-per-call` WHFast steps of every system of the
ensemble in ONE kernel launch, state resident in registers / shared memory. The default workload is BASELINE.json
config 4: cases/trappist1.py (8 bodies, tides + flattening + GR Kidder1995, democratic-heliocentric WHFast,
dt = 0.08 d) as a perturbed ensemble of 65536 systems IN TOTAL, sharded over the N GPUs by contiguous ranges
(posidonius_b200.shard.shard_range): STRONG scaling, 65536 / N systems per GPU, no collective on the hot path; one NCCL
all-reduce of the timings / status counts and one gather of the per-system summaries after the timed region. The weak-
scaling figure (65536 systems on every GPU) is measured next to it for N > 1 and reported as config.weak_value.

Prints ONE JSON line (rank 0). Keys follow the driver contract:
  value     system-steps/s, all GPUs, inputs resident in HBM, timed with CUDA events on the launching stream, max over ranks
  e2e       the same metric through the C-ABI call with HOST (pinned) buffers: H2D state upload + steps + D2H download
  roofline  FP64 vector pipe: algorithmic flops (exact count from the oracle's counting build) / kernel time / peak
  cpu_baseline  the CPU restatement of the reference (oracle/, validated bit-exact against the reference's goldens)
                on the box's host cores, bounded sample — the reference Rust binary cannot be built in this image.
  config.other_workloads  (N = 1 only) the other BASELINE configurations at their BASELINE ensemble sizes, measured briefly
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

# BASELINE.json configurations: (config index, BASELINE ensemble size, exact operation count of one system-step).
# The count is the oracle's counting build over the first 100 steps of the unperturbed case (every + - * / sqrt of the
# reference as written = 1 flop, SURVEY.md §8d counting rule; oracle.binding.count_flops; tests/test_bench_constants_cpu.py
# keeps the table honest) — see DESIGN.md §3.
WORKLOADS = {
    "c1_example": (1, 65536, 5981.0),
    "c2_case3": (2, 4096, 2763.0),
    "c3_case7": (3, 16384, 14059.0),
    "c3_case7_evolving": (3, 16384, 11308.0),
    "c4_trappist1": (4, 65536, 38471.0),
    "c5_circumbinary": (5, 65536, 9408.0),
}
DEFAULT_WORKLOAD = "c4_trappist1"
SEED = 20261017
AMPLITUDE = 1e-3
FP64_THEORETICAL_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12  # 37.2
ARITH = {"fast": 0, "strict": 1, "hybrid": 2}


def load_case(workload, history):
    """The workload's case image. Run-length knobs only are touched: the time limit is pushed out of the way of the timed
    steps and, unless --history, so is the historic-snapshot period (the snapshot at t = 0 still falls due)."""
    from posidonius_b200.case import case_from_dict
    with gzip.open(os.path.join(ROOT, "tests", "golden", "configs", workload + ".json.gz"), "rt") as f:
        d = json.load(f)
    d["universe"]["time_limit"] = max(d["universe"]["time_limit"], 1e9 * d["time_step"])
    if not history:
        d["historic_snapshot_period"] = 1e8 * d["time_step"]
    return case_from_dict(d)


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


def host_members(case, n_sys, seed):
    """Members 0 .. n_sys - 1 of the global synthetic ensemble as host case images (the recipe of
    pb200_ensemble_create_perturbed, stated on the host): what the reference arm / the CPU baseline integrate."""
    from posidonius_b200.perturb import splitmix_cases
    return splitmix_cases(case, n_sys, seed, AMPLITUDE)


def cpu_sample(case, tables, seed, steps_per_system, target_seconds, cores):
    """A bounded sample of the ensemble for the CPU restatement (oracle) on all host threads: calibrates on 4 x cores members."""
    from oracle.binding import run_ensemble
    cal_sys = 4 * cores
    cases = host_members(case, cal_sys, seed)
    _, _, secs = run_ensemble(cases, cal_sys, tables, 200, True, cores)
    rate = cal_sys * 200 / max(secs, 1e-6)
    n_sys = int(max(cores, min(65536, rate * target_seconds / steps_per_system)))
    n_sys = (n_sys // cores) * cores or cores
    return n_sys, host_members(case, n_sys, seed)


def rust_note():
    import shutil
    rust = [x for x in ("cargo", "rustc", "posidonius") if shutil.which(x)]
    return "found on this box but not used: " + ", ".join(rust) if rust else "no cargo / rustc / posidonius binary on this box"


def cpu_baseline(workload, case, tables, seed, steps_per_system=1000, target_seconds=12.0):
    from oracle.binding import run_ensemble
    cores = os.cpu_count() or 1
    n_sys, cases = cpu_sample(case, tables, seed, steps_per_system, target_seconds, cores)
    _, status, secs = run_ensemble(cases, n_sys, tables, steps_per_system, True, cores)
    return {"value": n_sys * steps_per_system / secs, "unit": "system-steps/s", "cores": cores, "kind": "port",
            "sample": "%d members of the %s ensemble x %d steps on %d host threads (%.1f s); CPU restatement of the "
                      "reference (oracle/, bit-exact vs the reference goldens), not the Rust binary (%s)"
                      % (n_sys, workload, steps_per_system, cores, secs, rust_note())}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port) on all host threads, same metric / workload / steps per
    bench step; the number of systems per step is a bounded sample of the ensemble (a few seconds of work per step)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.binding import run_ensemble
    workload = args.workload
    cfg_index, n_base, _ = WORKLOADS[workload]
    case, tables = load_case(workload, False)
    cores = os.cpu_count() or 1
    spc = args.steps_per_call
    n_total = args.systems or n_base
    n_sys, cases = cpu_sample(case, tables, SEED + cfg_index, spc, 4.0, cores)
    for _ in range(args.warmup):
        run_ensemble(cases, n_sys, tables, 50, True, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run_ensemble(cases, n_sys, tables, spc, True, cores)
    dt = time.perf_counter() - t0
    value = n_sys * spc * args.steps / dt
    sample = "the first %d of the %d members x %d steps x %d repeats on %d host threads" % (n_sys, n_total, spc, args.steps, cores)
    line = {
        "impl": "reference", "metric": "TRAPPIST-1 ensemble system-steps/s" if workload == DEFAULT_WORKLOAD else workload + " ensemble system-steps/s",
        "value": value, "unit": "system-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "systems_total": n_total, "systems_per_step": n_sys,
                   "whfast_steps_per_step": spc, "bodies": case.n_particles, "host_threads": cores,
                   "note": "CPU restatement of the reference algorithm on all %d host threads; the metric is a per-system-step rate, "
                           "so timing a bounded sample (%d systems per bench step instead of the whole ensemble, to end within "
                           "minutes) does not change it; the Rust binary cannot be built here (%s)" % (cores, n_sys, rust_note())},
        "cpu_baseline": {"value": value, "unit": "system-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "system-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def time_launches(ens, spc, reps):
    """Best-of-reps launch duration (ms) of `spc` steps, CUDA events on the launching stream."""
    ens.iterate(spc, synchronize=True)
    best = 1e30
    for _ in range(reps):
        ens.iterate(spc, synchronize=False)
        best = min(best, ens.last_step_ms())
    return best


def other_workloads(device, fp64_peak, arithmetic, skip):
    """The other BASELINE configurations at their BASELINE ensemble sizes (and at 65536 where that differs): value, fraction
    of the measured FP64 peak, live systems. Short runs (1000 steps per launch, best of 2) after the main measurement."""
    from posidonius_b200.ensemble import Ensemble
    out = {}
    for name, (idx, n_base, flops) in WORKLOADS.items():
        if name == skip:
            continue
        case, tables = load_case(name, False)
        for n_sys in sorted({n_base, 65536}):
            with Ensemble.perturbed(case, tables, n_sys, SEED + idx, AMPLITUDE, device=device, arithmetic=arithmetic) as ens:
                ens.initialize_physical_values()
                ms = time_launches(ens, 1000, 2)
                st, _, _ = ens.status()
                build = ens.last_kernel()
            rate = n_sys * 1000 / (ms * 1e-3)
            out["%s_x%d" % (name, n_sys)] = {
                "value": rate, "bodies": case.n_particles, "systems": n_sys, "flops_per_system_step": flops,
                "tflops": rate * flops / 1e12, "frac": rate * flops / fp64_peak if fp64_peak else None,
                "systems_alive": int((st == 0).sum()), "kernel_build": build}
    return out


def measured_traffic(workload, n_sys, spc, arithmetic):
    """dram__bytes_read.sum + dram__bytes_write.sum of one bench launch from the committed ncu capture of the same
    configuration (profiles/r2_traffic.json, written by scripts/ncu_traffic.py); None when there is no matching capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            for row in json.load(f):
                if (row["workload"], row["systems"], row["steps_per_call"], row["arithmetic"]) == (workload, n_sys, spc, arithmetic):
                    return float(row["dram_bytes"])
    except Exception:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--systems", type=int, default=0, help="systems IN TOTAL over all GPUs (default: the workload's BASELINE ensemble size)")
    ap.add_argument("--steps-per-call", type=int, default=2000,
                    help="WHFast steps per launch (one bench step); the default times 5 x 2000 = 10^4 steps (SURVEY §8d horizon)")
    ap.add_argument("--history", action="store_true",
                    help="keep the case's historic snapshot period and drain the 156-byte records to the host inside the timed regions")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-workloads", action="store_true")
    ap.add_argument("--no-weak", action="store_true")
    ap.add_argument("--arithmetic", default="hybrid", choices=sorted(ARITH),
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
    from posidonius_b200.shard import gather_summaries, reduce_timing, shard_range
    workload = args.workload
    cfg_index, n_base, flops_per_step = WORKLOADS[workload]
    case, tables = load_case(workload, args.history)
    n_total = args.systems or n_base
    spc = args.steps_per_call
    arithmetic = ARITH[args.arithmetic]
    first, last = shard_range(n_total, rank, world)
    n_sys = last - first
    # this rank's members of the ONE global ensemble, built on the device (pb200_ensemble_create_perturbed_range)
    ens = Ensemble.perturbed(case, tables, n_sys, SEED + cfg_index, AMPLITUDE, device=local, arithmetic=arithmetic, first_member=first)
    ens.initialize_physical_values()
    ens.synchronize()
    # pinned host buffers of the boundary call
    e2e_fields = ("position", "velocity", "acceleration", "angular_momentum", "spin", "velocity_errors", "angular_momentum_errors",
                  "radius", "radius_of_gyration_2", "moment_of_inertia", "current_time")
    host = ens.make_state_buffers(e2e_fields, pinned=True)
    ens.download(out=host)
    io_bytes = int(sum(a.nbytes for a in host.values()))
    hist_buf = None
    if args.history:
        cap = ens.history_capacity()
        hist_buf = torch.zeros(n_sys * cap * ens.n_particles * 156, dtype=torch.uint8).pin_memory().numpy()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    hist_bytes = [0]

    def drain():
        if hist_buf is not None:
            rec = ens.history_drain(out=hist_buf)
            hist_bytes[0] += rec.nbytes

    e_start, l_start = ens.summary()   # Universe::compute_total_energy / angular momentum of every member (untimed)
    for _ in range(args.warmup):
        ens.iterate(spc, synchronize=True)
        drain()
    fp64_peak = measure_fp64_peak(local, 30.0) if rank == 0 else 0.0

    # ---- device-resident throughput
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ens.launch_count()
    hist_bytes[0] = 0
    barrier()
    t0 = time.perf_counter()
    kernel_ms = 0.0
    for _ in range(args.steps):
        ens.iterate(spc, synchronize=False)
        kernel_ms += ens.last_step_ms()   # CUDA events on the launching stream (synchronizes on the end event)
        drain()
    barrier()
    wall = time.perf_counter() - t0
    launches = ens.launch_count() - launches0
    pieces = ens.last_pieces()
    kernel_name = ens.last_kernel()
    hist_timed = hist_bytes[0]
    clocks = sampler.summary()
    st, warn, _ = ens.status()
    alive = int(np.sum(st == 0))

    # ---- end to end through the boundary with host buffers
    ens.download(out=host)
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        ens.run_host(host, spc)
        drain()
    barrier()
    e2e_wall = time.perf_counter() - t1

    # ---- after the timed regions: per-system summaries {status, t, dE/E, dL/L}, gathered on rank 0 (the one gather of
    # ensemble data; NCCL when there are several ranks; shards may differ by one system: padded rows carry status -1)
    e_end, l_end = ens.summary()
    t_end = ens.get_current_time()
    n_pad = -(-n_total // world)
    rows_np = np.full((n_pad, 4), -1.0)
    rows_np[:n_sys] = np.stack([st.astype(np.float64), t_end, (e_end - e_start) / np.abs(e_start), (l_end - l_start) / np.abs(l_start)], axis=1)
    rows = gather_summaries(torch.from_numpy(rows_np).to("cuda"), dist if world > 1 else None)

    # ---- weak-scaling companion (N > 1): the BASELINE-size ensemble on EVERY GPU, two launches
    weak_ms = 0.0
    if world > 1 and not args.no_weak:
        with Ensemble.perturbed(case, tables, n_total, SEED + cfg_index + 1000 * (rank + 1), AMPLITUDE, device=local, arithmetic=arithmetic) as wens:
            wens.initialize_physical_values()
            wens.iterate(spc, synchronize=True)
            barrier()
            for _ in range(2):
                wens.iterate(spc, synchronize=False)
                weak_ms += wens.last_step_ms()
            barrier()

    elapsed = torch.tensor([kernel_ms * 1e-3, wall, e2e_wall, weak_ms * 1e-3], dtype=torch.float64, device="cuda")
    counts = torch.tensor([alive, n_sys, hist_timed], dtype=torch.int64, device="cuda")
    # the only collective: max of the timed regions, per-rank summaries — after the timed region
    elapsed, counts = reduce_timing(elapsed, counts, dist if world > 1 else None)
    kern_s, wall_s, e2e_s, weak_s = [float(x) for x in elapsed.tolist()]
    units = n_total * spc * args.steps

    if rank == 0:
        rows = rows[rows[:, 0] >= 0]
        value = units / kern_s
        achieved = n_sys * spc * args.steps * flops_per_step / (kernel_ms * 1e-3) / 1e12  # this GPU's dominant kernel
        config = {
            "workload": workload, "systems_total": n_total, "systems_per_gpu": n_sys, "bodies": case.n_particles,
            "whfast_steps_per_step": spc, "time_step_days": case.time_step, "arithmetic": args.arithmetic,
            "effects": {"tides": bool(case.consider_tides), "rotational_flattening": bool(case.consider_rotational_flattening),
                        "general_relativity": bool(case.consider_general_relativity), "evolution": bool(case.consider_evolution)},
            "coordinates": ["Jacobi", "DemocraticHeliocentric", "WHDS"][case.coordinates_type],
            "parallelism": "one global ensemble sharded by contiguous ranges over %d GPU(s), no collective on the hot path" % world,
            "time_slices_per_launch": int(pieces), "kernel_build": kernel_name,
            "l2": ("inputs larger than L2: %.0f MB of state per GPU; it crosses HBM once per time slice and stays in registers / "
                   "shared memory in between" % (io_bytes / 1e6)) if io_bytes > 126e6 / 2 else
                  ("state %.0f MB per GPU; the kernel keeps it in registers / shared memory for all the steps of a launch and "
                   "reads it from memory once per time slice, so cache residency between launches does not enter the timing" % (io_bytes / 1e6)),
            "wall_clock_value": units / wall_s, "systems_alive": int(counts[0]), "systems_counted": int(counts[1]),
            "ensemble_summary": {"gathered_systems": int(rows.shape[0]), "max_abs_dE_over_E": float(rows[:, 2].abs().max()),
                                 "max_abs_dL_over_L": float(rows[:, 3].abs().max()), "t_days_min": float(rows[:, 1].min()),
                                 "t_days_max": float(rows[:, 1].max())},
        }
        if weak_s > 0.0:
            config["weak_value"] = world * n_total * spc * 2 / weak_s
            config["weak_note"] = "%d systems on each of the %d GPUs, 2 launches of %d steps, max over ranks" % (n_total, world, spc)
        if args.history:
            config["history"] = {"snapshot_period_days": case.historic_snapshot_period,
                                 "record_bytes_drained_in_timed_region": int(counts[2]),
                                 "drain_gbs_of_wall": int(counts[2]) / wall_s / 1e9,
                                 "note": "156-byte records of every body packed on the device and copied to pinned host memory after each "
                                         "launch; the drain is inside wall_clock_value and e2e, not inside the kernel-event time of value"}
        line = {
            "metric": "TRAPPIST-1 ensemble system-steps/s" if workload == DEFAULT_WORKLOAD else workload + " ensemble system-steps/s",
            "value": value, "unit": "system-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * kern_s / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "clocks": clocks,
            "e2e": {"value": units / e2e_s, "unit": "system-steps/s", "h2d_bytes_per_step": io_bytes, "d2h_bytes_per_step": io_bytes},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
                         "frac": achieved / (fp64_peak / 1e12) if fp64_peak else None,
                         "traffic": measured_traffic(workload, n_sys, spc, args.arithmetic),
                         "peak_source": "measured here: DFMA-chain microbenchmark (pb200_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 figure",
                         "frac_of_theoretical_37.2": achieved / FP64_THEORETICAL_TFLOPS,
                         "flops_per_system_step": flops_per_step},
        }
        if world == 1 and not args.no_other_workloads:
            config["other_workloads"] = other_workloads(local, fp64_peak, arithmetic, None if args.systems else workload)
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(workload, case, tables, SEED + cfg_index)
        sys.stdout.flush()
        os.write(result_fd, (json.dumps(line) + "\n").encode())
    ens.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
