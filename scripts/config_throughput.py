#!/usr/bin/env python
"""Throughput of every BASELINE.json configuration on one GPU (system-steps/s and algorithmic TFLOP/s with the exact
flop counts of the oracle's counting build). usage: config_throughput.py [steps_per_launch] [modes, e.g. hybrid,fast,strict]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import config_case  # noqa: E402
from oracle.binding import count_flops  # noqa: E402
from posidonius_b200.case import case_from_dict  # noqa: E402
from posidonius_b200.ensemble import Ensemble, measure_fp64_peak  # noqa: E402
from posidonius_b200.perturb import make_ensemble_cases  # noqa: E402

CONFIGS = [("c1_example", 65536), ("c2_case3", 4096), ("c2_case3", 65536), ("c3_case7", 16384), ("c3_case7", 65536), ("c3_case7_evolving", 16384),
           ("c3_case7_evolving", 65536), ("c4_trappist1", 65536), ("c5_circumbinary", 65536)]
if os.environ.get("PB200_CONFIGS"):   # e.g. PB200_CONFIGS=c1_example:65536,c2_case3:4096
    CONFIGS = [(c.split(":")[0], int(c.split(":")[1])) for c in os.environ["PB200_CONFIGS"].split(",")]


def main():
    from posidonius_b200 import abi
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    modes = (sys.argv[2] if len(sys.argv) > 2 else "hybrid").split(",")
    mode_id = {"fast": abi.ARITH_FAST, "strict": abi.ARITH_STRICT, "hybrid": abi.ARITH_HYBRID}
    peak = measure_fp64_peak(0, 30.0)
    print("FP64 peak (DFMA chains) %.2f TFLOP/s" % (peak / 1e12))
    print("%-22s %-6s %8s %6s %14s %10s %8s" % ("config", "arith", "systems", "bodies", "system-steps/s", "TFLOP/s", "of peak"))
    for name, n_sys in CONFIGS:
      for mode in modes:
          d = config_case(name)
          d["universe"]["time_limit"] = 1e12
          d["historic_snapshot_period"] = 1e11
          case, tables = case_from_dict(d)
          flops = count_flops(case, tables, 100)["flops_per_step"]
          cases = make_ensemble_cases(case, n_sys, 5)
          with Ensemble(cases, tables, arithmetic=mode_id[mode]) as ens:
              ens.initialize_physical_values()
              ens.iterate(steps)
              best = 1e30
              for _ in range(3):
                  ens.iterate(steps, synchronize=False)
                  best = min(best, ens.last_step_ms())
              st, _, _ = ens.status()
          rate = n_sys * steps / (best * 1e-3)
          print("%-22s %-6s %8d %6d %14.4g %10.2f %7.1f%%   alive %d" % (name, mode, n_sys, case.n_particles, rate, rate * flops / 1e12,
                                                                  100.0 * rate * flops / peak, int((st == 0).sum())))


if __name__ == "__main__":
    main()
