#!/usr/bin/env python
"""BASELINE config 5 with historic snapshot output (cases/example_circumbinary_planet.py switched to WHFast/Jacobi,
65536 members, a snapshot every 36525 d = 1826 steps of 20 d): what the snapshots cost.
Prints the step-kernel time of launches with and without a snapshot inside, and the drain
(SoA planes -> 156-byte records on the device, then D2H) with its effective bandwidth.
usage: history_bench.py [n_systems]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import config_case  # noqa: E402
from posidonius_b200.case import case_from_dict  # noqa: E402
from posidonius_b200.ensemble import Ensemble  # noqa: E402
from posidonius_b200.perturb import make_ensemble_cases  # noqa: E402


def main():
    n_sys = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    case, tables = case_from_dict(config_case("c5_circumbinary"))
    n = case.n_particles
    cases = make_ensemble_cases(case, n_sys, 20261022)
    per = int(round(case.historic_snapshot_period / case.time_step))   # 1826 steps
    with Ensemble(cases, tables) as ens:
        ens.initialize_physical_values()
        ens.iterate(1)                      # the first snapshot (t = 0)
        ens.history_drain()
        ens.iterate(per - 2)                # no snapshot inside
        t_plain = ens.last_step_ms() / (per - 2)
        ens.iterate(per)                    # exactly one snapshot inside
        t_snap = ens.last_step_ms() / per
        ens.iterate(8 * per)                # eight snapshots buffered on the device
        pend = ens.history_pending()
        ens.synchronize()
        t0 = time.perf_counter()
        rec = ens.history_drain()
        dt = time.perf_counter() - t0
        st, _, _ = ens.status()
    rec_bytes = rec.nbytes
    plane_bytes = pend * 17 * 8 * n * n_sys
    print("config 5, %d systems x %d bodies, snapshot every %d steps" % (n_sys, n, per))
    print("step kernel: %.4f ms/step without a snapshot, %.4f ms/step with one snapshot per %d steps (+%.2f %%)" %
          (t_plain, t_snap, per, 100.0 * (t_snap / t_plain - 1.0)))
    print("throughput %.3e system-steps/s with snapshot output; %d systems alive" % (n_sys / (t_snap * 1e-3), int((st == 0).sum())))
    print("drain of %d snapshots: %.1f MB of records (%.1f MB of SoA planes read) in %.2f ms wall = %.1f GB/s of records to the host"
          % (pend, rec_bytes / 1e6, plane_bytes / 1e6, dt * 1e3, rec_bytes / dt / 1e9))
    assert rec.shape == (n_sys, pend, n, 156)
    t = np.frombuffer(rec[0, :, 0, :8].tobytes(), dtype="<f8")
    print("snapshot times of member 0 (days):", t.tolist())


if __name__ == "__main__":
    main()
