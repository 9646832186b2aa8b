#!/usr/bin/env python
"""Diagnostic run on a GPU box: CUDA ensemble vs CPU oracle on the golden fixtures and the configs.
Prints relative errors instead of asserting (used while bringing kernels up)."""
import gzip
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.binding import OracleSystem, run_ensemble  # noqa: E402
from parity_util import gpu_state_of, oracle_state_of, rel_err  # noqa: E402
from posidonius_b200.case import case_from_dict  # noqa: E402
from posidonius_b200.ensemble import Ensemble, measure_fp64_peak  # noqa: E402
from posidonius_b200.perturb import make_ensemble_cases  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")


def load(rel):
    with gzip.open(os.path.join(G, rel), "rt") as f:
        return json.load(f)


def main():
    steps_cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    arith = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    n_members = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    man = json.load(open(os.path.join(G, "manifest.json")))
    print("fp64 peak (DFMA chain): %.2f TFLOP/s" % (measure_fp64_peak() / 1e12))
    for name, fx in sorted(man["fixtures"].items()):
        case, tables = case_from_dict(load(fx["case"]))
        ens = Ensemble(case, tables, n_systems=3)
        ens.initialize_physical_values()
        ens.iterate(100000)
        st, w, it = ens.status()
        out = ens.get_case(1)
        err = 0.0
        for i, exp in enumerate(fx["particles"]):
            for key in ("inertial_position", "inertial_velocity", "inertial_acceleration"):
                got = np.array(getattr(out.bodies[i], key)[:])
                want = np.array([exp[key]["x"], exp[key]["y"], exp[key]["z"]])
                err = max(err, rel_err(got, want))
        print("%-48s status %s warn %s iter %s  rel err vs golden %.2e" % (name, st[:2], w[:1], it[:1], err))
        ens.close()
    for idx, name in enumerate(["c1_example", "c2_case3", "c3_case7", "c3_case7_evolving", "c4_trappist1", "c5_circumbinary"]):
        case, tables = case_from_dict(load("configs/%s.json.gz" % name))
        n_sys = n_members
        cases = make_ensemble_cases(case, n_sys, 20261017 + idx)
        ens = Ensemble(cases, tables, arithmetic=arith)
        ens.initialize_physical_values()
        t0 = time.time()
        ens.iterate(steps_cfg)
        tg = time.time() - t0
        g = gpu_state_of(ens)
        st, w, it = ens.status()
        t0 = time.time()
        oc, ost, secs = run_ensemble(cases, n_sys, tables, steps_cfg, True, 8)
        o = oracle_state_of(oc)
        errs = {k: rel_err(g[k], o[k]) for k in ("position", "velocity", "angular_momentum", "spin")}
        exact = np.mean(np.all(g["position"] == o["position"], axis=(1, 2)) & np.all(g["velocity"] == o["velocity"], axis=(1, 2)))
        errs["bit-identical r,v systems"] = exact
        e, l = ens.summary()
        print("%-20s N=%d steps=%d gpu %.2fs (kernel %.1f ms) oracle %.2fs status gpu %s oracle %s warn %s | %s" % (
            name, case.n_particles, steps_cfg, tg, ens.last_step_ms(), secs, np.unique(st), np.unique(ost), np.unique(w),
            " ".join("%s %.1e" % kv for kv in errs.items())))
        ens.close()


if __name__ == "__main__":
    main()
