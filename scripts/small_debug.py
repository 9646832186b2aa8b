#!/usr/bin/env python
"""Development check of the lane = planet builds (small_step.cuh) against the run-time-geometry kernel and the CPU oracle.

usage: small_debug.py [steps ...]   (default 1 2 10 300)
For every small configuration and arithmetic mode: state after n steps from the lane = planet build, from the generic
kernel (PB200_FORCE_GENERIC=1) and, in strict mode, from the oracle; prints the worst relative difference per field and
the share of bit-identical members.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import config_case  # noqa: E402
from parity_util import gpu_state_of, oracle_state_of, rel_err  # noqa: E402
from posidonius_b200.case import case_from_dict  # noqa: E402
from posidonius_b200.ensemble import Ensemble  # noqa: E402
from posidonius_b200.perturb import make_ensemble_cases  # noqa: E402

FIELDS = ("position", "velocity", "spin", "angular_momentum", "velocity_errors", "angular_momentum_errors", "acceleration")


def run(cases, tables, steps, arith, generic):
    os.environ["PB200_FORCE_GENERIC"] = "1" if generic else "0"
    with Ensemble(cases, tables, arithmetic=arith) as ens:
        ens.initialize_physical_values()
        ens.iterate(steps)
        g = gpu_state_of(ens)
        st, w, it = ens.status()
    return g, st, w, it


def main():
    steps_list = [int(a) for a in sys.argv[1:]] or [1, 2, 10, 300]
    from oracle.binding import run_ensemble
    n_sys = 77
    for name in ("c1_example", "c2_case3", "c3_case7", "c3_case7_evolving", "c5_circumbinary"):
        case, tables = case_from_dict(config_case(name))
        cases = make_ensemble_cases(case, n_sys, 13)
        for steps in steps_list:
            oc, ost, _ = run_ensemble(cases, n_sys, tables, steps, True, os.cpu_count() or 1)
            o = oracle_state_of(oc)
            for arith in (1, 2, 0):
                a, st, w, it = run(cases, tables, steps, arith, False)
                b, st2, w2, it2 = run(cases, tables, steps, arith, True)
                line = "%-18s steps %5d arith %d |" % (name, steps, arith)
                for k in FIELDS:
                    line += " %s %.1e/%.1e" % (k[:3] + k[-3:], rel_err(a[k], b[k]), rel_err(a[k], o[k]))
                same = np.all(a["position"] == o["position"], axis=(1, 2)) & np.all(a["velocity"] == o["velocity"], axis=(1, 2))
                line += " | bit-id vs oracle %.2f  status eq %s/%s warn %s it %s" % (same.mean(), np.array_equal(st, st2), np.array_equal(st, ost),
                                                                                   sorted(set(w.tolist())), sorted(set(it.tolist()))[:3])
                print(line, flush=True)


if __name__ == "__main__":
    main()
