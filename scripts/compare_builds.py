#!/usr/bin/env python
"""Bit-for-bit comparison of two builds of the library on the same ensemble (tuning aid: a change that only removes
redundant work must leave every bit of the state where it was).
usage: compare_builds.py libA.so libB.so [config] [systems] [steps]   (runs itself once per library in a subprocess)"""
import gzip
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_one(out, config, n_sys, steps):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from parity_util import gpu_state_of
    from posidonius_b200.case import case_from_dict
    from posidonius_b200.ensemble import Ensemble
    from posidonius_b200.perturb import make_ensemble_cases
    with gzip.open(os.path.join(ROOT, "tests", "golden", "configs", config + ".json.gz"), "rt") as f:
        d = json.load(f)
    d["universe"]["time_limit"] = 1e12
    case, tables = case_from_dict(d)
    cases = make_ensemble_cases(case, n_sys, 7)
    with Ensemble(cases, tables, device=0) as ens:
        ens.initialize_physical_values()
        ens.iterate(steps)
        g = gpu_state_of(ens)
    np.savez(out, **{k: np.asarray(v) for k, v in g.items()})


def main():
    if sys.argv[1] == "--one":
        return run_one(sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5]))
    a, b = sys.argv[1:3]
    config = sys.argv[3] if len(sys.argv) > 3 else "c4_trappist1"
    n_sys = sys.argv[4] if len(sys.argv) > 4 else "1000"
    steps = sys.argv[5] if len(sys.argv) > 5 else "700"
    outs = []
    for i, lib in enumerate((a, b)):
        out = "/tmp/pb200_cmp_%d.npz" % i
        env = dict(os.environ, PB200_LIB=os.path.abspath(lib))
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--one", out, config, n_sys, steps], env=env)
        outs.append(np.load(out))
    same = True
    for k in outs[0].files:
        x, y = outs[0][k], outs[1][k]
        eq = np.array_equal(x, y, equal_nan=True) if x.dtype.kind == "f" else np.array_equal(x, y)
        same &= bool(eq)
        print("%-12s %s" % (k, "identical" if eq else "DIFFERENT (max abs %.3e)" % float(np.nanmax(np.abs(x - y)))))
    print("BIT-IDENTICAL" if same else "NOT IDENTICAL")
    return 0 if same else 1


if __name__ == "__main__":
    sys.exit(main())
