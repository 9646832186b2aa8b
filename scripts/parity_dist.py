#!/usr/bin/env python
"""Distribution of the GPU-vs-oracle relative error after N steps, per configuration and arithmetic mode.

    python scripts/parity_dist.py [--members 1024] [--steps 10000] [--modes hybrid,fast] [--configs c5_circumbinary,...]

Per member: max over bodies of |x_gpu - x_oracle| / |x_oracle| (vector norms; the 1e-10 criterion of BASELINE.json).
Prints median / 99th percentile / max over the members for r, v, spin, and the fraction of members bit-identical in r and v.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import CONFIG_NAMES, config_case  # noqa: E402
from parity_util import gpu_state_of, oracle_state_of  # noqa: E402


def member_err(got, want):
    num = np.linalg.norm(got - want, axis=-1)
    den = np.linalg.norm(want, axis=-1)
    den = np.where(den > 0, den, 1.0)
    return np.max(num / den, axis=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--members", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=10000)
    ap.add_argument("--modes", default="hybrid,fast")
    ap.add_argument("--configs", default="", help="comma-separated subset of the configuration names")
    args = ap.parse_args()
    from oracle.binding import run_ensemble
    from posidonius_b200 import abi
    from posidonius_b200.case import case_from_dict
    from posidonius_b200.ensemble import Ensemble
    from posidonius_b200.perturb import make_ensemble_cases
    modes = {"fast": abi.ARITH_FAST, "strict": abi.ARITH_STRICT, "hybrid": abi.ARITH_HYBRID}
    print("members %d  steps %d  (per member: max over bodies of the relative error; median / p99 / max over members)" % (args.members, args.steps))
    for idx, name in enumerate(CONFIG_NAMES):
        if args.configs and name not in args.configs.split(","):
            continue
        case, tables = case_from_dict(config_case(name))
        cases = make_ensemble_cases(case, args.members, 20261017 + idx)
        t0 = time.time()
        oc, ost, _ = run_ensemble(cases, args.members, tables, args.steps, True, os.cpu_count() or 1)
        o = oracle_state_of(oc)
        t_or = time.time() - t0
        for mode in args.modes.split(","):
            with Ensemble(cases, tables, arithmetic=modes[mode]) as ens:
                ens.initialize_physical_values()
                ens.iterate(args.steps)
                g = gpu_state_of(ens)
                st, w, _ = ens.status()
            same = np.all(g["position"] == o["position"], axis=(1, 2)) & np.all(g["velocity"] == o["velocity"], axis=(1, 2))
            cols = []
            for k in ("position", "velocity", "spin"):
                e = member_err(g[k], o[k])
                cols.append("%s %.1e / %.1e / %.1e" % (k[:3], np.median(e), np.percentile(e, 99), e.max()))
            print("%-18s %-6s  %s   bit-identical r,v %5.1f %%   status equal %s  warnings %s  (oracle %.1f s)"
                  % (name, mode, "   ".join(cols), 100.0 * same.mean(), bool(np.array_equal(st, ost)), sorted(set(w.tolist())), t_or))
            sys.stdout.flush()


if __name__ == "__main__":
    main()
