#!/usr/bin/env python
"""How the hybrid arithmetic departs from the oracle over long horizons: share of members bit-identical in each state array at
a series of step counts (the exact committed evaluation keeps every array bit-identical until an uncommitted iterate flips
a rounding of its input). usage: hybrid_decay.py [config] [members]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import CONFIG_NAMES, config_case  # noqa: E402
from parity_util import gpu_state_of, oracle_state_of  # noqa: E402
from oracle.binding import run_ensemble  # noqa: E402
from posidonius_b200.case import case_from_dict  # noqa: E402
from posidonius_b200.ensemble import Ensemble  # noqa: E402
from posidonius_b200.perturb import make_ensemble_cases  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c4_trappist1"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
case, tables = case_from_dict(config_case(name))
cases = make_ensemble_cases(case, n, 20261017 + CONFIG_NAMES.index(name))
keys = ("position", "velocity", "angular_momentum", "spin", "velocity_errors", "angular_momentum_errors")
print(name, n, "members; share of members bit-identical to the oracle per array")
print("%8s " % "steps" + " ".join("%10s" % k[:10] for k in keys))
with Ensemble(cases, tables) as ens:
    ens.initialize_physical_values()
    done = 0
    for steps in (1000, 3000, 10000, 20000, 50000, 100000):
        ens.iterate(steps - done)
        done = steps
        g = gpu_state_of(ens)
        oc, _, _ = run_ensemble(cases, n, tables, steps, True, os.cpu_count() or 1)
        o = oracle_state_of(oc)
        print("%8d " % steps + " ".join("%10.3f" % np.all(g[k] == o[k], axis=(1, 2)).mean() for k in keys), flush=True)
