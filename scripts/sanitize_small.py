#!/usr/bin/env python
"""A short run of every lane = planet build (and, with PB200_PASSIVE_N members, the passive-planet build) for
compute-sanitizer:   compute-sanitizer --tool memcheck|racecheck|initcheck python scripts/sanitize_small.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import config_case  # noqa: E402
from posidonius_b200.case import case_from_dict  # noqa: E402
from posidonius_b200.ensemble import Ensemble  # noqa: E402
from posidonius_b200.perturb import make_ensemble_cases  # noqa: E402

for name, n_sys in (("c1_example", 70), ("c2_case3", 70), ("c3_case7", 70), ("c3_case7_evolving", 70), ("c5_circumbinary", 70),
                    ("c5_circumbinary", int(os.environ.get("PB200_PASSIVE_N", "0"))), ("c3_case7", int(os.environ.get("PB200_PASSIVE_N", "0")))):
    if n_sys == 0:
        continue
    d = config_case(name)
    d["historic_snapshot_period"] = 3 * d["time_step"]
    case, tables = case_from_dict(d)
    cases = make_ensemble_cases(case, n_sys, 3)
    for arith in (2, 1, 0):
        os.environ["PB200_PIECES"] = "2"
        with Ensemble(cases, tables, arithmetic=arith) as ens:
            ens.initialize_physical_values()
            ens.iterate(8)
            st, w, it = ens.status()
            h = ens.history_drain()
        print(name, n_sys, "arith", arith, "status", sorted(set(st.tolist())), "records", h.shape, flush=True)
