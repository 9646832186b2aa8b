#!/usr/bin/env python
"""A short run of the lane = body kernels (run-time geometry on a 5-body golden fixture of every coordinate system, the
8-body build in 64- and 384-thread CTAs) for compute-sanitizer:
    compute-sanitizer --tool memcheck|racecheck python scripts/sanitize_lane_body.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import GOLDEN, config_case, load_json_gz  # noqa: E402
from posidonius_b200.case import case_from_dict  # noqa: E402
from posidonius_b200.ensemble import Ensemble  # noqa: E402
from posidonius_b200.perturb import make_ensemble_cases  # noqa: E402

man = json.load(open(os.path.join(GOLDEN, "manifest.json")))["fixtures"]
runs = [(name, load_json_gz(man[name]["case"]), 9) for name in ("test_integrator-whfast_jacobi", "test_integrator-whfast_whds",
                                                                 "test_general_relativity-newhall1983", "test_evolution-solar_like_bolmontmathis2016")]
runs += [("c4_trappist1", config_case("c4_trappist1"), 70), ("c4_trappist1", config_case("c4_trappist1"), int(os.environ.get("PB200_WIDE_N", "0")))]
for name, d, n_sys in runs:
    if n_sys == 0:
        continue
    d["historic_snapshot_period"] = 3 * d["time_step"]
    case, tables = case_from_dict(d)
    cases = make_ensemble_cases(case, n_sys, 3)
    for arith in (2, 1, 0):
        os.environ["PB200_PIECES"] = "2"
        with Ensemble(cases, tables, arithmetic=arith) as ens:
            ens.initialize_physical_values()
            ens.iterate(6)
            st, w, it = ens.status()
            h = ens.history_drain()
            k = ens.last_kernel()
        print(name, n_sys, "arith", arith, "kernel", k, "status", sorted(set(st.tolist())), "records", h.shape, flush=True)
