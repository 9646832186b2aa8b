#!/usr/bin/env python
"""Generates the five BASELINE.json configurations as JSON cases under tests/golden/configs/.

Run in the build container only (needs /root/reference; the GPU box uses the committed outputs):
    python tests/golden/make_configs.py

The JSON is produced by the reference's OWN Python case generator (`posidonius/`, imported from a
scratch copy next to an empty `input/` directory, as SURVEY.md §8c describes) running the reference's
own case scripts. Two scripts need the stellar-evolution tables of the separate `input/` download,
which is absent; for those the script is run with the star NonEvolving and the evolver (table,
evolution type, radius / radius of gyration / angular momentum adjustment) is patched in afterwards,
restating posidonius/particles/universe.py:131-156 (= src/particles/universe.rs:127-157), with the
table taken from the reference's own test fixtures.

  c1_example.json.gz             cases/example.py
  c2_case3.json.gz               cases/Bolmont_et_al_2015/case3.py
  c3_case7.json.gz               cases/Bolmont_et_al_2015/case7.py (as shipped: non-evolving)
  c3_case7_evolving.json.gz      case7 with Leconte2011(0.08) on the host, initial_time = 4.5e6 yr
  c4_trappist1.json.gz           cases/trappist1.py
  c5_circumbinary.json.gz        cases/example_circumbinary_planet.py switched to WHFast/Jacobi,
                                 Baraffe2015(1.0) evolution on star 2
"""
import gzip
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "configs")


def run_case(workdir, script_text, name):
    script = os.path.join(workdir, name + ".py")
    with open(script, "w") as f:
        f.write(script_text)
    out = os.path.join(workdir, name + ".json")
    env = dict(os.environ, PYTHONPATH=workdir)
    subprocess.check_call([sys.executable, script, out], env=env, cwd=workdir, stdout=subprocess.DEVNULL)
    return json.load(open(out))


def linear_interpolation(t, x, y):
    """tools.rs:840-907 semantics (first x[i] > t)."""
    i = next((k for k, v in enumerate(x) if v > t), None)
    if i is None:
        return y[-1] if not (x[-1] > t) else y[0]
    if i == 0:
        return y[0]
    pct = (t - x[i - 1]) / (x[i] - x[i - 1])
    return y[i - 1] * (1.0 - pct) + y[i] * pct


def attach_evolver(case, body, evolution, table, table_initial_time):
    """Patch an evolving body into a NonEvolving case (universe.py:125-156)."""
    u = case["universe"]
    shift = table_initial_time - u["initial_time"]
    time = [t + shift for t in table["time"]]
    if time[0] > 0.0:
        raise RuntimeError("initial time younger than the table")
    ev = {"evolution": evolution, "time": time, "radius": table["radius"],
          "radius_of_gyration_2": table["radius_of_gyration_2"], "love_number": table["love_number"],
          "inverse_tidal_q_factor": table["inverse_tidal_q_factor"], "left_index": 0}
    u["particles_evolvers"][body] = ev
    u["consider_effects"]["evolution"] = True
    p = u["particles"][body]
    p["evolution"] = evolution
    update = False
    if ev["radius"]:
        r = linear_interpolation(0.0, time, ev["radius"])
        if abs(r - p["radius"]) > 1e-6:
            p["radius"] = r
            update = True
    if ev["radius_of_gyration_2"]:
        g = linear_interpolation(0.0, time, ev["radius_of_gyration_2"])
        if abs(g - p["radius_of_gyration_2"]) > 1e-6:
            p["radius_of_gyration_2"] = g
            update = True
    if update:
        p["moment_of_inertia"] = float(p["mass"]) * float(p["radius_of_gyration_2"]) * float(p["radius"]) * float(p["radius"])
        for ax in "xyz":
            p["angular_momentum"][ax] = p["spin"][ax] * p["moment_of_inertia"]


def dump(case, name):
    raw = json.dumps(case, sort_keys=True).encode()
    with gzip.GzipFile(os.path.join(OUT, name + ".json.gz"), "wb", compresslevel=9, mtime=0) as f:
        f.write(raw)
    print("wrote", name, len(raw), "bytes raw")


def main():
    os.makedirs(OUT, exist_ok=True)
    work = tempfile.mkdtemp(prefix="pb200_cfg_")
    try:
        shutil.copytree(os.path.join(REF, "posidonius"), os.path.join(work, "posidonius"))
        os.makedirs(os.path.join(work, "input"))
        read = lambda rel: open(os.path.join(REF, "cases", rel)).read()
        dump(run_case(work, read("example.py"), "c1"), "c1_example")
        dump(run_case(work, read("Bolmont_et_al_2015/case3.py"), "c2"), "c2_case3")
        c7 = read("Bolmont_et_al_2015/case7.py")
        dump(run_case(work, c7, "c3"), "c3_case7")
        # evolving variant: the script's own commented alternative initial_time (case7.py:14)
        c7e = re.sub(r"^(\s*)initial_time = 1\.0e6\*365\.25.*$", r"\1initial_time = 4.5e6*365.25", c7, count=1, flags=re.M)
        assert c7e != c7
        case = run_case(work, c7e, "c3e")
        fx = json.load(open(os.path.join(REF, "tests/data/test_integrator-whfast_jacobi/case.json")))
        attach_evolver(case, 0, {"Leconte2011": 0.08}, fx["universe"]["particles_evolvers"][0], fx["universe"]["initial_time"])
        dump(case, "c3_case7_evolving")
        dump(run_case(work, read("trappist1.py"), "c4"), "c4_trappist1")
        cb = read("example_circumbinary_planet.py")
        cb2 = cb.replace("star2_evolution = posidonius.Baraffe2015(star2_mass)", "star2_evolution = posidonius.NonEvolving()")
        cb2 = cb2.replace('universe.write(filename, integrator="IAS15")',
                          'universe.write(filename, integrator="WHFast", whfast_alternative_coordinates="Jacobi")')
        assert cb2.count("NonEvolving()") == cb.count("NonEvolving()") + 1 and "IAS15\")" not in cb2.split("#universe.write")[-1]
        case = run_case(work, cb2, "c5")
        fx = json.load(open(os.path.join(REF, "tests/data/test_evolution-solar_like_baraffe2015/case.json")))
        attach_evolver(case, 1, {"Baraffe2015": 1.0}, fx["universe"]["particles_evolvers"][0], fx["universe"]["initial_time"])
        dump(case, "c5_circumbinary")
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
