"""Pins the CPU oracle to the reference's own golden vectors (tests/golden/manifest.json).

The reference asserts |value - stored| < 1e-14 on inertial position, velocity and (Newtonian)
acceleration of every particle after running each fixture to completion (199 steps)
— reference tests/common/universe.rs:48-72. The oracle has to meet the same bar.
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_json_gz
from oracle.binding import OracleSystem
from posidonius_b200 import abi
from posidonius_b200.case import case_from_dict

with open(os.path.join(GOLDEN, "manifest.json")) as _f:
    _MANIFEST = json.load(_f)


@pytest.mark.parametrize("name", sorted(_MANIFEST["fixtures"]))
def test_oracle_reproduces_reference_golden(name):
    fx = _MANIFEST["fixtures"][name]
    case, tables = case_from_dict(load_json_gz(fx["case"]))
    o = OracleSystem(case, tables)
    assert o.initialize_physical_values() == 0
    steps = o.iterate(10 ** 6)
    assert steps == 199  # the accumulated-time termination test gives 199, not 200
    status, warnings, _ = o.status()
    assert status == abi.STATUS_COMPLETED and warnings == 0
    out = o.case()
    tol = fx["tolerance_abs"]
    for i, exp in enumerate(fx["particles"]):
        for key in ("inertial_position", "inertial_velocity", "inertial_acceleration"):
            got = np.array(getattr(out.bodies[i], key)[:])
            want = np.array([exp[key]["x"], exp[key]["y"], exp[key]["z"]])
            assert np.all(np.abs(got - want) < tol), (name, i, key, got, want)


def test_oracle_history_record_layout():
    fx = _MANIFEST["fixtures"]["test_integrator-whfast_jacobi"]
    case, tables = case_from_dict(load_json_gz(fx["case"]))
    o = OracleSystem(case, tables)
    o.initialize_physical_values()
    o.iterate(1)
    raw = o.history()
    n = case.n_particles
    assert len(raw) == n * abi.HISTORIC_RECORD_BYTES
    # reader dtype of the reference's posidonius/analysis/history.py:21-28
    dt = np.dtype([("current_time", "<f8"), ("time_step", "<f8"), ("particle", "<i4")] +
                  [(k, "<f8") for k in ("position_x", "position_y", "position_z", "spin_x", "spin_y", "spin_z",
                                         "velocity_x", "velocity_y", "velocity_z", "mass", "radius",
                                         "radius_of_gyration_2", "love_number", "scaled_dissipation_factor",
                                         "lag_angle", "denergy_dt", "migration_timescale")])
    assert dt.itemsize == abi.HISTORIC_RECORD_BYTES
    rec = np.frombuffer(raw, dtype=dt)
    assert list(rec["particle"]) == list(range(n))
    assert np.all(rec["current_time"] == 0.0) and np.all(rec["time_step"] == case.time_step)
    assert rec["mass"][0] == case.bodies[0].mass
    assert rec["position_x"][1] == case.bodies[1].inertial_position[0]


@pytest.mark.skipif(not os.path.isdir("/root/reference/posidonius"), reason="needs the reference's Python package (build container only)")
def test_history_file_is_readable_by_the_reference_reader(tmp_path):
    """a history file in our record layout, read back with the reference's OWN reader
    (posidonius/analysis/history.py:14-46), imported from a scratch copy next to an empty input/ directory (the package
    refuses to import without one). The same bytes come off the GPU: tests/test_gpu_parity.py compares the device records
    with these oracle records field by field."""
    import shutil
    import subprocess
    import sys
    fx = _MANIFEST["fixtures"]["test_evolution-solar_like_bolmontmathis2016"]
    d = load_json_gz(fx["case"])
    d["historic_snapshot_period"] = 0.8   # every 10 steps
    case, tables = case_from_dict(d)
    o = OracleSystem(case, tables)
    o.initialize_physical_values()
    o.iterate(95)
    path = tmp_path / "case_history.bin"
    path.write_bytes(o.history())
    pkg = tmp_path / "pkg"
    shutil.copytree("/root/reference/posidonius", pkg / "posidonius")
    (pkg / "input").mkdir()
    script = (
        "import json, sys\n"
        "import numpy as np\n"
        "from posidonius.analysis import history\n"
        "n, data = history.read(sys.argv[1])\n"
        "star, planets, keys = history.classify(n, data.copy(), reference_particle_index=0)\n"
        "print(json.dumps({'n': int(n), 'rows': int(len(data)), 'particle': [int(x) for x in data['particle'][:n]],\n"
        "  'times': sorted(set(float(x) for x in data['current_time'])), 'dt': sorted(set(float(x) for x in data['time_step'])),\n"
        "  'mass0': float(data['mass'][0]), 'radius1': float(data['radius'][1]), 'lag': [float(x) for x in data['lag_angle'][n:2 * n]],\n"
        "  'n_planets': len(keys), 'star_rows': int(len(star))}))\n")
    out = subprocess.run([sys.executable, "-c", script, str(path)], cwd=str(pkg), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert out.returncode == 0, out.stderr[-2000:]
    r = json.loads(out.stdout.strip().splitlines()[-1])
    n = case.n_particles
    assert r["n"] == n and r["rows"] == 10 * n and r["particle"] == list(range(n))
    # snapshots fall on the first step at or after each multiple of the period on the ACCUMULATED clock (whfast.rs:237-239)
    assert len(r["times"]) == 10 and np.all(np.abs(np.array(r["times"]) - 0.8 * np.arange(10)) < 0.0801) and r["dt"] == [case.time_step]
    assert r["mass0"] == case.bodies[0].mass and r["radius1"] == case.bodies[1].radius
    assert r["lag"][0] > 0.0 and all(x == 0.0 for x in r["lag"][1:])   # the star's dynamical-tide lag angle (evolution.rs:548-567)
    assert r["n_planets"] == n - 1 and r["star_rows"] == 10
