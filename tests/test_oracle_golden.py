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
