"""The `posidonius-b200 start | resume | ensemble` command line (host C++ over the C ABI) on a GPU box."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, config_case

pytestmark = pytest.mark.gpu

CLI = os.path.join(ROOT, "posidonius_b200", "bin", "posidonius-b200")
# the reader dtype of the reference's posidonius/analysis/history.py:21-28
HISTORY_DTYPE = np.dtype([("current_time", "<f8"), ("time_step", "<f8"), ("particle", "<i4")] + [(k, "<f8") for k in (
    "position_x", "position_y", "position_z", "spin_x", "spin_y", "spin_z", "velocity_x", "velocity_y", "velocity_z", "mass", "radius",
    "radius_of_gyration_2", "love_number", "scaled_dissipation_factor", "lag_angle", "denergy_dt", "migration_timescale")])


def run(*args, expect=0, env=None):
    p = subprocess.run([CLI] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600,
                       env=dict(os.environ, **(env or {})))
    assert p.returncode == expect, p.stdout
    return p.stdout


def _case(tmp_path, name, time_limit, hist_period, rec_period):
    d = config_case(name)
    d["universe"]["time_limit"] = time_limit
    d["historic_snapshot_period"] = hist_period
    d["recovery_snapshot_period"] = rec_period
    p = tmp_path / "case.json"
    p.write_text(json.dumps(d))
    return d, p


def test_start_writes_the_reference_history_layout_and_matches_the_oracle(tmp_path):
    from oracle.binding import OracleSystem
    from posidonius_b200.case import case_from_dict
    d, case_path = _case(tmp_path, "c3_case7_evolving", 40.0, 8.0, 16.0)
    out = run("start", case_path, tmp_path / "rec.bin", tmp_path / "hist.bin", "--silent", "--strict")
    assert "Simulation completed" in out
    case, tables = case_from_dict(d)
    o = OracleSystem(case, tables)
    o.initialize_physical_values()
    o.iterate(10 ** 6)
    want = np.frombuffer(o.history(), dtype=HISTORY_DTYPE)
    got = np.fromfile(tmp_path / "hist.bin", dtype=HISTORY_DTYPE)
    assert got.shape == want.shape and len(got) == 5 * case.n_particles   # t = 0, 8, 16, 24, 32
    for name in HISTORY_DTYPE.names:
        a, b = got[name], want[name]
        if name == "denergy_dt":
            nan = np.isnan(a) & np.isnan(b)
            assert np.all(nan | (np.abs(a - b) <= 1e-12 * np.max(np.abs(b[~nan])))), name
        else:
            assert np.array_equal(a, b), name   # strict arithmetic: bit for bit
    # start refuses to overwrite (main.rs:144-148)
    out = run("start", case_path, tmp_path / "rec.bin", tmp_path / "hist2.bin", "--silent", expect=101)
    assert "already exists" in out


def test_resume_continues_bit_for_bit(tmp_path):
    from posidonius_b200.case import load_case_file
    # uninterrupted run to t = 64
    a = tmp_path / "a"
    a.mkdir()
    _, case_a = _case(a, "c2_case3", 64.0, 8.0, 16.0)
    run("start", case_a, a / "rec.bin", a / "hist.bin", "--silent")
    # same case stopped at t = 40, then resumed with a new time limit of 64
    b = tmp_path / "b"
    b.mkdir()
    _, case_b = _case(b, "c2_case3", 40.0, 8.0, 16.0)
    run("start", case_b, b / "rec.bin", b / "hist.bin", "--silent")
    rec, _ = load_case_file(b / "rec.bin")
    assert 0 < rec.current_time <= 40.0 and rec.last_recovery_snapshot_time == rec.current_time
    out = run("resume", b / "rec.bin", b / "hist.bin", "--silent", "--time-limit", 64)
    assert "Restored previous simulation" in out and "Simulation completed" in out
    ha = np.fromfile(a / "hist.bin", dtype=HISTORY_DTYPE)
    hb = np.fromfile(b / "hist.bin", dtype=HISTORY_DTYPE)
    assert ha.shape == hb.shape
    for name in HISTORY_DTYPE.names:
        if name == "denergy_dt":
            continue   # its tidal scratch is not part of the recovery image (first record after a resume may differ)
        assert np.array_equal(ha[name], hb[name]), name
    fa, _ = load_case_file(a / "rec.bin")
    fb, _ = load_case_file(b / "rec.bin")
    assert fa.current_time == fb.current_time
    for i in range(fa.n_particles):
        assert fa.bodies[i].inertial_position[:] == fb.bodies[i].inertial_position[:]
        assert fa.bodies[i].inertial_velocity[:] == fb.bodies[i].inertial_velocity[:]
        assert fa.bodies[i].angular_momentum[:] == fb.bodies[i].angular_momentum[:]


def test_recovery_snapshots_are_written_once_per_recovery_period(tmp_path):
    """main.rs:158-176 + whfast.rs:237-239, 303-308: a recovery snapshot after the very first step and then whenever
    last_recovery_snapshot_time + period <= current_time at the start of a step — counted over four recovery periods."""
    _, case_path = _case(tmp_path, "c2_case3", 64.0, 8.0, 16.0)
    out = run("start", case_path, tmp_path / "rec.bin", tmp_path / "hist.bin", "--silent", env={"PB200_CLI_TRACE": "1"})
    got = [float(line.split("t = ")[1].split()[0]) for line in out.splitlines() if line.startswith("[TRACE] recovery snapshot")]
    # the reference's loop on the accumulated clock
    t, dt, last_rec, last_hist, want = 0.0, 0.08, -1.0, -1.0, []
    while True:
        first = last_hist < 0.0
        trigger = last_rec + 16.0 <= t
        if first or last_hist + 8.0 <= t:
            last_hist = 0.0 if first else last_hist + 8.0
        t += dt
        if t + dt > 64.0:
            break
        if first or trigger:
            last_rec = t
            want.append(t)
    assert len(want) == 4 and got == want, (got, want)


def test_ensemble_subcommand_stays_inside_a_small_history_buffer(tmp_path):
    """More snapshot periods than history slots: the subcommand bounds its calls by pb200_ensemble_history_capacity."""
    d = config_case("c2_case3")
    d["historic_snapshot_period"] = 0.8   # every 10 steps
    p = tmp_path / "case.json"
    p.write_text(json.dumps(d))
    out_dir = tmp_path / "ens"
    out = run("ensemble", p, out_dir, "--systems", 512, "--steps", 200, "--silent", env={"PB200_HISTORY_SLOTS": "3"})
    t, last_hist, n_snap = 0.0, -1.0, 0   # whfast.rs:237-253 on the accumulated clock
    for _ in range(200):
        first = last_hist < 0.0
        if first or last_hist + 0.8 <= t:
            n_snap += 1
            last_hist = 0.0 if first else last_hist + 0.8
        t += 0.08
    assert n_snap > 3
    assert "ensemble of 512 systems x 200 steps: 512 running" in out and "%d snapshot(s) per system" % n_snap in out
    hist = np.fromfile(out_dir / "ensemble_history.bin", dtype=HISTORY_DTYPE)
    assert len(hist) == 512 * 2 * n_snap


def test_ensemble_subcommand(tmp_path):
    d = config_case("c4_trappist1")
    p = tmp_path / "case.json"
    p.write_text(json.dumps(d))
    out_dir = tmp_path / "ens"
    out = run("ensemble", p, out_dir, "--systems", 96, "--steps", 300, "--seed", 7, "--silent")
    assert "ensemble of 96 systems x 300 steps: 96 running" in out
    rows = (out_dir / "summary.csv").read_text().strip().splitlines()
    assert len(rows) == 97
    cols = rows[1].split(",")
    assert cols[1] == "0" and abs(float(cols[4]) - 300 * 0.08) < 1e-9
    hist = np.fromfile(out_dir / "ensemble_history.bin", dtype=HISTORY_DTYPE)
    assert len(hist) == 96 * 8 and np.all(hist["current_time"] == 0.0)
    assert (out_dir / "recovery_000095.bin").exists()


def test_unsupported_case_is_rejected_not_run_on_a_fallback(tmp_path):
    d = config_case("c4_trappist1")
    d["universe"]["consider_effects"]["disk"] = True
    p = tmp_path / "case.json"
    p.write_text(json.dumps(d))
    out = run("start", p, tmp_path / "rec.bin", tmp_path / "hist.bin", expect=101)
    assert "outside the B200 hot path" in out
