"""C++ case I/O of the library (no GPU): JSON reader vs the Python reader, bincode / JSON round trips, the size of the
bincode image against an independently computed layout (no upstream fixture exists), rejections."""
import ctypes as C
import gzip
import json
import os

import numpy as np
import pytest

from conftest import CONFIG_NAMES, GOLDEN, config_case, load_json_gz
from posidonius_b200 import abi
from posidonius_b200.case import (InvalidCaseError, UnsupportedCaseError, case_from_dict, load_case_file, save_case_file)

with open(os.path.join(GOLDEN, "manifest.json")) as _f:
    _MANIFEST = json.load(_f)


def _bytes(case):
    return bytes(memoryview(case))


def _tables_equal(a, b):
    assert len(a) == len(b)
    ta, tb = a.as_ctypes(), b.as_ctypes()
    for i in range(len(a)):
        assert ta[i].n_rows == tb[i].n_rows
        for col in ("time", "radius", "radius_of_gyration_2", "love_number", "inverse_tidal_q_factor"):
            pa, pb = getattr(ta[i], col), getattr(tb[i], col)
            assert bool(pa) == bool(pb), col
            if pa:
                assert np.array_equal(np.ctypeslib.as_array(pa, (ta[i].n_rows,)), np.ctypeslib.as_array(pb, (tb[i].n_rows,))), col


def _write_json(tmp_path, d, name="case.json"):
    p = tmp_path / name
    p.write_text(json.dumps(d))
    return p


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_cpp_json_reader_equals_python_reader_on_configs(tmp_path, name):
    d = config_case(name)
    want, wt = case_from_dict(d)
    got, gt = load_case_file(_write_json(tmp_path, d))
    assert _bytes(got) == _bytes(want)
    _tables_equal(gt, wt)


@pytest.mark.parametrize("name", ["test_integrator-whfast_jacobi", "test_integrator-whfast_whds", "test_evolution-m_dwarf_baraffe1998",
                                  "test_general_relativity-newhall1983"])
def test_cpp_json_reader_equals_python_reader_on_rust_written_fixtures(tmp_path, name):
    d = load_json_gz(_MANIFEST["fixtures"][name]["case"])
    want, wt = case_from_dict(d)
    got, gt = load_case_file(_write_json(tmp_path, d))
    assert _bytes(got) == _bytes(want)
    _tables_equal(gt, wt)


@pytest.mark.parametrize("ext", ["bin", "json"])
@pytest.mark.parametrize("name", ["c4_trappist1", "c3_case7_evolving", "c5_circumbinary"])
def test_recovery_snapshot_round_trip(tmp_path, name, ext):
    case, tables = case_from_dict(config_case(name))
    case.current_time = 123.25
    case.current_iteration = 1541
    case.n_historic_snapshots = 3
    case.last_recovery_snapshot_time = 100.0
    case.last_historic_snapshot_time = 73.05
    case.inertial_velocity_errors[1][0] = 1.25e-20
    case.bodies[1].tides_denergy_dt = -3.5e-17
    path = tmp_path / ("rec." + ext)
    save_case_file(path, case, tables)
    back, bt = load_case_file(path)
    assert _bytes(back) == _bytes(case)
    _tables_equal(bt, tables)
    # writing over an existing snapshot keeps a backup next to it (output.rs:63-69)
    save_case_file(path, case, tables)
    backups = [f for f in os.listdir(tmp_path) if f.startswith("rec.") and f != "rec." + ext]
    assert len(backups) == 1 and backups[0].endswith(".bin")


def _particle_bytes(role_t, role_f, role_g, ref_particle, evo_type, disk_central):
    """Size of one bincode `Particle` (particles/particle.rs:16-52) from the struct definitions, independently of case_io.cpp."""
    f64, u64, u32, axes = 8, 8, 4, 24
    n = u64 + 3 * f64 + 4 * axes + 2 * axes + 4 * f64 + axes + f64 + 2 * axes + 2 * f64
    n += u32 + (u64 if ref_particle else 0)                                    # Reference
    tides_effect = u32 + (u32 + 3 * f64 if role_t != 2 else 0)                 # TidesEffect(TidalModel::ConstantTimeLag{3 f64})
    n += tides_effect + (9 * f64 + axes + 2 * f64) + 2 * axes + 2 * axes       # internal(12 f64 + shape), output, coordinates
    flat_effect = u32 + (u32 + f64 if role_f != 2 else 0)
    n += flat_effect + (8 * f64 + axes) + 2 * axes + 2 * axes
    gr_effect = u32 + (u32 if role_g == 0 else 0)
    n += gr_effect + 5 * f64 + 2 * axes + 2 * axes
    n += u32 + 2 * f64 + f64 + axes                                            # Wind
    n += u32 + (6 * f64 if disk_central else 0) + 4 * f64 + axes + 2 * axes    # Disk
    n += u32 + (0 if evo_type == 6 else (1 if evo_type == 5 else f64))         # EvolutionType
    return n


def test_bincode_image_size_matches_the_struct_layout(tmp_path):
    case, tables = case_from_dict(config_case("c3_case7_evolving"))
    path = tmp_path / "rec.bin"
    save_case_file(path, case, tables)
    n = case.n_particles
    total = 2 * 8 + 2 * 8                                                      # time_step, half_time_step, initial_time, time_limit
    rows = 0
    for i in range(10):
        if i < n:
            b = case.bodies[i]
            total += _particle_bytes(b.tides_role, b.flattening_role, b.general_relativity_role, b.reference >= 0, b.evolution_type, b.disk_role == 0)
        else:
            total += _particle_bytes(2, 2, 2, False, 6, False)
    total += 8                                                                 # Vec<Evolver> length
    tc = tables.as_ctypes()
    for i in range(10):
        b = case.bodies[i]
        has = i < n and b.evolution_table >= 0
        total += 4 + (0 if not has else (1 if b.evolution_type == 5 else 8))   # evolver.evolution
        if has:
            t = tc[b.evolution_table]
            cols = sum(1 for c in ("time", "radius", "radius_of_gyration_2", "love_number", "inverse_tidal_q_factor") if getattr(t, c))
            rows += cols * t.n_rows
        total += 5 * 8 + 8                                                     # five Vec lengths + left_index
    total += rows * 8
    total += 8 + 6 + 4 + 5 * 8 + 5 + 8 + 100 * 8                               # n_particles, consider_effects, GR impl, hosts, map, roche
    total += 8 + 8 + 4 * 8 + 8 + 8                                             # current_time, iteration, periods/last times, n_hist, hash
    total += 10 * (2 * 8 + 3 * 24) + 4 + 8 + 2 * 10 * 24                       # alternative coordinates, type, timestep_warning, errors
    assert os.path.getsize(path) == total


def test_cpp_reader_rejections(tmp_path):
    for name, fx in _MANIFEST["reject"].items():
        with pytest.raises(UnsupportedCaseError):
            load_case_file(_write_json(tmp_path, load_json_gz(fx["case"]), name + ".json"))
    with pytest.raises(InvalidCaseError):
        load_case_file(tmp_path / "does_not_exist.json")
    (tmp_path / "garbage.bin").write_bytes(b"\x00" * 100)
    with pytest.raises((InvalidCaseError, UnsupportedCaseError)):
        load_case_file(tmp_path / "garbage.bin")
    d = config_case("c4_trappist1")
    model = d["universe"]["particles"][1]["tides"]["effect"]["OrbitingBody"]
    d["universe"]["particles"][1]["tides"]["effect"]["OrbitingBody"] = {"Kaula": model["ConstantTimeLag"]}
    with pytest.raises(UnsupportedCaseError):
        load_case_file(_write_json(tmp_path, d, "kaula.json"))


def test_json_recovery_is_readable_by_plain_json_and_keeps_the_reference_key_order(tmp_path):
    case, tables = case_from_dict(config_case("c1_example"))
    path = tmp_path / "rec.json"
    save_case_file(path, case, tables)
    d = json.loads(path.read_text())
    assert list(d)[:3] == ["time_step", "half_time_step", "universe"]           # declaration order of WHFast (whfast.rs:98-120)
    assert list(d["universe"])[:3] == ["initial_time", "time_limit", "particles"]
    back, _ = case_from_dict(d)
    assert _bytes(back) == _bytes(case)
