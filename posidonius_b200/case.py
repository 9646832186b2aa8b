"""JSON case reader/writer: the serde image of `WHFast` <-> the flat `pb200_case_t`.

Host-side mirror of `output::restore_snapshot` / `deserialize_json_snapshot`
(reference src/integrator/output.rs:206-269): the reference tries WHFast, then Ias15,
then LeapFrog; this reader accepts only the WHFast image and raises
`UnsupportedCaseError` for everything the hot path does not cover (IAS15, LeapFrog,
Kaula and creep tides, creep flattening, disk) — there is no CPU fallback.
"""
import ctypes as C
import json

import numpy as np

from . import abi


class UnsupportedCaseError(ValueError):
    """The case needs an effect/integrator outside the B200 hot path (PB200_E_UNSUPPORTED)."""


class InvalidCaseError(ValueError):
    """The case is malformed (PB200_E_INVALID)."""


def _axes(a):
    return (float(a["x"]), float(a["y"]), float(a["z"]))


def _effect_role_and_payload(effect):
    """serde externally-tagged enum: "Disabled" | "OrbitingBody" | {"CentralBody": payload} ..."""
    if isinstance(effect, str):
        return effect, None
    if isinstance(effect, dict) and len(effect) == 1:
        (k, v), = effect.items()
        return k, v
    raise InvalidCaseError("unrecognised effect encoding: %r" % (effect,))


class CaseTables:
    """Owns the evolution-table arrays referenced by a `pb200_table_t[]`."""

    def __init__(self):
        self.columns = []  # list of dicts of np arrays
        self._ctypes = None

    def add(self, time, radius, rg2, love, qinv):
        def arr(x):
            a = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
            return a if a.size else None
        t = arr(time)
        if t is None:
            raise InvalidCaseError("evolving body with an empty time table")
        cols = {"time": t, "radius": arr(radius), "radius_of_gyration_2": arr(rg2), "love_number": arr(love),
                "inverse_tidal_q_factor": arr(qinv)}
        for k, v in cols.items():
            if v is not None and v.size != t.size:
                raise InvalidCaseError("evolution table column %s has %d rows, time has %d" % (k, v.size, t.size))
        self.columns.append(cols)
        self._ctypes = None
        return len(self.columns) - 1

    def __len__(self):
        return len(self.columns)

    def as_ctypes(self):
        if self._ctypes is None:
            n = len(self.columns)
            arr = (abi.Table * max(n, 1))()
            for i, cols in enumerate(self.columns):
                arr[i].n_rows = cols["time"].size
                for k in ("time", "radius", "radius_of_gyration_2", "love_number", "inverse_tidal_q_factor"):
                    v = cols[k]
                    setattr(arr[i], k, v.ctypes.data_as(C.POINTER(C.c_double)) if v is not None else None)
            self._ctypes = arr
        return self._ctypes


def _set_dummy_body(b):
    """Unused particle slots: Particle::new_dummy() (particle.rs:100-146) — every effect Disabled, NonEvolving."""
    C.memset(C.byref(b), 0, C.sizeof(abi.Body))
    b.tides_role = b.flattening_role = b.general_relativity_role = b.disk_role = abi.ROLE_DISABLED
    b.wind_role = 1
    b.evolution_type = abi.EVO_NONEVOLVING
    b.evolution_table = -1
    b.reference = -1


def case_from_dict(d):
    """Flatten the serde JSON image of a WHFast integrator. Returns (abi.Case, CaseTables)."""
    if "alternative_coordinates_type" not in d or "universe" not in d:
        # Ias15 has `time_step_fraction`/`b`,`br`...; LeapFrog has neither alternative coordinates (ias15.rs, leapfrog.rs)
        kind = "IAS15" if "n_particles" in d or "b" in d else "LeapFrog or unknown"
        raise UnsupportedCaseError("only the WHFast integrator image is supported (got %s)" % kind)
    u = d["universe"]
    c = abi.Case()
    c.time_step = d["time_step"]
    c.half_time_step = d["half_time_step"]
    c.initial_time = u["initial_time"]
    c.time_limit = u["time_limit"]
    c.current_time = d["current_time"]
    c.recovery_snapshot_period = d["recovery_snapshot_period"]
    c.historic_snapshot_period = d["historic_snapshot_period"]
    c.last_recovery_snapshot_time = d["last_recovery_snapshot_time"]
    c.last_historic_snapshot_time = d["last_historic_snapshot_time"]
    c.current_iteration = d["current_iteration"]
    c.n_historic_snapshots = d["n_historic_snapshots"]
    c.timestep_warning = d["timestep_warning"]
    try:
        c.coordinates_type = abi.COORDINATES[d["alternative_coordinates_type"]]
    except KeyError:
        raise InvalidCaseError("unknown coordinates type %r" % d["alternative_coordinates_type"])
    n = int(u["n_particles"])
    if not (1 <= n <= abi.MAX_PARTICLES):
        raise InvalidCaseError("n_particles = %d out of range" % n)
    c.n_particles = n
    ce = u["consider_effects"]
    c.consider_tides = int(ce["tides"])
    c.consider_rotational_flattening = int(ce["rotational_flattening"])
    c.consider_general_relativity = int(ce["general_relativity"])
    c.consider_disk = int(ce["disk"])
    c.consider_wind = int(ce["wind"])
    c.consider_evolution = int(ce["evolution"])
    if c.consider_disk:
        raise UnsupportedCaseError("disk interaction is outside the B200 hot path")
    c.general_relativity_implementation = abi.GR_IMPLEMENTATIONS[u["general_relativity_implementation"]]
    hi = u["hosts"]["index"]
    c.host_most_massive = hi["most_massive"]
    c.host_tides = hi["tides"]
    c.host_rotational_flattening = hi["rotational_flattening"]
    c.host_general_relativity = hi["general_relativity"]
    c.host_disk = hi["disk"]
    # HashMap<usize, f64> (universe.rs:61), serde_json writes the keys as strings; NaN = absent
    for k in range(abi.MAX_PARTICLES * abi.MAX_PARTICLES):
        c.pair_dependent_scaled_dissipation_factor[k] = float("nan")
    for key, value in (u.get("pair_dependent_scaled_dissipation_factor") or {}).items():
        k = int(key)
        if not (0 <= k < abi.MAX_PARTICLES * abi.MAX_PARTICLES):
            raise InvalidCaseError("pair_dependent_scaled_dissipation_factor key %r out of range" % (key,))
        c.pair_dependent_scaled_dissipation_factor[k] = float(value)
    tables = CaseTables()
    evolvers = u["particles_evolvers"]
    for i in range(n):
        p = u["particles"][i]
        b = c.bodies[i]
        b.id = p["id"]
        b.mass = p["mass"]
        b.mass_g = p["mass_g"]
        b.radius = p["radius"]
        b.radius_of_gyration_2 = p["radius_of_gyration_2"]
        b.moment_of_inertia = p["moment_of_inertia"]
        b.inertial_position[:] = _axes(p["inertial_position"])
        b.inertial_velocity[:] = _axes(p["inertial_velocity"])
        b.inertial_acceleration[:] = _axes(p["inertial_acceleration"])
        b.heliocentric_position[:] = _axes(p["heliocentric_position"])
        b.heliocentric_velocity[:] = _axes(p["heliocentric_velocity"])
        b.spin[:] = _axes(p["spin"])
        b.angular_momentum[:] = _axes(p["angular_momentum"])
        # tides
        role, model = _effect_role_and_payload(p["tides"]["effect"])
        b.tides_role = abi.ROLES[role]
        if model is not None:
            (mname, params), = model.items()
            if mname != "ConstantTimeLag":
                if c.consider_tides:
                    raise UnsupportedCaseError("tidal model %s is outside the B200 hot path (ConstantTimeLag only)" % mname)
                b.tides_role = abi.ROLE_DISABLED
            else:
                b.tides_dissipation_factor = params["dissipation_factor"]
                b.tides_dissipation_factor_scale = params["dissipation_factor_scale"]
                b.tides_love_number = params["love_number"]
        ti = p["tides"]["parameters"]["internal"]
        b.tides_scaled_dissipation_factor = ti["scaled_dissipation_factor"]
        b.tides_lag_angle = ti["lag_angle"]
        b.tides_denergy_dt = ti["denergy_dt"]
        # rotational flattening
        role, model = _effect_role_and_payload(p["rotational_flattening"]["effect"])
        b.flattening_role = abi.ROLES[role]
        if model is not None:
            (mname, params), = model.items()
            if mname != "OblateSpheroid":
                if c.consider_rotational_flattening:
                    raise UnsupportedCaseError("rotational flattening model %s is outside the B200 hot path (OblateSpheroid only)" % mname)
                b.flattening_role = abi.ROLE_DISABLED
            else:
                b.flattening_love_number = params["love_number"]
        # general relativity
        role, impl = _effect_role_and_payload(p["general_relativity"]["effect"])
        b.general_relativity_role = abi.ROLES[role]
        b.general_relativity_factor = p["general_relativity"]["parameters"]["internal"]["factor"]
        # wind (wind.rs:6-39); the disk must be inert, its description is carried for recovery images
        wrole, _ = _effect_role_and_payload(p["wind"]["effect"])
        b.wind_role = 0 if wrole == "Interaction" else 1
        b.wind_k_factor = p["wind"]["parameters"]["input"]["k_factor"]
        b.wind_rotation_saturation = p["wind"]["parameters"]["input"]["rotation_saturation"]
        drole, dprops = _effect_role_and_payload(p["disk"]["effect"])
        b.disk_role = abi.ROLES[drole]
        if dprops is not None:
            for k, key in enumerate(("inner_edge_distance", "outer_edge_distance", "lifetime", "alpha",
                                     "surface_density_normalization", "mean_molecular_weight")):
                b.disk_properties[k] = dprops[key]
        ref = p.get("reference", "MostMassiveParticle")
        b.reference = -1 if isinstance(ref, str) else int(ref["Particle"])
        # evolution
        etype, eparam = _effect_role_and_payload(p["evolution"])
        b.evolution_type = abi.EVOLUTION_TYPES[etype]
        b.evolution_parameter = float(eparam) if eparam is not None else 0.0
        b.evolution_table = -1
        ev = evolvers[i]
        b.evolution_left_index = ev.get("left_index", 0)
        if b.evolution_type != abi.EVO_NONEVOLVING and c.consider_evolution:
            b.evolution_table = tables.add(ev["time"], ev["radius"], ev["radius_of_gyration_2"], ev["love_number"],
                                           ev["inverse_tidal_q_factor"])
        c.inertial_velocity_errors[i][:] = _axes(d["inertial_velocity_errors"][i])
        c.particle_angular_momentum_errors[i][:] = _axes(d["particle_angular_momentum_errors"][i])
    for i in range(n, abi.MAX_PARTICLES):
        _set_dummy_body(c.bodies[i])
    rr = u["roche_radiuses"]
    for k in range(abi.MAX_PARTICLES * abi.MAX_PARTICLES):
        c.roche_radiuses[k] = rr[k]
    return c, tables


def load_case_json(path):
    with open(path) as f:
        return case_from_dict(json.load(f))


def copy_case(c):
    out = abi.Case()
    C.memmove(C.byref(out), C.byref(c), C.sizeof(abi.Case))
    return out


class NativeTables:
    """Evolution tables owned by the C++ reader (pb200_table_store_t); same interface as CaseTables."""

    def __init__(self, handle):
        self._h = handle

    def __len__(self):
        from ._lib import lib
        return lib().pb200_table_store_count(self._h)

    def as_ctypes(self):
        from ._lib import lib
        p = lib().pb200_table_store_tables(self._h)
        return p if len(self) else (abi.Table * 1)()

    def __del__(self):
        try:
            from ._lib import lib
            if self._h:
                lib().pb200_table_store_free(self._h)
                self._h = None
        except Exception:
            pass


def load_case_file(path):
    """pb200_case_load: the C++ reader of the library (JSON for *.json, bincode otherwise) — what the CLI uses."""
    from ._lib import lib, last_error
    c = abi.Case()
    h = C.c_void_p()
    rc = lib().pb200_case_load(str(path).encode(), C.byref(c), C.byref(h))
    if rc == abi.E_UNSUPPORTED:
        raise UnsupportedCaseError(last_error())
    if rc != abi.OK:
        raise InvalidCaseError(last_error())
    return c, NativeTables(h)


def save_case_file(path, case, tables):
    """pb200_case_save: bincode recovery snapshot (or pretty JSON for *.json) in the reference's layout."""
    from ._lib import lib, last_error
    rc = lib().pb200_case_save(str(path).encode(), C.byref(case), tables.as_ctypes(), len(tables))
    if rc != abi.OK:
        raise InvalidCaseError(last_error())
