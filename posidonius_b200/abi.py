"""ctypes mirror of include/posidonius_b200.h (field order and types must match the header)."""
import ctypes as C

MAX_PARTICLES = 10
HISTORIC_RECORD_BYTES = 156

OK, E_INVALID, E_UNSUPPORTED, E_CUDA, E_NOMEM = 0, -1, -2, -3, -4
ARITH_FAST, ARITH_STRICT, ARITH_HYBRID = 0, 1, 2

COORD_JACOBI, COORD_DEMOCRATIC_HELIOCENTRIC, COORD_WHDS = 0, 1, 2
COORDINATES = {"Jacobi": COORD_JACOBI, "DemocraticHeliocentric": COORD_DEMOCRATIC_HELIOCENTRIC, "WHDS": COORD_WHDS}
COORDINATES_INV = {v: k for k, v in COORDINATES.items()}

ROLE_CENTRAL, ROLE_ORBITING, ROLE_DISABLED = 0, 1, 2
ROLES = {"CentralBody": ROLE_CENTRAL, "OrbitingBody": ROLE_ORBITING, "Disabled": ROLE_DISABLED}
ROLES_INV = {v: k for k, v in ROLES.items()}

GR_KIDDER1995, GR_ANDERSON1975, GR_NEWHALL1983, GR_DISABLED = 0, 1, 2, 3
GR_IMPLEMENTATIONS = {"Kidder1995": 0, "Anderson1975": 1, "Newhall1983": 2, "Disabled": 3}
GR_IMPLEMENTATIONS_INV = {v: k for k, v in GR_IMPLEMENTATIONS.items()}

EVOLUTION_TYPES = {
    "GalletBolmont2017": 0,
    "BolmontMathis2016": 1,
    "Baraffe2015": 2,
    "Leconte2011": 3,
    "Baraffe1998": 4,
    "LeconteChabrier2013": 5,
    "NonEvolving": 6,
}
EVOLUTION_TYPES_INV = {v: k for k, v in EVOLUTION_TYPES.items()}
EVO_NONEVOLVING = 6

STATUS_OK, STATUS_COMPLETED, STATUS_ROCHE_DESTROYED, STATUS_COLLISION, STATUS_EJECTED, STATUS_ZERO_INERTIA = range(6)
WARN_MIDPOINT_NOT_CONVERGED, WARN_TIMESTEP_GT_PERIOD, WARN_HISTORY_DROPPED = 1, 2, 4

_d3 = C.c_double * 3


class Body(C.Structure):
    _fields_ = [
        ("mass", C.c_double),
        ("mass_g", C.c_double),
        ("radius", C.c_double),
        ("radius_of_gyration_2", C.c_double),
        ("moment_of_inertia", C.c_double),
        ("inertial_position", _d3),
        ("inertial_velocity", _d3),
        ("inertial_acceleration", _d3),
        ("heliocentric_position", _d3),
        ("heliocentric_velocity", _d3),
        ("spin", _d3),
        ("angular_momentum", _d3),
        ("tides_dissipation_factor", C.c_double),
        ("tides_dissipation_factor_scale", C.c_double),
        ("tides_love_number", C.c_double),
        ("tides_scaled_dissipation_factor", C.c_double),
        ("tides_lag_angle", C.c_double),
        ("tides_denergy_dt", C.c_double),
        ("flattening_love_number", C.c_double),
        ("general_relativity_factor", C.c_double),
        ("evolution_parameter", C.c_double),
        ("tides_role", C.c_int32),
        ("flattening_role", C.c_int32),
        ("general_relativity_role", C.c_int32),
        ("evolution_type", C.c_int32),
        ("evolution_table", C.c_int32),
        ("evolution_left_index", C.c_int32),
        ("id", C.c_int32),
        ("reference", C.c_int32),
        ("wind_role", C.c_int32),
        ("disk_role", C.c_int32),
        ("wind_k_factor", C.c_double),
        ("wind_rotation_saturation", C.c_double),
        ("disk_properties", C.c_double * 6),
    ]


class Case(C.Structure):
    _fields_ = [
        ("time_step", C.c_double),
        ("half_time_step", C.c_double),
        ("initial_time", C.c_double),
        ("time_limit", C.c_double),
        ("current_time", C.c_double),
        ("recovery_snapshot_period", C.c_double),
        ("historic_snapshot_period", C.c_double),
        ("last_recovery_snapshot_time", C.c_double),
        ("last_historic_snapshot_time", C.c_double),
        ("current_iteration", C.c_uint64),
        ("n_historic_snapshots", C.c_uint64),
        ("timestep_warning", C.c_uint64),
        ("coordinates_type", C.c_int32),
        ("n_particles", C.c_int32),
        ("consider_tides", C.c_int32),
        ("consider_rotational_flattening", C.c_int32),
        ("consider_general_relativity", C.c_int32),
        ("consider_disk", C.c_int32),
        ("consider_wind", C.c_int32),
        ("consider_evolution", C.c_int32),
        ("general_relativity_implementation", C.c_int32),
        ("host_most_massive", C.c_int32),
        ("host_tides", C.c_int32),
        ("host_rotational_flattening", C.c_int32),
        ("host_general_relativity", C.c_int32),
        ("host_disk", C.c_int32),
        ("bodies", Body * MAX_PARTICLES),
        ("inertial_velocity_errors", _d3 * MAX_PARTICLES),
        ("particle_angular_momentum_errors", _d3 * MAX_PARTICLES),
        ("roche_radiuses", C.c_double * (MAX_PARTICLES * MAX_PARTICLES)),
        ("pair_dependent_scaled_dissipation_factor", C.c_double * (MAX_PARTICLES * MAX_PARTICLES)),
    ]


class Table(C.Structure):
    _fields_ = [
        ("n_rows", C.c_size_t),
        ("time", C.POINTER(C.c_double)),
        ("radius", C.POINTER(C.c_double)),
        ("radius_of_gyration_2", C.POINTER(C.c_double)),
        ("love_number", C.POINTER(C.c_double)),
        ("inverse_tidal_q_factor", C.POINTER(C.c_double)),
    ]


class StateView(C.Structure):
    _fields_ = [
        ("position", C.POINTER(C.c_double)),
        ("velocity", C.POINTER(C.c_double)),
        ("acceleration", C.POINTER(C.c_double)),
        ("angular_momentum", C.POINTER(C.c_double)),
        ("spin", C.POINTER(C.c_double)),
        ("velocity_errors", C.POINTER(C.c_double)),
        ("angular_momentum_errors", C.POINTER(C.c_double)),
        ("radius", C.POINTER(C.c_double)),
        ("radius_of_gyration_2", C.POINTER(C.c_double)),
        ("moment_of_inertia", C.POINTER(C.c_double)),
        ("current_time", C.POINTER(C.c_double)),
    ]


STATE_FIELDS_VEC = ("position", "velocity", "acceleration", "angular_momentum", "spin", "velocity_errors",
                    "angular_momentum_errors")
STATE_FIELDS_BODY = ("radius", "radius_of_gyration_2", "moment_of_inertia")
STATE_FIELDS_SYS = ("current_time",)
