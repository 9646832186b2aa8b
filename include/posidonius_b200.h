/*
 * posidonius_b200.h — C ABI of the B200 ensemble integrator.
 *
 * This is the drop-in boundary for ONE path of marblestation/posidonius: the
 * WHFast kick-drift-kick step with constant-time-lag tides, oblate-spheroid
 * rotational flattening, general relativity (Kidder1995 / Anderson1975 /
 * Newhall1983), the implicit-midpoint spin update and evolution-table
 * interpolation, run over an ENSEMBLE of independent systems on one GPU.
 *
 * The reference has no FFI; its boundary is `trait Integrator`
 * (src/integrator/mod.rs:16-26) implemented by `WHFast`
 * (src/integrator/whfast.rs:169-318) over the serde image of the `WHFast`
 * struct (src/integrator/whfast.rs:98-120).  Every entry point below names the
 * reference item it replaces.  Plain pointers and sizes only; no exceptions
 * cross this boundary; all functions return 0 on success or a negative
 * PB200_E_* code (message via pb200_last_error()).
 *
 * Layout conventions
 *   - pb200_case_t is the flattened image of one `WHFast` value (one system).
 *   - Ensemble state moves as SoA: field-major, then body, then system:
 *       a[(c * n_bodies + b) * n_systems + s]   for 3-vectors (c = x,y,z)
 *       a[b * n_systems + s]                    for per-body scalars
 *       a[s]                                    for per-system scalars
 */
#ifndef POSIDONIUS_B200_H
#define POSIDONIUS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB200_MAX_PARTICLES 10 /* src/constants.rs:3 */
#define PB200_HISTORIC_RECORD_BYTES 156 /* src/integrator/output.rs:85-89 */

/* error codes */
#define PB200_OK 0
#define PB200_E_INVALID (-1)      /* bad argument / inconsistent case */
#define PB200_E_UNSUPPORTED (-2)  /* effect/integrator outside the hot path (Kaula, creep, disk, IAS15, LeapFrog...) */
#define PB200_E_CUDA (-3)         /* CUDA runtime failure */
#define PB200_E_NOMEM (-4)

/* CoordinatesType, enum tag order of src/integrator/whfast.rs:91-95 */
#define PB200_COORD_JACOBI 0
#define PB200_COORD_DEMOCRATIC_HELIOCENTRIC 1
#define PB200_COORD_WHDS 2

/* TidesEffect / RotationalFlatteningEffect / GeneralRelativityEffect tag order
 * (tides/common.rs:89-93, rotational_flattening/common.rs:51-55, general_relativity.rs:48-52) */
#define PB200_ROLE_CENTRAL 0
#define PB200_ROLE_ORBITING 1
#define PB200_ROLE_DISABLED 2

/* GeneralRelativityImplementation tag order (general_relativity.rs:40-45) */
#define PB200_GR_KIDDER1995 0
#define PB200_GR_ANDERSON1975 1
#define PB200_GR_NEWHALL1983 2
#define PB200_GR_DISABLED 3

/* EvolutionType tag order (effects/evolution.rs:9-17) */
#define PB200_EVO_GALLETBOLMONT2017 0
#define PB200_EVO_BOLMONTMATHIS2016 1
#define PB200_EVO_BARAFFE2015 2
#define PB200_EVO_LECONTE2011 3
#define PB200_EVO_BARAFFE1998 4
#define PB200_EVO_LECONTECHABRIER2013 5
#define PB200_EVO_NONEVOLVING 6

/* per-system status word (replaces the reference's panic!/warnings) */
#define PB200_STATUS_OK 0
#define PB200_STATUS_COMPLETED 1          /* t + dt > time_limit, whfast.rs:300 */
#define PB200_STATUS_ROCHE_DESTROYED 2    /* universe.rs:225-228 */
#define PB200_STATUS_COLLISION 3          /* universe.rs:230-233 */
#define PB200_STATUS_EJECTED 4            /* universe.rs:234-237 */
#define PB200_STATUS_ZERO_INERTIA 5       /* particles/common.rs:5-7 */
/* warning bits, OR-ed into pb200 "warnings" word (do not stop the system) */
#define PB200_WARN_MIDPOINT_NOT_CONVERGED 1u /* whfast.rs:389-391 */
#define PB200_WARN_TIMESTEP_GT_PERIOD 2u     /* whfast.rs:702-707 */
#define PB200_WARN_HISTORY_DROPPED 4u        /* a historic snapshot found the device history buffer full (not a reference condition) */

/* One body: the fields of `Particle` (src/particles/particle.rs:16-52) that the
 * hot path reads or carries between steps.  Per-effect scratch that the
 * reference recomputes before every use is not part of the image. */
typedef struct pb200_body {
    double mass;
    double mass_g;
    double radius;
    double radius_of_gyration_2;
    double moment_of_inertia;
    double inertial_position[3];
    double inertial_velocity[3];
    double inertial_acceleration[3]; /* Newtonian acc. left by the last gravity evaluation; read by Anderson/Newhall GR */
    double heliocentric_position[3];
    double heliocentric_velocity[3]; /* host's value is read stale by universe.rs:335-337 */
    double spin[3];                  /* spin of the previous evaluation feeds r.omega (universe.rs:429-430 order) */
    double angular_momentum[3];
    /* tides: ConstantTimeLagParameters (constant_time_lag.rs:12-18) + internal.scaled_dissipation_factor */
    double tides_dissipation_factor;
    double tides_dissipation_factor_scale;
    double tides_love_number;
    double tides_scaled_dissipation_factor;
    double tides_lag_angle;
    double tides_denergy_dt;
    /* rotational flattening: OblateSpheroidParameters (oblate_spheroid.rs:8-10) */
    double flattening_love_number;
    /* general relativity: internal.factor (general_relativity.rs:16) */
    double general_relativity_factor;
    /* EvolutionType payload: stellar mass (f64 variants) or 0/1 (LeconteChabrier2013(bool)) */
    double evolution_parameter;
    int32_t tides_role;       /* PB200_ROLE_* */
    int32_t flattening_role;  /* PB200_ROLE_* */
    int32_t general_relativity_role; /* PB200_ROLE_* */
    int32_t evolution_type;   /* PB200_EVO_* */
    int32_t evolution_table;  /* index into the table pool, -1 when NonEvolving */
    int32_t evolution_left_index; /* Evolver.left_index cursor (evolution.rs:27), carried for recovery images */
    int32_t id;
    /* Reference (particle.rs:9-13): -1 = MostMassiveParticle, k >= 0 = Particle(k); wind (wind.rs:6-39) role and input
     * parameters (dL/dt of wind.rs:72-91 is part of the hot path); the disk (disk.rs:6-56) role and parameters are inert
     * on the GPU path, carried so that a recovery image round-trips through the reference's own reader. */
    int32_t reference;
    int32_t wind_role;        /* 0 = Interaction, 1 = Disabled (WindEffect tag order) */
    int32_t disk_role;        /* PB200_ROLE_* (DiskEffect tag order) */
    double wind_k_factor;
    double wind_rotation_saturation;
    double disk_properties[6]; /* DiskProperties, only meaningful when disk_role == PB200_ROLE_CENTRAL */
} pb200_body_t;

/* One system: flattened `WHFast` (whfast.rs:98-120) + `Universe` (universe.rs:50-63). */
typedef struct pb200_case {
    double time_step;
    double half_time_step;
    double initial_time;
    double time_limit;
    double current_time;
    double recovery_snapshot_period;
    double historic_snapshot_period;
    double last_recovery_snapshot_time;
    double last_historic_snapshot_time;
    uint64_t current_iteration;
    uint64_t n_historic_snapshots;
    uint64_t timestep_warning;
    int32_t coordinates_type;   /* PB200_COORD_* */
    int32_t n_particles;
    /* ConsiderEffects (universe.rs:40-47) */
    int32_t consider_tides;
    int32_t consider_rotational_flattening;
    int32_t consider_general_relativity;
    int32_t consider_disk;
    int32_t consider_wind;
    int32_t consider_evolution;
    int32_t general_relativity_implementation; /* PB200_GR_* */
    /* HostIndices (universe.rs:16-22); MAX_PARTICLES+1 when absent */
    int32_t host_most_massive;
    int32_t host_tides;
    int32_t host_rotational_flattening;
    int32_t host_general_relativity;
    int32_t host_disk;
    pb200_body_t bodies[PB200_MAX_PARTICLES];
    double inertial_velocity_errors[PB200_MAX_PARTICLES][3];         /* whfast.rs:117 */
    double particle_angular_momentum_errors[PB200_MAX_PARTICLES][3]; /* whfast.rs:119 */
    double roche_radiuses[PB200_MAX_PARTICLES * PB200_MAX_PARTICLES]; /* universe.rs:62, row stride = n_particles */
    /* Universe.pair_dependent_scaled_dissipation_factor (universe.rs:61; tides/constant_time_lag.rs:152-165): the
     * HashMap<usize, f64> keyed id * MAX_PARTICLES + depends_on_id, flattened; NaN = key absent. Only bodies whose
     * EvolutionType drives dynamical tides (BolmontMathis2016, GalletBolmont2017, LeconteChabrier2013(true)) use it. */
    double pair_dependent_scaled_dissipation_factor[PB200_MAX_PARTICLES * PB200_MAX_PARTICLES];
} pb200_case_t;

/* One evolution table = the vectors of `Evolver` (evolution.rs:19-28). Columns
 * that the EvolutionType does not use may be NULL. `time` is in days since
 * initial_time, as stored in the JSON. */
typedef struct pb200_table {
    size_t n_rows;
    const double* time;
    const double* radius;
    const double* radius_of_gyration_2;
    const double* love_number;
    const double* inverse_tidal_q_factor;
} pb200_table_t;

typedef struct pb200_ensemble pb200_ensemble_t;

/* Library/version probes. */
const char* pb200_version(void);
const char* pb200_last_error(void);
/* Number of CUDA devices visible; <0 on CUDA error. */
int pb200_device_count(void);

/* Validates one case against the hot-path scope. Returns PB200_OK,
 * PB200_E_UNSUPPORTED or PB200_E_INVALID. Mirrors the trial deserialisation +
 * checks of output::restore_snapshot (output.rs:206-231) and
 * Universe::new (universe.rs:73-175) — no CPU fallback exists behind it. */
int pb200_case_validate(const pb200_case_t* c, const pb200_table_t* tables, size_t n_tables);

/* ---- case files (host side, no GPU needed) -------------------------------------------------------------------
 * pb200_case_load replaces output::restore_snapshot's readers (output.rs:206-307): a path ending in ".json" is
 * parsed as the serde_json image of `WHFast` (what posidonius' Python package writes and what
 * write_recovery_snapshot writes for .json paths), anything else as the bincode 1.3.3 image. The WHFast image is
 * the only one accepted: Ias15 / LeapFrog images and cases with effects outside the hot path return
 * PB200_E_UNSUPPORTED. The evolution tables are owned by the returned store.
 * pb200_case_save replaces write_recovery_snapshot (output.rs:56-82): bincode unless the path ends in ".json";
 * an existing file is first renamed to a 12-hourly backup like the reference does (output.rs:63-69). */
typedef struct pb200_table_store pb200_table_store_t;
int pb200_case_load(const char* path, pb200_case_t* out, pb200_table_store_t** tables_out);
int pb200_case_save(const char* path, const pb200_case_t* c, const pb200_table_t* tables, size_t n_tables);
const pb200_table_t* pb200_table_store_tables(const pb200_table_store_t* s);
size_t pb200_table_store_count(const pb200_table_store_t* s);
void pb200_table_store_free(pb200_table_store_t* s);

/* Replaces output::restore_snapshot + Universe ownership (output.rs:206-231):
 * builds a device-resident ensemble of n_systems systems on `device`.
 * n_cases must be 1 (the case is replicated) or n_systems. All cases must
 * share structure (n_particles, coordinates, effects, roles, tables, dt). */
int pb200_ensemble_create(const pb200_case_t* cases, size_t n_cases, size_t n_systems,
                          const pb200_table_t* tables, size_t n_tables, int device,
                          pb200_ensemble_t** out);
void pb200_ensemble_destroy(pb200_ensemble_t* e);

/* Synthetic perturbed ensemble built ON THE DEVICE (no per-member case images on the host): member 0 is
 * `base`; member k > 0 has every non-host body's heliocentric position and velocity multiplied component-wise by
 * (1 + delta), delta uniform in (-amplitude, amplitude) from a SplitMix64 stream seeded with seed * 0x100000001b3 + k
 * (drawn in the order [body][x, y, z, vx, vy, vz]), and the barycentric coordinates recomputed as Universe::new does
 * (universe.rs:95-105, 663-697). Masses, radii, spins and parameters are those of `base`. Bit-identical to building the
 * members on the host with the same recipe (posidonius-b200 ensemble, posidonius_b200/perturb.py::splitmix_cases). */
int pb200_ensemble_create_perturbed(const pb200_case_t* base, size_t n_systems, uint64_t seed, double amplitude,
                                    const pb200_table_t* tables, size_t n_tables, int device, pb200_ensemble_t** out);
/* The same for one SHARD of that ensemble: members first_member .. first_member + n_systems - 1 of the global ensemble that
 * pb200_ensemble_create_perturbed(seed) defines (one process per GPU, each building its contiguous range — the reference
 * runs one Universe per process, src/main.rs:124-176). Member k's stream depends on (seed, k) only, so the union of the
 * shards is the unsharded ensemble bit for bit. */
int pb200_ensemble_create_perturbed_range(const pb200_case_t* base, uint64_t first_member, size_t n_systems, uint64_t seed,
                                          double amplitude, const pb200_table_t* tables, size_t n_tables, int device,
                                          pb200_ensemble_t** out);

/* Integrator::get_n_particles / get_current_time / get_n_historic_snapshots (mod.rs:18-20). */
int pb200_ensemble_n_particles(const pb200_ensemble_t* e);
size_t pb200_ensemble_n_systems(const pb200_ensemble_t* e);
/* Integrator::set_time_limit / set_snapshot_periods (whfast.rs:187-224); applied to every system. */
int pb200_ensemble_set_time_limit(pb200_ensemble_t* e, double time_limit);
int pb200_ensemble_set_snapshot_periods(pb200_ensemble_t* e, double historic_snapshot_period,
                                        double recovery_snapshot_period);

/* Arithmetic of the perturbation forces (tides, flattening, GR). The WHFast core is always strict IEEE in the
 * reference's operation order.
 *   PB200_ARITH_HYBRID (default): the implicit midpoint (whfast.rs:322-466) evaluates the forces at least three times per
 *     half step and commits the increments of the last evaluation only. The first two evaluations use the fast forces, the
 *     committed one the exact forces, so v, L and their Kahan residuals carry the reference's roundings:
 *     within 1e-10 relative of the reference after 10^4 steps on every configuration, most members bit for bit.
 *   PB200_ARITH_STRICT: every evaluation with the exact forces, operation by operation as the reference writes them; the
 *     whole step is bit-reproducible against the CPU restatement of the reference (every GR variant, any particle order).
 *   PB200_ARITH_FAST: every evaluation with the fast forces (FMA contraction, reciprocal reuse, hoisted constants); agrees
 *     with the reference to roundoff-growth level only (1e-10 .. 4e-10 after 10^4 steps). */
#define PB200_ARITH_FAST 0
#define PB200_ARITH_STRICT 1
#define PB200_ARITH_HYBRID 2
int pb200_ensemble_set_arithmetic(pb200_ensemble_t* e, int mode);

/* Integrator::initialize_physical_values (whfast.rs:226-233): spin = L/I, evolving
 * quantities at t = 0, Roche radii. Fails with PB200_E_INVALID on a resumed ensemble
 * (current_time != 0), like the reference's panic. */
int pb200_ensemble_initialize_physical_values(pb200_ensemble_t* e);

/* Integrator::iterate called n_steps times on every live system (whfast.rs:235-305).
 * Systems stop individually when completed (t + dt > time_limit) or on a physical
 * failure (status word). Historic snapshot records that fall due inside the call are
 * appended to the ensemble's device-side history buffer (see pb200_ensemble_history_*).
 * Asynchronous on the ensemble's stream; pb200_ensemble_synchronize() waits. */
int pb200_ensemble_step(pb200_ensemble_t* e, uint64_t n_steps);
int pb200_ensemble_synchronize(pb200_ensemble_t* e);
/* Duration in ms of the last pb200_ensemble_step kernel(s), measured with CUDA events
 * on the ensemble's stream. Synchronizes. */
int pb200_ensemble_last_step_ms(pb200_ensemble_t* e, float* ms);
/* Kernel launches issued by this ensemble so far. */
uint64_t pb200_ensemble_launch_count(const pb200_ensemble_t* e);
/* Time slices of the last step launch (1 = every block of systems ran its steps in one piece). */
unsigned pb200_ensemble_last_pieces(const pb200_ensemble_t* e);
/* Name of the step-kernel build that ran the last pb200_ensemble_step ("" before the first one): "generic" (lane = body, any
 * geometry), "n8" / "n8w" (8 bodies, 64- / 384-thread CTAs), "s2", "s2t", "s3", "s3e", "s3j", "s3p" (lane = planet, compile-time
 * effect sets), "s2any", "s3any", "s3jany" (lane = planet, effect set read at run time). Diagnostics: every build computes the
 * same step. The string is static storage. */
const char* pb200_ensemble_last_kernel(const pb200_ensemble_t* e);
/* The same name for a case that has not been uploaded: which build would integrate an ensemble of n_systems copies /
 * perturbed members of `c` on a GPU with sm_count SMs in the given arithmetic. Needs no device. "" for a null / malformed case. */
const char* pb200_case_step_kernel(const pb200_case_t* c, size_t n_systems, int sm_count, int arithmetic);

/* Per-system status (PB200_STATUS_*), warning bits and the iteration index of the event. */
int pb200_ensemble_status(pb200_ensemble_t* e, int32_t* status, uint32_t* warnings,
                          uint64_t* iteration_of_event);

/* SoA download/upload of the dynamic state (layout at top). Any pointer may be NULL.
 *   position, velocity, acceleration, angular_momentum, spin, velocity_errors,
 *   angular_momentum_errors : 3 * n_bodies * n_systems
 *   radius, radius_of_gyration_2, moment_of_inertia : n_bodies * n_systems
 *   current_time : n_systems */
typedef struct pb200_state_view {
    double* position;
    double* velocity;
    double* acceleration;
    double* angular_momentum;
    double* spin;
    double* velocity_errors;
    double* angular_momentum_errors;
    double* radius;
    double* radius_of_gyration_2;
    double* moment_of_inertia;
    double* current_time;
} pb200_state_view_t;
int pb200_ensemble_download(pb200_ensemble_t* e, const pb200_state_view_t* dst);
int pb200_ensemble_upload(pb200_ensemble_t* e, const pb200_state_view_t* src);
/* Full image of system `s` as a pb200_case_t (what write_recovery_snapshot serialises, whfast.rs:307-316). */
int pb200_ensemble_get_case(pb200_ensemble_t* e, size_t s, pb200_case_t* out);

/* Historic snapshots (output.rs:119-163; whfast.rs:237-261). Records are produced on
 * the device in the reference's 156-byte little-endian layout, one per body per
 * snapshot. `pb200_ensemble_history_pending` = snapshots (per system) buffered since
 * the last drain; `pb200_ensemble_history_drain` copies them system-major:
 * dst[((s * n_snapshots + k) * n_bodies + b) * 156 ...] and empties the buffer. */
size_t pb200_ensemble_history_pending(pb200_ensemble_t* e);
/* Snapshots per system that the device-side history buffer holds between drains: pb200_ensemble_step refuses a call that
 * could overflow it, so callers bound their calls by (capacity - pending) snapshot periods. */
size_t pb200_ensemble_history_capacity(const pb200_ensemble_t* e);
int pb200_ensemble_history_drain(pb200_ensemble_t* e, void* dst, size_t dst_bytes);

/* Diagnostics with the formulas of Universe::compute_total_energy /
 * compute_total_angular_momentum (universe.rs:625-658), evaluated on the device after
 * a heliocentric refresh. energy / angular_momentum: n_systems each. */
int pb200_ensemble_summary(pb200_ensemble_t* e, double* energy, double* angular_momentum);

/* End-to-end convenience for benchmarking the boundary with HOST buffers: uploads the
 * SoA state in `io`, advances n_steps, downloads the state back into `io`. */
int pb200_ensemble_run_host(pb200_ensemble_t* e, const pb200_state_view_t* io, uint64_t n_steps);

/* DFMA-chain microbenchmark used as the FP64 roofline denominator: returns the
 * sustained fp64 FLOP/s (FMA = 2) measured over `ms_target` milliseconds. */
int pb200_measure_fp64_peak(int device, double ms_target, double* flops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* POSIDONIUS_B200_H */
