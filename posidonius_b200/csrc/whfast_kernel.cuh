// whfast_kernel.cuh — the ensemble WHFast step for sm_100a (fp64 vector pipe, no tensor cores).
//
// Mapping: one body per lane, one system per aligned group of W lanes (W = 2,4,8,16), 32/W systems
// per warp. Dynamic state lives in registers (and per-thread shared-memory columns, `Cold`) for all
// the steps a CTA runs; the lanes of a group exchange values through those columns (one LDS.64 per
// double after a __syncwarp) — the host body's quantities, the terms of the ordered sums, the
// contributions to the host sums; warp shuffles remain in the cold paths and the GR variants.
// HBM is touched only when a CTA starts and ends (coalesced SoA [field][body][system]) and when a
// historic snapshot falls due.
//
// Reference path restated here (file:line under /root/reference/src):
//   WHFast::iterate                          integrator/whfast.rs:235-305
//   integrate_velocity_dependent_forces      integrator/whfast.rs:322-466
//   kepler_individual_step / stumpff         integrator/whfast.rs:676-876
//   coordinate transforms, jump, kick        integrator/whfast.rs:495-672, 881-1155
//   gravity_calculate_acceleration           particles/universe.rs:198-303
//   inertial_to_heliocentric                 particles/universe.rs:318-351
//   calculate_additional_effects             particles/universe.rs:428-614
//   constant-time-lag tides                  effects/tides/{common,constant_time_lag}.rs
//   oblate-spheroid flattening               effects/rotational_flattening/{common,oblate_spheroid}.rs
//   GR Kidder1995/Anderson1975/Newhall1983   effects/general_relativity.rs:177-895
//   evolution interpolation                  effects/evolution.rs:449-567, tools.rs:840-907
//
// Arithmetic: the WHFast core (transforms, Kepler drift, jump, kick, gravity, compensated v/L updates) is
// strict IEEE in the reference's association order (strict.cuh) and reproduces the CPU oracle's roundings;
// the perturbation forces use FMA contraction, reciprocal reuse, hoisted powers and a transposed reduction.
// Parity is asserted against the CPU oracle at 1e-10 relative after 10^4 steps.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/posidonius_b200.h"
#include "strict.cuh"

namespace pb200 {

// ---- constants (constants.rs), same expressions as the reference
#define PB_HOUR 3600.
#define PB_DAY (24. * PB_HOUR)
#define PB_M_SUN 1.9818e30
#define PB_G_SI 6.67428e-11
#define PB_AU 1.49597870700e11
__device__ constexpr double kG = PB_G_SI / (PB_AU * PB_AU * PB_AU) * PB_M_SUN * (PB_DAY * PB_DAY);
__device__ constexpr double kK2 = kG;
__device__ constexpr double kC = (2.99792458e8 / PB_AU) * PB_DAY;
__device__ constexpr double kC2 = kC * kC;
__device__ constexpr double kInvC2 = 1.0 / kC2;
__device__ constexpr double kEps2 = 2.2204460492503131e-16 * 2.2204460492503131e-16;
__device__ constexpr double kMaxDistance2 = 100. * 100.;
__device__ constexpr double kPi = 3.14159265358979323846264338327950288;
__device__ constexpr double kRSun = 6.957e8 / PB_AU;                      // constants.rs:52
__device__ constexpr double kSunDynFreq2 = kK2 / (kRSun * kRSun * kRSun);  // constants.rs:66
__device__ constexpr double kSmoothDynTide = 1.0e-5 * 86400.;              // constants.rs:67

// FLAG_WIND: some body has WindEffect::Interaction (wind.rs:72-91). FLAG_DYN: some body's EvolutionType drives dynamical
// tides, i.e. pair-dependent dissipation factors (tides/constant_time_lag.rs:20-165). Both live in the run-time
// geometry build of the kernel only (PB_FIXED_N == 0).
enum : int { FLAG_TIDES = 1, FLAG_FLAT = 2, FLAG_GR = 4, FLAG_EVO = 8, FLAG_WIND = 16, FLAG_DYN = 32 };

// Device-side description of one evolution table (effects/evolution.rs:19-28).
struct DevTable {
    const double* time;
    const double* radius;
    const double* rg2;
    int n_rows;
    int interp_radius;  // 1 unless NonEvolving
    int reserved;
    const double* qinv; // inverse tidal Q factor column, null when the table has none
    // (which columns a BODY interpolates depends on its own EvolutionType: KParams::evo_rg2 / dyn_evo bit masks; a table
    // shared by bodies of different types carries every column it was given)
};

// Kernel parameters: everything uniform across the ensemble + SoA pointers.
struct KParams {
    int n_sys, n_bodies, W, shift;
    int host;      // index of the most massive body = host of every enabled effect
    int flags;     // FLAG_*
    int spin_on;   // integrate_spin (whfast.rs:270-275)
    double dt, half_dt, time_limit, hist_period;
    // per-body roles, uniform across systems (bit b set = body b is OrbitingBody for the effect)
    uint32_t tides_orbiting, flat_orbiting, gr_orbiting, gr_enabled /* != Disabled */;
    int tides_host_central, flat_host_central;
    int evo_table[PB200_MAX_PARTICLES];  // -1 = NonEvolving
    uint32_t evo_rg2;                    // bit b: body b's EvolutionType interpolates the radius of gyration (Baraffe2015 / Leconte2011 / LeconteChabrier2013)
    DevTable tables[PB200_MAX_PARTICLES];
    // dynamic state, [c][b][s]
    double *pos, *vel, *acc, *L, *spin, *verr, *lerr;
    // per body per system, [b][s]
    double *radius, *rg2, *moi;
    const double *mass, *mass_g, *sigma, *k2t, *k2f;
    const double* roche;  // [i][j][s], i,j < n_bodies
    // per system
    double *t, *last_hist;
    unsigned long long *iteration, *n_hist, *event_iteration, *tswarn;
    int* status;
    unsigned int* warnings;
    // historic snapshot planes: hist[slot][field(17)][b][s] doubles; slot = snapshots since launch start
    double* hist;
    int hist_capacity;          // slots available
    int* hist_count;            // per system: slots used
    // stale tidal internals for denergy_dt (tides/common.rs:263-279 reads the last evaluation), [k(13)][b][s]
    double* tide_scratch;
    // wind (wind.rs) and dynamical tides (constant_time_lag.rs:20-165); arrays are null unless the flag is set
    uint32_t wind_on;   // bit b: WindEffect::Interaction
    uint32_t dyn_evo;   // bit b: the body's EvolutionType drives dynamical tides
    const double *wind_k, *wind_sat;        // [b][s]
    const double *diss, *diss_scale;        // [b][s] ConstantTimeLagParameters.dissipation_factor(_scale)
    double* lag;                            // [b][s] tides.parameters.internal.lag_angle (evolution.rs:548-567)
    double *pair_h, *pair_p;                // [b][s] map entries (host, b) and (b, host); NaN = absent
    // Time slicing of a launch (whfast_step.cuh): the n_steps of a call are cut into n_pieces consecutive pieces per block of
    // systems ("group"); a CTA takes a ticket, runs one piece of one group and hands the state over through HBM.
    // sched[0] = ticket counter, sched[1 + g] = pieces of group g completed. n_pieces = 1: plain one-CTA-per-group launch.
    unsigned int* sched;
    unsigned int n_groups, n_pieces;
};

// Loads of state that another CTA of the same launch may have written (time slicing): L2, never a stale L1 line.
template <class T> __device__ __forceinline__ T ldm(const T* p) { return __ldcg(p); }

#define FULL 0xffffffffu

// Geometry accessors: PB_FIXED_N > 0 compiles the kernel for exactly that many bodies with the host at index 0
// (loops over bodies unroll, shuffle lanes become immediates); 0 = any geometry at run time. Set per inclusion of the
// device code in pb200_api.cu.
#define PB_N(P) (PB_FIXED_N ? PB_FIXED_N : (P).n_bodies)
#define PB_HOST(P) (PB_FIXED_N ? 0 : (P).host)
#define PB_W(P) (PB_FIXED_N ? PB_FIXED_W : (P).W)
#define PB_SHIFT(P) (PB_FIXED_N ? PB_FIXED_SHIFT : (P).shift)
// the fixed-geometry build also fixes the effect set (tides + flattening + GR, no evolution: the TRAPPIST-1 case)
#define PB_FLAGS(P) (PB_FIXED_N ? PB_FIXED_FLAGS : (P).flags)
#define PB_SPIN(P) (PB_FIXED_N ? 1 : (P).spin_on)

__device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(FULL, v, src); }
__device__ __forceinline__ double shfl_xor(double v, int m) { return __shfl_xor_sync(FULL, v, m); }

// all-reduce (sum) inside the aligned group of W lanes
__device__ __forceinline__ double group_sum(double v, int W) {
    for (int off = W >> 1; off > 0; off >>= 1) v += shfl_xor(v, off);
    return v;
}

struct V3 { double x, y, z; };
__device__ __forceinline__ V3 v3(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(double s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ V3 shfl3(V3 a, int src) { return v3(shfl(a.x, src), shfl(a.y, src), shfl(a.z, src)); }

// ---------------------------------------------------------------------------------------------
// Stumpff functions c0..c3 (whfast.rs:844-876) and Stiefel G-functions (whfast.rs:835-842), strict arithmetic
// 1/13!, 1/12!, ..., 1/2! for the series below. Deliberately NOT const: as constant-bank operands the coefficients cost
// nothing (DADD R, -R, c[3][..]); as literals each one is materialised by two UMOVs before every use.
static __constant__ double kInvFactorial[12] = {1. / 6227020800., 1. / 479001600., 1. / 39916800., 1. / 3628800., 1. / 362880., 1. / 40320.,
                                         1. / 5040., 1. / 720., 1. / 120., 1. / 24., 1. / 6., 1. / 2.};
__device__ __forceinline__ void stumpff_cs3(sd z, sd& c0, sd& c1, sd& c2, sd& c3) {
    int n = 0;
    // z / 4 (exact scaling, same value). Bounded (deviation D3): the reference's loop never ends for an infinite z
    // (a blown-up state) and a kernel must not hang; finite doubles need at most 515 trips.
    while (fabs(z.v) > 0.1 && n < 600) { z = z * sd(0.25); n++; }
    const double* F = kInvFactorial;
    sd c_odd = sd(F[0]);    // 1/13!
    sd c_even = sd(F[1]);   // 1/12!
    c_odd = sd(F[2]) - z * c_odd;    c_even = sd(F[3]) - z * c_even;   // 11!, 10!
    c_odd = sd(F[4]) - z * c_odd;    c_even = sd(F[5]) - z * c_even;   // 9!, 8!
    c_odd = sd(F[6]) - z * c_odd;    c_even = sd(F[7]) - z * c_even;   // 7!, 6!
    c_odd = sd(F[8]) - z * c_odd;    c_even = sd(F[9]) - z * c_even;   // 5!, 4!
    c_odd = sd(F[10]) - z * c_odd;   c_even = sd(F[11]) - z * c_even;  // 3!, 2!
    c3 = c_odd; c2 = c_even;
    c1 = sd(1.) - z * c_odd;
    c0 = sd(1.) - z * c_even;
    for (; n > 0; n--) {
        c3 = (c2 + c0 * c3) * sd(0.25);
        c2 = c1 * c1 * sd(0.5);
        c1 = c0 * c1;
        c0 = sd(2.) * c0 * c0 - sd(1.);
    }
}
__device__ __forceinline__ void stiefel_gs3(sd beta, sd x, sd& g0, sd& g1, sd& g2, sd& g3) {
    sd x2 = x * x;
    stumpff_cs3(beta * x2, g0, g1, g2, g3);
    g1 = g1 * x; g2 = g2 * x2; g3 = g3 * (x2 * x);
}

// Universal-variable Kepler drift of one body (whfast.rs:676-833), strict arithmetic in the reference's
// association order. `work` lanes only; the warp iterates until every working lane has met the reference's
// exit test (exact repetition of x).
__device__ __forceinline__ void kepler_step(bool work, S3& pos, S3& vel, sd mu_in, sd dt, bool& tswarn, unsigned int& warnings) {
    // idle lanes (host slot, padding, stopped systems) get a benign state so that no lane drags the warp into the
    // slow paths of the IEEE division / square root (zero or NaN operands)
    const S3 p1 = work ? pos : s3(sd(1.), sd(0.), sd(0.));
    const S3 v1 = work ? vel : s3(sd(0.), sd(0.), sd(0.));
    const sd mu = work ? mu_in : sd(0.);
    const sd r0 = ssqrt(p1.x * p1.x + p1.y * p1.y + p1.z * p1.z);
    const sd r0i = sd(1.) / r0;
    const sd v2 = v1.x * v1.x + v1.y * v1.y + v1.z * v1.z;
    const sd beta = sd(2.) * mu * r0i - v2;
    const sd eta0 = p1.x * v1.x + p1.y * v1.y + p1.z * v1.z;
    const sd zeta0 = mu - beta * r0;
    sd x, g0, g1, g2, g3;
    const bool elliptic = beta.v > 0.;
    const sd two_pi = sd(2. * kPi);
    // invperiod = sqrt(beta) beta / (2 pi mu) and x_per_period = 2 pi / sqrt(beta) only feed two threshold tests and the
    // rare fallbacks. The tests are first decided on squared quantities with a factor-2 guard band (no sqrt/div);
    // only a lane inside the band evaluates the reference's exact expression.
    if (elliptic) {
        if (work && !tswarn) {
            double w = (dt.v * dt.v) * (beta.v * beta.v * beta.v), lim = (two_pi.v * mu.v) * (two_pi.v * mu.v);
            bool warn = w > 2. * lim;
            if (!warn && w > 0.5 * lim) { sd ip = div_ieee(ssqrt_ieee(beta) * beta, two_pi * mu); warn = fabs(dt.v) * ip.v > 1.; }
            if (warn) { tswarn = true; warnings |= PB200_WARN_TIMESTEP_GT_PERIOD; }
        }
        sd dtr0i = dt * r0i;
        x = dtr0i * (sd(1.) - dtr0i * eta0 * sd(0.5) * r0i);
    } else {
        x = sd(0.);
    }
    bool converged = false;
    sd old_x = x;
    stiefel_gs3(beta, x, g0, g1, g2, g3);
    sd e1 = eta0 * g1 + zeta0 * g2;
    sd ri = sd(1.) / (r0 + e1);
    x = ri * (x * e1 - eta0 * g2 - zeta0 * g3 + dt);
    bool quartic = false;
    if (work && elliptic) {
        double dx = (x - old_x).v;
        double qv = dx * dx * beta.v, lim = (0.01 * two_pi.v) * (0.01 * two_pi.v);
        quartic = qv > 2. * lim;
        if (!quartic && qv > 0.5 * lim) { sd xpp = div_ieee(two_pi, ssqrt_ieee(beta)); quartic = fabs(dx) > (sd(0.01) * xpp).v; }
    }
    if (__any_sync(FULL, quartic)) {
        // Laguerre-like quartic solver (whfast.rs:732-755), rare: large steps only.
        if (quartic) {
            x = div_ieee(beta * dt, mu);
            double prev_x[64];
            for (int k = 0; k < 64; k++) prev_x[k] = 0.;
            for (int n_lag = 1; n_lag < 64; n_lag++) {
                stiefel_gs3(beta, x, g0, g1, g2, g3);
                sd f = r0 * x + eta0 * g2 + zeta0 * g3 - dt;
                sd fp = r0 + eta0 * g1 + zeta0 * g2;
                sd fpp = eta0 * g0 + zeta0 * g1;
                sd denom = fp + ssqrt_ieee(sabs(sd(16.) * fp * fp - sd(20.) * f * fpp));
                old_x = x;   // the G-functions stay those of this x (the Newton loop below re-evaluates them in place)
                x = div_ieee(x * denom - sd(5.) * f, denom);
                bool hit = false;
                for (int k = 1; k < n_lag; k++) if (x.v == prev_x[k]) hit = true;
                if (hit) { converged = true; break; }
                prev_x[n_lag] = x.v;
            }
            ri = div_ieee(sd(1.), r0 + (eta0 * g1 + zeta0 * g2));
        }
    }
    {
        // Newton's method (whfast.rs:757-773): at most 31 passes, exit on x == old_x || x == old_x2.
        bool active = work && !quartic;
#pragma unroll 1
        for (int k = 1; k < 32; k++) {
            if (!__any_sync(FULL, active)) break;
            // The G-functions and 1/r are evaluated in place: a lane that has finished (or never worked) re-evaluates them at
            // the x they were computed from (old_x) and gets the same bits back, instead of every lane copying eight doubles
            // under a predicate on every pass.
            const sd xe = active ? x : old_x;
            stiefel_gs3(beta, xe, g0, g1, g2, g3);
            sd e = eta0 * g1 + zeta0 * g2;
            ri = sd(1.) / (r0 + e);
            sd xn = ri * (xe * e - eta0 * g2 - zeta0 * g3 + dt);
            if (active) {
                sd old_x2 = old_x;
                old_x = x; x = xn;
                if (x.v == old_x.v || x.v == old_x2.v) { converged = true; active = false; }
            }
        }
    }
    const bool bisect = work && !converged;
    if (__any_sync(FULL, bisect)) {
        if (bisect) {
            sd x_min, x_max;
            if (elliptic) {
                sd sqrt_beta = ssqrt_ieee(beta);
                sd invperiod = div_ieee(sqrt_beta * beta, two_pi * mu);
                sd x_per_period = div_ieee(two_pi, sqrt_beta);
                x_min = x_per_period * sd(floor((dt * invperiod).v));
                x_max = x_min + x_per_period;
            } else {
                sd h2 = r0 * r0 * v2 - eta0 * eta0;
                sd q = div_ieee(div_ieee(h2, mu), sd(1.) + ssqrt_ieee(sd(1.) - div_ieee(h2 * beta, mu * mu)));
                sd vq = div_ieee(ssqrt_ieee(h2), q);
                x_min = div_ieee(sd(1.), vq + div_ieee(r0, dt));
                x_max = div_ieee(dt, q);
            }
            x = div_ieee(x_max + x_min, sd(2.));
            for (int guard = 0; guard < 200; guard++) {   // the reference's `loop {}` never ends on NaN input
                stiefel_gs3(beta, x, g0, g1, g2, g3);
                sd s = r0 * x + eta0 * g2 + zeta0 * g3 - dt;
                if (s.v >= 0.) x_max = x; else x_min = x;
                x = div_ieee(x_max + x_min, sd(2.));
                if (div_ieee(sabs(x_max - x_min), x_max).v <= 1e-15) break;
            }
            ri = div_ieee(sd(1.), r0 + (eta0 * g1 + zeta0 * g2));
        }
    }
    if (isnan(ri.v)) { ri = sd(0.); g1 = sd(0.); g2 = sd(0.); g3 = sd(0.); }
    sd f = -mu * g2 * r0i;
    sd g = dt - mu * g3;
    sd fd = -mu * g1 * r0i * ri;
    sd gd = -mu * g2 * ri;
    if (work) {
        pos.x = pos.x + (f * p1.x + g * v1.x); pos.y = pos.y + (f * p1.y + g * v1.y); pos.z = pos.z + (f * p1.z + g * v1.z);
        vel.x = vel.x + (fd * p1.x + gd * v1.x); vel.y = vel.y + (fd * p1.y + gd * v1.y); vel.z = vel.z + (fd * p1.z + gd * v1.z);
    }
}

// upper-bound search + linear interpolation (tools.rs:840-907): first row with time > t
__device__ __forceinline__ int table_upper(const double* __restrict__ time, int n, double t) {
    int lo = 0, hi = n;  // first index with time[i] > t, n if none
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (__ldg(time + mid) > t) hi = mid; else lo = mid + 1;
    }
    return lo;
}
__device__ __forceinline__ double table_interp(const double* __restrict__ time, const double* __restrict__ y, int n, int i,
                                               double t) {
    if (i == 0) return __ldg(y);
    if (i >= n) return __ldg(y + n - 1);
    // strict, in the order of tools.rs:853-855
    sd xl = sd(__ldg(time + i - 1)), xr = sd(__ldg(time + i));
    sd pct = div_ieee(sd(t) - xl, xr - xl);
    return (sd(__ldg(y + i - 1)) * (sd(1.) - pct) + sd(__ldg(y + i)) * pct).v;
}

// ---------------------------------------------------------------------------------------------
// Register / shared-memory split. Only what every phase touches lives in registers (Lane). Everything that is
// read once per evaluation or once per step — per-system constants, the Kahan residuals, the midpoint's
// originals, the position while the midpoint runs — lives in per-thread shared-memory slots (Cold), addressed
// [slot][thread] so that a warp's access is one conflict-free 256-byte row. `volatile` keeps the compiler from
// forwarding the stored values back into registers.
#ifndef PB_BLOCK
#define PB_BLOCK 64
#endif

// (the slot map itself — enum ColdSlot — is per geometry build: cold_slots.cuh)
//
// The same columns double as the exchange medium inside a group: a lane leaves a value in its own column of a slot,
// __syncwarp(), and any lane of the group reads it with one LDS.64 (`getk`: column of body k). That replaces the
// two 32-bit shuffles per double plus the moves ptxas needs to re-pair the halves under register pressure
// (profiles/r1_variants.md). While the drift-kick-drift core runs, the midpoint's working set (S_VOX .. S_DLZ, S_RX..)
// is dead and serves as exchange space (E_A .. E_R).
struct Cold {
    volatile double* base;  // shared memory + threadIdx.x
    volatile double* grp;   // base - b: the column of the group's body 0
    volatile double* pair;  // this thread's 16-byte cell in the first two slots of a pair region (inside the warp's own columns)
    __device__ __forceinline__ double getk(int k, int slot) const { return grp[k + slot * PB_BLOCK]; }
    __device__ __forceinline__ double get(int slot) const { return base[slot * PB_BLOCK]; }
    __device__ __forceinline__ void set(int slot, double v) const { base[slot * PB_BLOCK] = v; }
    __device__ __forceinline__ V3 get3(int slot) const { return v3(get(slot), get(slot + 1), get(slot + 2)); }
    __device__ __forceinline__ void set3(int slot, V3 v) const { set(slot, v.x); set(slot + 1, v.y); set(slot + 2, v.z); }
    __device__ __forceinline__ V3 getk3(int k, int slot) const { return v3(getk(k, slot), getk(k, slot + 1), getk(k, slot + 2)); }
    // Pair regions: values that are always read together by their own thread (force constants, the midpoint's originals,
    // increments and Kahan residuals) sit as 16-byte cells inside a run of 2 n consecutive slots (cell p of a warp = its 32 columns of slots 2p and 2p + 1), so one
    // LDS.128 / STS.128 moves two of them: the same shared-memory wavefronts, half the instructions. `region` is the first
    // slot of the run; a region is only ever accessed through these four functions while it holds pairs.
    __device__ __forceinline__ unsigned pair_addr(int region, int p) const {
        return (unsigned)__cvta_generic_to_shared((const void*)(pair + (region + 2 * p) * PB_BLOCK));
    }
    __device__ __forceinline__ double2 get2(int region, int p) const {
        double2 v;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(pair_addr(region, p)) : "memory");
        return v;
    }
    __device__ __forceinline__ void set2(int region, int p, double a, double b) const {
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" : : "r"(pair_addr(region, p)), "d"(a), "d"(b) : "memory");
    }
    // two vectors in three cells: (a.x, a.y) (a.z, b.x) (b.y, b.z)
    __device__ __forceinline__ void get6(int region, V3& a, V3& b) const {
        const double2 p0 = get2(region, 0), p1 = get2(region, 1), p2 = get2(region, 2);
        a = v3(p0.x, p0.y, p1.x); b = v3(p1.y, p2.x, p2.y);
    }
    __device__ __forceinline__ void set6(int region, V3 a, V3 b) const {
        set2(region, 0, a.x, a.y); set2(region, 1, a.z, b.x); set2(region, 2, b.y, b.z);
    }
};
// Per-lane register state.
struct Lane {
    S3 r, v;            // inertial position / velocity (strict arithmetic only)
    V3 L, s;            // angular momentum, spin (of the previous evaluation)
    double rs_s, rs_p;  // fast mode: r . w_host, r . w_planet with the spins of the previous evaluation (Q3)
};
__device__ __forceinline__ V3 plain(S3 a) { return v3(a.x.v, a.y.v, a.z.v); }
__device__ __forceinline__ S3 strict(V3 a) { return s3(sd(a.x), sd(a.y), sd(a.z)); }
__device__ __forceinline__ S3 shfl3(S3 a, int src) { return s3(sd(shfl(a.x.v, src)), sd(shfl(a.y.v, src)), sd(shfl(a.z.v, src))); }
__device__ __forceinline__ sd shfl(sd a, int src) { return sd(shfl(a.v, src)); }

__device__ __forceinline__ double pow5(double x) { double x2 = x * x; return x2 * x2 * x; }

struct Roles {
    bool valid;     // lane belongs to a live system slot and b < n_bodies
    bool host;
    bool planet;    // valid && !host
    bool t_on, f_on, g_on;  // OrbitingBody for tides / flattening / GR
};

}  // namespace pb200
