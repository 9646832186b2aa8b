// oracle_core.hpp — CPU restatement of the Posidonius WHFast hot path.
//
// TEST INFRASTRUCTURE ONLY. This file is the parity oracle and the timed CPU
// baseline. Nothing in the product (posidonius_b200/) may include, link or call
// it; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs do.
//
// Parity status: PINNED. The restatement reproduces the reference's own golden
// vectors tests/data/<fixture>/particle_{0..4}.json (199 steps, dt = 0.08 d,
// absolute 1e-14 — tests/common/universe.rs:48-72) for every in-scope fixture;
// the vectors are committed under tests/golden/ with the script that copied
// them. The reference itself is Rust and cannot be built in this image (no
// cargo/rustc), so there is no oracle/_ref.
//
// Every function cites the reference lines it restates (paths relative to
// /root/reference/src). Arithmetic is written in the reference's association
// order and compiled with -ffp-contract=off so that results agree to the last
// bits; `powi` follows LLVM's square-and-multiply expansion.
//
// The code is a template on the scalar type: R = double is the oracle proper,
// R = Counted (oracle_count.cpp) counts + - * / sqrt for the roofline numerator.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "../include/posidonius_b200.h"

namespace pb200_oracle {

// ---- constants (constants.rs:20-67), evaluated in the same order as rustc's const folding
constexpr int MAXP = PB200_MAX_PARTICLES;
constexpr double C_PI = 3.14159265358979323846264338327950288;
constexpr int WHFAST_NMAX_QUART = 64;
constexpr int WHFAST_NMAX_NEWT = 32;
constexpr double DBL_EPSILON_ = 2.2204460492503131e-16;
constexpr double DBL_EPSILON_2 = DBL_EPSILON_ * DBL_EPSILON_;
constexpr int IMPLICIT_MIDPOINT_MIN_ITER = 3;
constexpr int IMPLICIT_MIDPOINT_MAX_ITER = 10;
constexpr double MAX_DISTANCE = 100.;
constexpr double MAX_DISTANCE_2 = MAX_DISTANCE * MAX_DISTANCE;
constexpr double HOUR = 3600.;
constexpr double DAY = 24. * HOUR;
constexpr double M_SUN = 1.9818e30;
constexpr double G_SI = 6.67428e-11;
constexpr double AU = 1.49597870700e11;
constexpr double SPEED_OF_LIGHT = (2.99792458e8 / AU) * DAY;
constexpr double SPEED_OF_LIGHT_2 = SPEED_OF_LIGHT * SPEED_OF_LIGHT;
constexpr double R_SUN = 6.957e8 / AU;
constexpr double G_MERCURY = G_SI / (AU * AU * AU) * M_SUN * (DAY * DAY);
constexpr double G = G_MERCURY;
constexpr double K2 = G_MERCURY;
constexpr double SUN_DYN_FREQ_2 = K2 / (R_SUN * R_SUN * R_SUN);
constexpr double SMOOTHING_FACTOR_DYN_TIDE_COROTATION = 1.0e-5 * 86400.;

inline double o_sqrt(double x) { return std::sqrt(x); }
inline double o_abs(double x) { return std::fabs(x); }
inline double o_floor(double x) { return std::floor(x); }
inline bool o_isnan(double x) { return std::isnan(x); }
inline double o_pow(double x, double y) { return std::pow(x, y); }
inline double o_val(double x) { return x; }

// f64::powi as lowered by LLVM (SelectionDAGBuilder ExpandPowI): square-and-multiply, LSB first.
template <class R>
inline R powi(R x, int n) {
    R res = x;
    bool have = false;
    R sq = x;
    unsigned v = (unsigned)n;
    while (v) {
        if (v & 1u) {
            if (have) res = res * sq; else { res = sq; have = true; }
        }
        v >>= 1;
        if (v) sq = sq * sq;
    }
    return res;
}

template <class R> struct V3 { R x, y, z; };

// particles/particle.rs:16-52 — the fields the path touches, with the per-effect scratch
template <class R>
struct Body {
    int id = 0;
    R mass = 0., mass_g = 0., radius = 0.;
    V3<R> ipos{0., 0., 0.}, ivel{0., 0., 0.}, iacc{0., 0., 0.}, iadd{0., 0., 0.};
    V3<R> hpos{0., 0., 0.}, hvel{0., 0., 0.};
    R hdist = 0., hradvel = 0., hnormv = 0., hnormv2 = 0.;
    V3<R> spin{0., 0., 0.};
    R norm_spin2 = 0.;
    V3<R> L{0., 0., 0.}, dLdt{0., 0., 0.};
    R rg2 = 0., moi = 0.;
    // tides (tides/common.rs:41-100, constant_time_lag.rs:12-18)
    int t_role = PB200_ROLE_DISABLED;
    R t_dissipation_factor = 0., t_dissipation_factor_scale = 0., t_k2 = 0., t_sigma = 0.;
    R t_dist = 0., t_radvel = 0., t_rs_star = 0., t_rs_planet = 0., t_orth_star = 0., t_orth_planet = 0.;
    R t_radial = 0., t_radial_diss_pm = 0., t_denergy = 0., t_lag = 0.;
    V3<R> t_pos{0., 0., 0.}, t_vel{0., 0., 0.}, t_acc{0., 0., 0.}, t_dL{0., 0., 0.};
    // rotational flattening (rotational_flattening/common.rs:8-62)
    int f_role = PB200_ROLE_DISABLED;
    R f_k2 = 0.;
    R f_dist = 0., f_rs_star = 0., f_rs_planet = 0., f_radial = 0., f_factor_star = 0., f_factor_planet = 0.;
    R f_orth_star = 0., f_orth_planet = 0.;
    V3<R> f_pos{0., 0., 0.}, f_vel{0., 0., 0.}, f_acc{0., 0., 0.}, f_dL{0., 0., 0.};
    // general relativity (general_relativity.rs:10-60)
    int g_role = PB200_ROLE_DISABLED;
    R g_factor = 0., g_dist = 0., g_radvel = 0., g_normv = 0., g_normv2 = 0.;
    V3<R> g_pos{0., 0., 0.}, g_vel{0., 0., 0.}, g_acc{0., 0., 0.}, g_dL{0., 0., 0.};
    // wind (wind.rs:6-39); the disk output stays zero (effect rejected by the product, kept for the sums of universe.rs:540-614)
    int w_role = 1;  // WindEffect tag order: 0 = Interaction, 1 = Disabled
    R w_k = 0., w_sat = 0., w_sat2 = 0.;
    V3<R> w_dL{0., 0., 0.}, d_acc{0., 0., 0.};
    int evo_type = PB200_EVO_NONEVOLVING;
    double evo_param = 0.;
    int evo_table = -1;
};

struct Table {
    std::vector<double> time, radius, rg2, love, qinv;
    size_t left_index = 0;
};

template <class R>
struct Alt { R mass = 0., mass_g = 0.; V3<R> pos{0., 0., 0.}, vel{0., 0., 0.}, acc{0., 0., 0.}; };

enum Ignore { IgnoreNone = 0, WHFastOne = 1, WHFastTwo = 2 };

template <class R>
struct System {
    // WHFast (integrator/whfast.rs:98-120)
    double time_step = 0., half_time_step = 0.;
    double current_time = 0.;
    uint64_t current_iteration = 0;
    double recovery_snapshot_period = 0., historic_snapshot_period = 0.;
    double last_recovery_snapshot_time = -1., last_historic_snapshot_time = -1.;
    uint64_t n_historic_snapshots = 0;
    Alt<R> alt[MAXP];
    int coord = PB200_COORD_JACOBI;
    uint64_t timestep_warning = 0;
    V3<R> verr[MAXP], lerr[MAXP];
    // Universe (particles/universe.rs:50-63)
    double initial_time = 0., time_limit = 0.;
    Body<R> p[MAXP];
    std::vector<Table> evolvers;  // one per body slot, own cursor each (Evolver.left_index)
    int n = 0;
    bool c_tides = false, c_flat = false, c_gr = false, c_disk = false, c_wind = false, c_evo = false;
    int gr_impl = PB200_GR_DISABLED;
    int h_mm = 0, h_tides = 0, h_flat = 0, h_gr = 0;
    bool mm_all = true, mm_tides = true, mm_flat = true, mm_gr = true;
    double roche[MAXP * MAXP];
    // ensemble status (replaces panic!)
    int status = PB200_STATUS_OK;
    uint32_t warnings = 0;
    uint64_t event_iteration = 0;
    // last historic record bytes requested by the caller
    std::vector<uint8_t> history;

    System() {
        for (int i = 0; i < MAXP; i++) { verr[i] = {0., 0., 0.}; lerr[i] = {0., 0., 0.}; }
        std::memset(roche, 0, sizeof(roche));
    }

    // ------------------------------------------------------------------ tools.rs:840-907
    static void find_indices(const double* data, size_t ndata, double target, size_t& left, size_t& right) {
        size_t last = ndata - 1;
        size_t i = 0;
        bool found = false;
        for (; i < ndata; i++) if (data[i] > target) { found = true; break; }
        if (!found) {
            if (data[last] > target) { left = 0; right = 0; } else { left = last; right = last; }
        } else if (i == 0) { left = 0; right = 0; }
        else { left = i - 1; right = i; }
    }
    static double linear_interpolation(double target_x, const double* x, const double* y, size_t n, size_t& left_out) {
        size_t left, right;
        find_indices(x, n, target_x, left, right);
        double target_y;
        if (left == right) target_y = y[left];
        else {
            double x_left = x[left];
            double pct = (target_x - x_left) / (x[right] - x_left);
            target_y = y[left] * (1. - pct) + y[right] * pct;
        }
        left_out = left;
        return target_y;
    }
    // effects/evolution.rs:449-456
    static size_t evo_idx(const Table& t) { return t.left_index > 10 ? t.left_index - 10 : 0; }
    double evo_interp(Table& t, const std::vector<double>& col, double time) {
        size_t i0 = evo_idx(t), left;
        double y = linear_interpolation(time, t.time.data() + i0, col.data() + i0, t.time.size() - i0, left);
        t.left_index += left;
        return y;
    }
    // effects/evolution.rs:458-513
    double evo_radius(int type, Table& t, double time, double cur) {
        if (type == PB200_EVO_NONEVOLVING) return cur;
        return evo_interp(t, t.radius, time);
    }
    double evo_rg2(int type, Table& t, double time, double cur) {
        if (type == PB200_EVO_BARAFFE2015 || type == PB200_EVO_LECONTE2011 || type == PB200_EVO_LECONTECHABRIER2013)
            return evo_interp(t, t.rg2, time);
        return cur;
    }
    double evo_love(int type, Table& t, double time, double cur) {
        if (type == PB200_EVO_LECONTECHABRIER2013) return evo_interp(t, t.love, time);
        return cur;
    }
    static bool evo_dynamical(int type, double param) {
        return type == PB200_EVO_BOLMONTMATHIS2016 || type == PB200_EVO_GALLETBOLMONT2017 ||
               (type == PB200_EVO_LECONTECHABRIER2013 && param != 0.);
    }
    double evo_qinv(int type, double param, Table& t, double time, double cur) {
        if (evo_dynamical(type, param)) return evo_interp(t, t.qinv, time);
        return cur;
    }

    // effects/evolution.rs:516-546
    void evolve_non_spin(double time) {
        for (int i = 0; i < n; i++) {
            Body<R>& b = p[i];
            Table& t = evolvers[i];
            double new_radius = evo_radius(b.evo_type, t, time, o_val(b.radius));
            double new_rg2 = evo_rg2(b.evo_type, t, time, o_val(b.rg2));
            if (new_radius != o_val(b.radius) || new_rg2 != o_val(b.rg2)) {
                b.radius = new_radius;
                b.rg2 = new_rg2;
                b.moi = b.mass * b.rg2 * powi(b.radius, 2);
            }
            // Q6: the love number is interpolated into a copy (evolution.rs:534-541); only the cursor moves.
            if (b.t_role != PB200_ROLE_DISABLED) (void)evo_love(b.evo_type, t, time, o_val(b.t_k2));
        }
    }
    // effects/evolution.rs:548-567
    void evolve_spin_dependent(double time) {
        for (int i = 0; i < n; i++) {
            Body<R>& b = p[i];
            if (evo_dynamical(b.evo_type, b.evo_param)) {
                R qinv = evo_qinv(b.evo_type, b.evo_param, evolvers[i], time, 0.);
                R eps2 = b.norm_spin2 / SUN_DYN_FREQ_2;
                b.t_lag = 3.0 * eps2 * qinv / 4.0;
            } else b.t_lag = 0.;
        }
    }
    // particles/common.rs:3-15
    void calculate_spin() {
        for (int i = 0; i < n; i++) {
            Body<R>& b = p[i];
            if (o_val(b.moi) == 0.) { fail(PB200_STATUS_ZERO_INERTIA); }
            b.spin.x = b.L.x / b.moi;
            b.spin.y = b.L.y / b.moi;
            b.spin.z = b.L.z / b.moi;
            b.norm_spin2 = (powi(b.spin.x, 2)) + (powi(b.spin.y, 2)) + (powi(b.spin.z, 2));
        }
    }
    // particles/universe.rs:305-316
    void calculate_spin_and_evolving_quantities(double time, bool evolution) {
        if (evolution && c_evo) evolve_non_spin(time);
        calculate_spin();
        if (evolution && c_evo) evolve_spin_dependent(time);
    }
    // particles/universe.rs:177-196
    void calculate_roche_radiuses() {
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) {
                if (i == j) continue;
                double ma = o_val(p[i].mass), mb = o_val(p[j].mass);
                if (ma > mb) roche[i * n + j] = (o_val(p[j].radius) / 0.462) * std::pow(ma / mb, 1. / 3.);
                else roche[i * n + j] = (o_val(p[i].radius) / 0.462) * std::pow(mb / ma, 1. / 3.);
            }
    }
    // integrator/whfast.rs:226-233
    int initialize_physical_values() {
        if (current_time != 0.) return PB200_E_INVALID;
        calculate_spin_and_evolving_quantities(current_time, true);
        calculate_roche_radiuses();
        return PB200_OK;
    }

    void fail(int st) {
        if (status == PB200_STATUS_OK) { status = st; event_iteration = current_iteration; }
    }

    // ------------------------------------------------------------------ particles/universe.rs:198-303
    void gravity_calculate_acceleration(Ignore ign) {
        V3<R> acc[MAXP];
        for (int i = 0; i < n; i++) acc[i] = {0., 0., 0.};
        for (int i = 0; i < n; i++) {
            for (int j = 0; j < n; j++) {
                if (i == j) continue;
                R dx = p[i].ipos.x - p[j].ipos.x;
                R dy = p[i].ipos.y - p[j].ipos.y;
                R dz = p[i].ipos.z - p[j].ipos.z;
                R d2 = dx * dx + dy * dy + dz * dz;
                if (i < j) {
                    double rr = roche[i * n + j];
                    if (o_val(d2) <= rr * rr) fail(PB200_STATUS_ROCHE_DESTROYED);
                    if (o_val(d2) <= o_val(powi(p[i].radius + p[j].radius, 2))) fail(PB200_STATUS_COLLISION);
                    if (MAX_DISTANCE_2 > 0. && i == h_mm && o_val(d2) > MAX_DISTANCE_2) fail(PB200_STATUS_EJECTED);
                }
                if (ign == WHFastOne && ((i == h_mm && h_mm == 0 && j == 1) || (i == h_mm && h_mm > 0 && j == 0) ||
                                         (j == h_mm && h_mm == 0 && i == 1) || (j == h_mm && h_mm > 0 && i == 0)))
                    continue;
                if (ign == WHFastTwo && (i == h_mm || j == h_mm)) continue;
                R d = o_sqrt(d2);
                R prefact = -G / (d * d * d) * p[j].mass;
                acc[i].x = acc[i].x + prefact * dx;
                acc[i].y = acc[i].y + prefact * dy;
                acc[i].z = acc[i].z + prefact * dz;
            }
        }
        for (int i = 0; i < n; i++) p[i].iacc = acc[i];
    }

    // ------------------------------------------------------------------ particles/universe.rs:318-351
    void inertial_to_heliocentric() {
        Body<R>& h = p[h_mm];
        for (int i = 0; i < n; i++) {
            if (i == h_mm) continue;
            Body<R>& b = p[i];
            b.hpos.x = b.ipos.x - h.ipos.x;
            b.hpos.y = b.ipos.y - h.ipos.y;
            b.hpos.z = b.ipos.z - h.ipos.z;
            b.hvel.x = b.ivel.x - h.ivel.x;
            b.hvel.y = b.ivel.y - h.ivel.y;
            b.hvel.z = b.ivel.z - h.ivel.z;
            b.hdist = o_sqrt(powi(b.hpos.x, 2) + powi(b.hpos.y, 2) + powi(b.hpos.z, 2));
            b.hradvel = (b.hpos.x * b.hvel.x + b.hpos.y * b.hvel.y + b.hpos.z * b.hvel.z) / b.hdist;
            // Q4: the host's heliocentric velocity is read before it is zeroed below
            b.hnormv2 = powi(b.hvel.x - h.hvel.x, 2) + powi(b.hvel.y - h.hvel.y, 2) + powi(b.hvel.z - h.hvel.z, 2);
            b.hnormv = o_sqrt(b.hnormv2);
        }
        h.hpos = {0., 0., 0.};
        h.hvel = {0., 0., 0.};
        h.hdist = 0.; h.hradvel = 0.; h.hnormv2 = 0.; h.hnormv = 0.;
    }

    // ------------------------------------------------------------------ per-effect coordinate copies + initialize
    // tides/common.rs:143-217
    void tides_copy_helio(int host) {
        Body<R>& h = p[host];
        if (h.t_role != PB200_ROLE_CENTRAL) return;
        h.t_pos = h.hpos; h.t_vel = h.hvel; h.t_dist = h.hdist; h.t_radvel = h.hradvel;
        for (int i = 0; i < n; i++) {
            if (i == host) continue;
            Body<R>& b = p[i];
            if (b.t_role != PB200_ROLE_ORBITING) continue;
            b.t_pos = b.hpos; b.t_vel = b.hvel; b.t_dist = b.hdist; b.t_radvel = b.hradvel;
        }
    }
    void tides_inertial_to_helio(int host) {
        Body<R>& h = p[host];
        if (h.t_role != PB200_ROLE_CENTRAL) return;
        h.t_pos = {0., 0., 0.}; h.t_vel = {0., 0., 0.}; h.t_dist = 0.; h.t_radvel = 0.;
        for (int i = 0; i < n; i++) {
            if (i == host) continue;
            Body<R>& b = p[i];
            if (b.t_role != PB200_ROLE_ORBITING) continue;
            b.t_pos.x = b.ipos.x - h.ipos.x; b.t_pos.y = b.ipos.y - h.ipos.y; b.t_pos.z = b.ipos.z - h.ipos.z;
            b.t_vel.x = b.ivel.x - h.ivel.x; b.t_vel.y = b.ivel.y - h.ivel.y; b.t_vel.z = b.ivel.z - h.ivel.z;
            b.t_dist = o_sqrt(powi(b.t_pos.x, 2) + powi(b.t_pos.y, 2) + powi(b.t_pos.z, 2));
            b.t_radvel = (b.t_pos.x * b.t_vel.x + b.t_pos.y * b.t_vel.y + b.t_pos.z * b.t_vel.z) / b.t_dist;
        }
    }
    void tides_initialize(int host) {
        Body<R>& h = p[host];
        if (h.t_role != PB200_ROLE_CENTRAL) return;
        h.t_rs_star = 0.; h.t_rs_planet = 0.; h.t_acc = {0., 0., 0.}; h.t_dL = {0., 0., 0.};
        for (int i = 0; i < n; i++) {
            if (i == host) continue;
            Body<R>& b = p[i];
            if (b.t_role != PB200_ROLE_ORBITING) continue;
            // Q3: spins are those of the previous evaluation
            b.t_rs_star = b.t_pos.x * h.spin.x + b.t_pos.y * h.spin.y + b.t_pos.z * h.spin.z;
            b.t_rs_planet = b.t_pos.x * b.spin.x + b.t_pos.y * b.spin.y + b.t_pos.z * b.spin.z;
            b.t_acc = {0., 0., 0.}; b.t_dL = {0., 0., 0.};
        }
    }
    // rotational_flattening/common.rs:93-161
    void flat_copy_helio(int host) {
        Body<R>& h = p[host];
        if (h.f_role != PB200_ROLE_CENTRAL) return;
        h.f_pos = h.hpos; h.f_vel = h.hvel; h.f_dist = h.hdist;
        for (int i = 0; i < n; i++) {
            if (i == host) continue;
            Body<R>& b = p[i];
            if (b.f_role != PB200_ROLE_ORBITING) continue;
            b.f_pos = b.hpos; b.f_vel = b.hvel; b.f_dist = b.hdist;
        }
    }
    void flat_inertial_to_helio(int host) {
        Body<R>& h = p[host];
        if (h.f_role != PB200_ROLE_CENTRAL) return;
        h.f_pos = {0., 0., 0.}; h.f_vel = {0., 0., 0.}; h.f_dist = 0.;
        for (int i = 0; i < n; i++) {
            if (i == host) continue;
            Body<R>& b = p[i];
            if (b.f_role != PB200_ROLE_ORBITING) continue;
            b.f_pos.x = b.ipos.x - h.ipos.x; b.f_pos.y = b.ipos.y - h.ipos.y; b.f_pos.z = b.ipos.z - h.ipos.z;
            b.f_vel.x = b.ivel.x - h.ivel.x; b.f_vel.y = b.ivel.y - h.ivel.y; b.f_vel.z = b.ivel.z - h.ivel.z;
            b.f_dist = o_sqrt(powi(b.f_pos.x, 2) + powi(b.f_pos.y, 2) + powi(b.f_pos.z, 2));
        }
    }
    void flat_initialize(int host) {
        Body<R>& h = p[host];
        if (h.f_role != PB200_ROLE_CENTRAL) return;
        h.f_rs_star = 0.; h.f_rs_planet = 0.; h.f_acc = {0., 0., 0.}; h.f_dL = {0., 0., 0.};
        for (int i = 0; i < n; i++) {
            if (i == host) continue;
            Body<R>& b = p[i];
            if (b.f_role != PB200_ROLE_ORBITING) continue;
            b.f_rs_star = b.f_pos.x * h.spin.x + b.f_pos.y * h.spin.y + b.f_pos.z * h.spin.z;
            b.f_rs_planet = b.f_pos.x * b.spin.x + b.f_pos.y * b.spin.y + b.f_pos.z * b.spin.z;
            b.f_acc = {0., 0., 0.}; b.f_dL = {0., 0., 0.};
        }
    }
    // general_relativity.rs:86-174
    bool gr_host_active(const Body<R>& h) const { return h.g_role == PB200_ROLE_CENTRAL && gr_impl_of_host != PB200_GR_DISABLED; }
    int gr_impl_of_host = PB200_GR_DISABLED;  // payload of GeneralRelativityEffect::CentralBody(impl) on the host
    void gr_copy_helio(int host) {
        Body<R>& h = p[host];
        if (!gr_host_active(h)) return;
        h.g_pos = h.hpos; h.g_vel = h.hvel; h.g_dist = h.hdist; h.g_radvel = h.hradvel; h.g_normv = h.hnormv; h.g_normv2 = h.hnormv2;
        for (int i = 0; i < n; i++) {
            if (i == host) continue;
            Body<R>& b = p[i];
            if (b.g_role != PB200_ROLE_ORBITING) continue;
            b.g_pos = b.hpos; b.g_vel = b.hvel; b.g_dist = b.hdist; b.g_radvel = b.hradvel; b.g_normv = b.hnormv; b.g_normv2 = b.hnormv2;
        }
    }
    void gr_inertial_to_helio(int host) {
        Body<R>& h = p[host];
        if (!gr_host_active(h)) return;
        h.g_pos = {0., 0., 0.}; h.g_vel = {0., 0., 0.}; h.g_dist = 0.; h.g_radvel = 0.; h.g_normv = 0.; h.g_normv2 = 0.;
        for (int i = 0; i < n; i++) {
            if (i == host) continue;
            Body<R>& b = p[i];
            if (b.g_role != PB200_ROLE_ORBITING) continue;
            b.g_pos.x = b.ipos.x - h.ipos.x; b.g_pos.y = b.ipos.y - h.ipos.y; b.g_pos.z = b.ipos.z - h.ipos.z;
            b.g_vel.x = b.ivel.x - h.ivel.x; b.g_vel.y = b.ivel.y - h.ivel.y; b.g_vel.z = b.ivel.z - h.ivel.z;
            b.g_dist = o_sqrt(powi(b.g_pos.x, 2) + powi(b.g_pos.y, 2) + powi(b.g_pos.z, 2));
            b.g_radvel = (b.g_pos.x * b.g_vel.x + b.g_pos.y * b.g_vel.y + b.g_pos.z * b.g_vel.z) / b.g_dist;
            b.g_normv2 = powi(b.g_vel.x - h.g_vel.x, 2) + powi(b.g_vel.y - h.g_vel.y, 2) + powi(b.g_vel.z - h.g_vel.z, 2);
            b.g_normv = o_sqrt(b.g_normv2);
        }
    }
    void gr_initialize(int host) {
        Body<R>& h = p[host];
        if (h.g_role != PB200_ROLE_CENTRAL) return;
        if (gr_impl_of_host == PB200_GR_NEWHALL1983 || gr_impl_of_host == PB200_GR_DISABLED) return;
        h.g_factor = 0.; h.g_acc = {0., 0., 0.}; h.g_dL = {0., 0., 0.};
        for (int i = 0; i < n; i++) {
            if (i == host) continue;
            Body<R>& b = p[i];
            if (b.g_role != PB200_ROLE_ORBITING) continue;
            b.g_factor = h.mass_g * b.mass_g / powi(h.mass_g + b.mass_g, 2);
            b.g_acc = {0., 0., 0.}; b.g_dL = {0., 0., 0.};
        }
    }

    // particles/universe.rs:353-426
    void initialize(bool dL, bool acc) {
        bool init_tides = (dL && c_tides) || (acc && c_tides);
        bool init_flat = (dL && c_flat) || (acc && c_flat);
        bool init_gr = (dL && c_gr && gr_impl == PB200_GR_KIDDER1995) || (acc && c_gr);
        if (init_tides || init_flat || init_gr) {
            if (init_tides && mm_tides) { tides_copy_helio(h_mm); tides_initialize(h_mm); }
            if (init_flat && mm_flat) { flat_copy_helio(h_mm); flat_initialize(h_mm); }
            if (init_gr && mm_gr) { gr_copy_helio(h_mm); gr_initialize(h_mm); }
            if (!mm_all) {
                if (init_tides && !mm_tides) { tides_inertial_to_helio(h_tides); tides_initialize(h_tides); }
                if (init_flat && !mm_flat) { flat_inertial_to_helio(h_flat); flat_initialize(h_flat); }
                if (init_gr && !mm_gr) { gr_inertial_to_helio(h_gr); gr_initialize(h_gr); }
            }
        }
        // wind::initialize (wind.rs:62-70; universe.rs:356-366)
        if (dL && c_wind) for (int i = 0; i < n; i++) if (p[i].w_role == 0) p[i].w_dL = {0., 0., 0.};
        if (acc) for (int i = 0; i < n; i++) p[i].iadd = {0., 0., 0.};
    }

    // ------------------------------------------------------------------ wind.rs:72-91
    void calculate_wind_factor() {
        for (int i = 0; i < n; i++) {
            Body<R>& b = p[i];
            if (b.w_role != 0) continue;
            R threshold = o_sqrt(b.norm_spin2);
            if (o_val(threshold) >= o_val(b.w_sat)) {
                // fast rotator
                b.w_dL.x = -1. * b.w_k * b.spin.x * b.w_sat2 * o_sqrt(b.radius / R_SUN * 1. / b.mass);
                b.w_dL.y = -1. * b.w_k * b.spin.y * b.w_sat2 * o_sqrt(b.radius / R_SUN * 1. / b.mass);
                b.w_dL.z = -1. * b.w_k * b.spin.z * b.w_sat2 * o_sqrt(b.radius / R_SUN * 1. / b.mass);
            } else {
                b.w_dL.x = -1. * b.w_k * b.spin.x * b.norm_spin2 * o_sqrt(b.radius / R_SUN * 1. / b.mass);
                b.w_dL.y = -1. * b.w_k * b.spin.y * b.norm_spin2 * o_sqrt(b.radius / R_SUN * 1. / b.mass);
                b.w_dL.z = -1. * b.w_k * b.spin.z * b.norm_spin2 * o_sqrt(b.radius / R_SUN * 1. / b.mass);
            }
        }
    }

    // ------------------------------------------------------------------ constant-time-lag tides
    // pair-dependent sigma (constant_time_lag.rs:20-165) — only dynamical-tide evolution types
    double pair_sigma[MAXP * MAXP];
    bool pair_sigma_set[MAXP * MAXP];
    void pair_clear() { for (int i = 0; i < MAXP * MAXP; i++) { pair_sigma[i] = 0.; pair_sigma_set[i] = false; } }
    // tools.rs:251-285
    static void perihelion_and_ecc(R gm, V3<R> pos, V3<R> vel, R& q, R& e) {
        R x = pos.x, y = pos.y, z = pos.z, u = vel.x, v = vel.y, w = vel.z;
        R hx = y * w - z * v, hy = z * u - x * w, hz = x * v - y * u;
        R h2 = o_pow(hx, 2.) + o_pow(hy, 2.) + o_pow(hz, 2.);
        R v2 = u * u + v * v + w * w;
        R r = o_sqrt(x * x + y * y + z * z);
        R s = h2 / gm;
        R temp = 1. + s * (v2 / gm - 2. / r);
        if (o_val(temp) <= 0.) e = 0.; else e = o_sqrt(temp);
        q = s / (1. + e);
    }
    R get_sigma(int id, int dep, int evo_type, double evo_param, R fallback) {
        if (evo_dynamical(evo_type, evo_param)) {
            int key = id * MAXP + dep;
            if (pair_sigma_set[key]) return pair_sigma[key];
        }
        return fallback;
    }
    void tides_pair_sigma(int host) {
        Body<R>& h = p[host];
        if (evo_dynamical(h.evo_type, h.evo_param)) {
            R host_norm_spin = o_sqrt(h.norm_spin2);
            for (int i = 0; i < n; i++) {
                if (i == host) continue;
                Body<R>& b = p[i];
                if (b.t_role != PB200_ROLE_ORBITING) continue;
                R scale = h.t_role == PB200_ROLE_CENTRAL ? h.t_dissipation_factor_scale : R(0.);
                R diss = h.t_role == PB200_ROLE_CENTRAL ? h.t_dissipation_factor : R(0.);
                R gm = h.mass_g + b.mass_g;
                R q, e;
                perihelion_and_ecc(gm, b.t_pos, b.t_vel, q, e);
                R mean_motion = o_sqrt(gm) * o_pow(q / (1.0 - e), -1.5);
                R half = o_abs(host_norm_spin - mean_motion);
                if (o_val(half) < o_val(host_norm_spin)) {
                    if (o_val(half) < SMOOTHING_FACTOR_DYN_TIDE_COROTATION) half = SMOOTHING_FACTOR_DYN_TIDE_COROTATION;
                    R inv = 1. / half;
                    R s = scale * (2.0 * K2 / (3.0 * powi(h.radius, 5)) * h.t_lag * inv + diss);
                    pair_sigma[h.id * MAXP + b.id] = o_val(s); pair_sigma_set[h.id * MAXP + b.id] = true;
                } else pair_sigma_set[h.id * MAXP + b.id] = false;
            }
        }
        for (int i = 0; i < n; i++) {
            if (i == host) continue;
            Body<R>& b = p[i];
            if (!evo_dynamical(b.evo_type, b.evo_param)) continue;
            if (b.t_role != PB200_ROLE_ORBITING) continue;
            R norm_spin = o_sqrt(b.norm_spin2);
            R gm = h.mass_g + b.mass_g;
            R q, e;
            perihelion_and_ecc(gm, b.t_pos, b.t_vel, q, e);
            R mean_motion = o_sqrt(gm) * o_pow(q / (1.0 - e), -1.5);
            R half = o_abs(norm_spin - mean_motion);
            if (o_val(half) < o_val(norm_spin)) {
                if (o_val(half) < SMOOTHING_FACTOR_DYN_TIDE_COROTATION) half = SMOOTHING_FACTOR_DYN_TIDE_COROTATION;
                R inv = 1. / half;
                R s = b.t_dissipation_factor_scale * (2.0 * K2 / (3.0 * powi(b.radius, 5)) * b.t_lag * inv + b.t_dissipation_factor);
                pair_sigma[b.id * MAXP + h.id] = o_val(s); pair_sigma_set[b.id * MAXP + h.id] = true;
            } else pair_sigma_set[h.id * MAXP + b.id] = false;  // Q7: (host, particle) key
        }
    }
    // constant_time_lag.rs:206-264
    void tides_orthogonal(int host) {
        Body<R>& h = p[host];
        for (int pass = 0; pass < 2; pass++) {
            bool central = pass == 0;
            for (int i = 0; i < n; i++) {
                if (i == host) continue;
                Body<R>& b = p[i];
                if (b.t_role != PB200_ROLE_ORBITING) continue;
                R d7 = powi(b.t_dist, 7);
                if (central) {
                    R sig = get_sigma(h.id, b.id, h.evo_type, h.evo_param, h.t_sigma);
                    b.t_orth_star = 4.5 * (powi(b.mass, 2)) * (powi(h.radius, 10)) * sig / d7;
                } else {
                    R sig = get_sigma(b.id, h.id, b.evo_type, b.evo_param, b.t_sigma);
                    b.t_orth_planet = 4.5 * (powi(h.mass, 2)) * (powi(b.radius, 10)) * sig / d7;
                }
            }
        }
    }
    // constant_time_lag.rs:266-307
    void tides_radial(int host) {
        Body<R>& h = p[host];
        R host_mass_2 = h.mass * h.mass;
        for (int i = 0; i < n; i++) {
            if (i == host) continue;
            Body<R>& b = p[i];
            if (b.t_role != PB200_ROLE_ORBITING) continue;
            R host_k2 = h.t_role == PB200_ROLE_CENTRAL ? h.t_k2 : R(0.);
            R m2 = b.mass * b.mass;
            R cons = -3.0 * K2 / powi(b.t_dist, 7) * (m2 * powi(h.radius, 5) * host_k2 + host_mass_2 * powi(b.radius, 5) * b.t_k2);
            R factor1 = -13.5 * b.t_radvel / powi(b.t_dist, 8);
            R hs = get_sigma(h.id, b.id, h.evo_type, h.evo_param, h.t_sigma);
            R ps = get_sigma(b.id, h.id, b.evo_type, b.evo_param, b.t_sigma);
            R term1 = m2 * powi(h.radius, 10) * hs;
            R term2 = host_mass_2 * powi(b.radius, 10) * ps;
            b.t_radial_diss_pm = factor1 * term2;
            R diss = b.t_radial_diss_pm + factor1 * term1;
            b.t_radial = cons + diss;
        }
    }
    // constant_time_lag.rs:309-332
    V3<R> tides_force(const Body<R>& h, const Body<R>& b) {
        R f3 = b.t_radial + (b.t_orth_star + b.t_orth_planet) * b.t_radvel / b.t_dist;
        V3<R> F;
        F.x = f3 * b.t_pos.x / b.t_dist
            + b.t_orth_star / b.t_dist * (h.spin.y * b.t_pos.z - h.spin.z * b.t_pos.y - b.t_vel.x)
            + b.t_orth_planet / b.t_dist * (b.spin.y * b.t_pos.z - b.spin.z * b.t_pos.y - b.t_vel.x);
        F.y = f3 * b.t_pos.y / b.t_dist
            + b.t_orth_star / b.t_dist * (h.spin.z * b.t_pos.x - h.spin.x * b.t_pos.z - b.t_vel.y)
            + b.t_orth_planet / b.t_dist * (b.spin.z * b.t_pos.x - b.spin.x * b.t_pos.z - b.t_vel.y);
        F.z = f3 * b.t_pos.z / b.t_dist
            + b.t_orth_star / b.t_dist * (h.spin.x * b.t_pos.y - h.spin.y * b.t_pos.x - b.t_vel.z)
            + b.t_orth_planet / b.t_dist * (b.spin.x * b.t_pos.y - b.spin.y * b.t_pos.x - b.t_vel.z);
        return F;
    }
    // tides/common.rs:312-345 (the Kaula branch :347-369 is out of scope)
    void tides_acceleration(int host) {
        Body<R>& h = p[host];
        R factor2 = 1. / h.mass;
        V3<R> sum{0., 0., 0.};
        for (int i = 0; i < n; i++) {
            if (i == host) continue;
            Body<R>& b = p[i];
            if (b.t_role != PB200_ROLE_ORBITING) continue;
            V3<R> F = tides_force(h, b);
            R factor1 = 1. / b.mass;
            sum.x = sum.x + F.x; sum.y = sum.y + F.y; sum.z = sum.z + F.z;
            b.t_acc.x = factor1 * F.x; b.t_acc.y = factor1 * F.y; b.t_acc.z = factor1 * F.z;
        }
        h.t_acc.x = -1.0 * factor2 * sum.x;
        h.t_acc.y = -1.0 * factor2 * sum.y;
        h.t_acc.z = -1.0 * factor2 * sum.z;
    }
    // constant_time_lag.rs:171-204
    V3<R> tides_torque(const Body<R>& h, const Body<R>& b, bool central) {
        V3<R> ref; R orth, rs;
        if (!central) { ref = b.spin; rs = b.t_rs_planet; orth = b.t_orth_planet; }
        else { ref = h.spin; rs = b.t_rs_star; orth = b.t_orth_star; }
        R d = b.t_dist;
        V3<R> N;
        N.x = orth * (d * ref.x - rs * b.t_pos.x / d - 1.0 / d * (b.t_pos.y * b.t_vel.z - b.t_pos.z * b.t_vel.y));
        N.y = orth * (d * ref.y - rs * b.t_pos.y / d - 1.0 / d * (b.t_pos.z * b.t_vel.x - b.t_pos.x * b.t_vel.z));
        N.z = orth * (d * ref.z - rs * b.t_pos.z / d - 1.0 / d * (b.t_pos.x * b.t_vel.y - b.t_pos.y * b.t_vel.x));
        return N;
    }
    // tides/common.rs:223-261
    void tides_dangular_momentum_dt(int host) {
        Body<R>& h = p[host];
        R factor = -1.0;
        for (int i = 0; i < n; i++) {
            if (i == host) continue;
            Body<R>& b = p[i];
            if (b.t_role != PB200_ROLE_ORBITING) continue;
            V3<R> N = tides_torque(h, b, false);
            b.t_dL.x = factor * N.x; b.t_dL.y = factor * N.y; b.t_dL.z = factor * N.z;
        }
        V3<R> s{0., 0., 0.};
        for (int i = 0; i < n; i++) {
            if (i == host) continue;
            if (h.t_role != PB200_ROLE_CENTRAL) continue;
            // DEVIATION D1 (documented in DESIGN.md): the reference sums over EVERY non-host particle, and a
            // tides-Disabled one contributes 0 * (0*0/0) = NaN (its tidal distance is never set), which poisons the
            // host spin and then hangs the Kepler bisection. Disabled bodies are skipped here (their term is 0 * x).
            if (p[i].t_role != PB200_ROLE_ORBITING) continue;
            V3<R> N = tides_torque(h, p[i], true);
            s.x = s.x + factor * N.x; s.y = s.y + factor * N.y; s.z = s.z + factor * N.z;
        }
        h.t_dL = s;
    }
    // tides/common.rs:263-279
    void tides_denergy_dt(int host) {
        for (int i = 0; i < n; i++) {
            if (i == host) continue;
            Body<R>& b = p[i];
            if (b.t_role != PB200_ROLE_ORBITING) continue;
            R factor2 = b.t_orth_planet / b.t_dist;
            b.t_denergy = -((1.0 / b.t_dist * (b.t_radial_diss_pm + factor2 * b.t_radvel))
                              * (b.t_pos.x * b.t_vel.x + b.t_pos.y * b.t_vel.y + b.t_pos.z * b.t_vel.z)
                          + factor2
                              * ((b.spin.y * b.t_pos.z - b.spin.z * b.t_pos.y - b.t_vel.x) * b.t_vel.x
                               + (b.spin.z * b.t_pos.x - b.spin.x * b.t_pos.z - b.t_vel.y) * b.t_vel.y
                               + (b.spin.x * b.t_pos.y - b.spin.y * b.t_pos.x - b.t_vel.z) * b.t_vel.z))
                          - (b.t_dL.x * b.spin.x + b.t_dL.y * b.spin.y + b.t_dL.z * b.spin.z);
        }
    }

    // ------------------------------------------------------------------ oblate-spheroid rotational flattening
    // oblate_spheroid.rs:12-49
    void flat_orthogonal(int host) {
        Body<R>& h = p[host];
        for (int pass = 0; pass < 2; pass++) {
            bool central = pass == 0;
            for (int i = 0; i < n; i++) {
                if (i == host) continue;
                Body<R>& b = p[i];
                if (b.f_role != PB200_ROLE_ORBITING) continue;
                R host_k2 = h.f_role == PB200_ROLE_CENTRAL ? h.f_k2 : R(0.);
                if (central) {
                    b.f_factor_star = b.mass * host_k2 * h.norm_spin2 * powi(h.radius, 5) / 6.;
                    b.f_orth_star = -6. * b.f_factor_star * b.f_rs_star / (h.norm_spin2 * powi(b.f_dist, 5));
                } else {
                    b.f_factor_planet = h.mass * b.f_k2 * b.norm_spin2 * powi(b.radius, 5) / 6.;
                    b.f_orth_planet = -6. * b.f_factor_planet * b.f_rs_planet / (b.norm_spin2 * powi(b.f_dist, 5));
                }
            }
        }
    }
    // oblate_spheroid.rs:51-60
    void flat_radial(int host) {
        Body<R>& h = p[host];
        for (int i = 0; i < n; i++) {
            if (i == host) continue;
            Body<R>& b = p[i];
            if (b.f_role != PB200_ROLE_ORBITING) continue;
            b.f_radial = -3. / powi(b.f_dist, 5) * (b.f_factor_planet + b.f_factor_star)
                + 15. / powi(b.f_dist, 7) * (b.f_factor_star * b.f_rs_star * b.f_rs_star / h.norm_spin2
                                             + b.f_factor_planet * b.f_rs_planet * b.f_rs_planet / b.norm_spin2);
        }
    }
    // rotational_flattening/common.rs:203-237 + oblate_spheroid.rs:83-97
    void flat_acceleration(int host) {
        Body<R>& h = p[host];
        R factor2 = 1. / h.mass;
        V3<R> sum{0., 0., 0.};
        for (int i = 0; i < n; i++) {
            if (i == host) continue;
            Body<R>& b = p[i];
            if (b.f_role != PB200_ROLE_ORBITING) continue;
            V3<R> F;
            F.x = b.f_radial * b.f_pos.x + b.f_orth_planet * b.spin.x + b.f_orth_star * h.spin.x;
            F.y = b.f_radial * b.f_pos.y + b.f_orth_planet * b.spin.y + b.f_orth_star * h.spin.y;
            F.z = b.f_radial * b.f_pos.z + b.f_orth_planet * b.spin.z + b.f_orth_star * h.spin.z;
            sum.x = sum.x + F.x; sum.y = sum.y + F.y; sum.z = sum.z + F.z;
            R factor1 = 1. / b.mass;
            b.f_acc.x = factor1 * F.x; b.f_acc.y = factor1 * F.y; b.f_acc.z = factor1 * F.z;
        }
        h.f_acc.x = -1.0 * factor2 * sum.x;
        h.f_acc.y = -1.0 * factor2 * sum.y;
        h.f_acc.z = -1.0 * factor2 * sum.z;
    }
    // oblate_spheroid.rs:62-81
    V3<R> flat_torque(const Body<R>& h, const Body<R>& b, bool central) {
        V3<R> ref = h.spin; R orth;
        if (!central) { ref = b.spin; orth = b.f_orth_planet; } else orth = b.f_orth_star;
        V3<R> N;
        N.x = orth * (b.f_pos.y * ref.z - b.f_pos.z * ref.y);
        N.y = orth * (b.f_pos.z * ref.x - b.f_pos.x * ref.z);
        N.z = orth * (b.f_pos.x * ref.y - b.f_pos.y * ref.x);
        return N;
    }
    // rotational_flattening/common.rs:165-201
    void flat_dangular_momentum_dt(int host) {
        Body<R>& h = p[host];
        R factor = -1.0;
        for (int i = 0; i < n; i++) {
            if (i == host) continue;
            Body<R>& b = p[i];
            if (b.f_role != PB200_ROLE_ORBITING) continue;
            V3<R> N = flat_torque(h, b, false);
            b.f_dL.x = factor * N.x; b.f_dL.y = factor * N.y; b.f_dL.z = factor * N.z;
        }
        V3<R> s{0., 0., 0.};
        for (int i = 0; i < n; i++) {
            if (i == host) continue;
            if (h.f_role != PB200_ROLE_CENTRAL) continue;
            V3<R> N = flat_torque(h, p[i], true);
            s.x = s.x + factor * N.x; s.y = s.y + factor * N.y; s.z = s.z + factor * N.z;
        }
        h.f_dL = s;
    }

    // ------------------------------------------------------------------ GR Kidder1995 (general_relativity.rs:177-456)
    void gr_kidder(int host) {
        Body<R>& h = p[host];
        // 1PN :187-239
        V3<R> sum{0., 0., 0.};
        for (int i = 0; i < n; i++) {
            if (i == host) continue;
            Body<R>& b = p[i];
            if (b.g_role != PB200_ROLE_ORBITING) continue;
            R mg = h.mass_g + b.mass_g;
            R d2 = powi(b.g_dist, 2);
            R rv2 = powi(b.g_radvel, 2);
            R radial = -mg / (d2 * SPEED_OF_LIGHT_2)
                * ((1.0 + 3.0 * b.g_factor) * b.g_normv2 - 2.0 * (2.0 + b.g_factor) * mg / b.g_dist - 1.5 * b.g_factor * rv2);
            R orth = mg / (d2 * SPEED_OF_LIGHT_2) * 2.0 * (2.0 - b.g_factor) * b.g_radvel * b.g_normv;
            R ax = radial * b.g_pos.x / b.g_dist + orth * b.g_vel.x / b.g_normv;
            R ay = radial * b.g_pos.y / b.g_dist + orth * b.g_vel.y / b.g_normv;
            R az = radial * b.g_pos.z / b.g_dist + orth * b.g_vel.z / b.g_normv;
            sum.x = sum.x + b.mass / h.mass * ax;
            sum.y = sum.y + b.mass / h.mass * ay;
            sum.z = sum.z + b.mass / h.mass * az;
            b.g_acc.x = ax; b.g_acc.y = ay; b.g_acc.z = az;
        }
        h.g_acc.x = -1.0 * sum.x; h.g_acc.y = -1.0 * sum.y; h.g_acc.z = -1.0 * sum.z;
        // 2PN :241-298
        sum = {0., 0., 0.};
        for (int i = 0; i < n; i++) {
            if (i == host) continue;
            Body<R>& b = p[i];
            if (b.g_role != PB200_ROLE_ORBITING) continue;
            R mg = h.mass_g + b.mass_g;
            R d2 = powi(b.g_dist, 2);
            R v2 = b.g_normv2;
            R v4 = powi(v2, 2);
            R rv2 = powi(b.g_radvel, 2);
            R rv4 = powi(rv2, 2);
            R f = b.g_factor;
            R f2 = powi(f, 2);
            R radial = -mg / (d2 * SPEED_OF_LIGHT_2)
                * (3.0 / 4.0 * (12.0 + 29.0 * f) * (powi(mg, 2) / d2)
                   + f * (3.0 - 4.0 * f) * v4
                   + 15.0 / 8.0 * f * (1.0 - 3.0 * f) * rv4
                   - 3.0 / 2.0 * f * (3.0 - 4.0 * f) * rv2 * v2
                   - 0.5 * f * (13.0 - 4.0 * f) * (mg / b.g_dist) * v2
                   - (2.0 + 25.0 * f + 2.0 * f2) * (mg / b.g_dist) * rv2);
            R orth = -mg / (d2 * SPEED_OF_LIGHT_2) * (-0.5) * b.g_radvel
                * (f * (15.0 + 4.0 * f) * v2 - (4.0 + 41.0 * f + 8.0 * f2) * (mg / b.g_dist) - 3.0 * f * (3.0 + 2.0 * f) * rv2);
            R ax = radial * b.g_pos.x / b.g_dist + orth * b.g_vel.x;
            R ay = radial * b.g_pos.y / b.g_dist + orth * b.g_vel.y;
            R az = radial * b.g_pos.z / b.g_dist + orth * b.g_vel.z;
            sum.x = sum.x + b.mass / h.mass * ax;
            sum.y = sum.y + b.mass / h.mass * ay;
            sum.z = sum.z + b.mass / h.mass * az;
            b.g_acc.x = b.g_acc.x + ax; b.g_acc.y = b.g_acc.y + ay; b.g_acc.z = b.g_acc.z + az;
        }
        h.g_acc.x = h.g_acc.x + -1.0 * sum.x; h.g_acc.y = h.g_acc.y + -1.0 * sum.y; h.g_acc.z = h.g_acc.z + -1.0 * sum.z;
        // 1.5PN spin-orbit :300-456
        V3<R> Ls{h.moi * h.spin.x, h.moi * h.spin.y, h.moi * h.spin.z};
        sum = {0., 0., 0.};
        h.g_dL = {0., 0., 0.};
        for (int i = 0; i < n; i++) {
            if (i == host) continue;
            Body<R>& b = p[i];
            if (b.g_role != PB200_ROLE_ORBITING) continue;
            R msum = h.mass + b.mass;
            R mdiff = h.mass - b.mass;
            R mfac = mdiff / msum;
            V3<R> Lp{b.moi * b.spin.x, b.moi * b.spin.y, b.moi * b.spin.z};
            V3<R> nn{b.g_pos.x / b.g_dist, b.g_pos.y / b.g_dist, b.g_pos.z / b.g_dist};
            R msx = mfac * msum * (Lp.x / b.mass - Ls.x / h.mass);
            R msy = mfac * msum * (Lp.y / b.mass - Ls.y / h.mass);
            R msz = mfac * msum * (Lp.z / b.mass - Ls.z / h.mass);
            R e1x = 6. * nn.x * ((nn.y * b.g_vel.z - nn.z * b.g_vel.y) * (2. * (Ls.x + Lp.x) + msx));
            R e1y = 6. * nn.y * ((nn.z * b.g_vel.x - nn.x * b.g_vel.z) * (2. * (Ls.y + Lp.y) + msy));
            R e1z = 6. * nn.z * ((nn.x * b.g_vel.y - nn.y * b.g_vel.x) * (2. * (Ls.z + Lp.z) + msz));
            V3<R> e7{7. * (Ls.x + Lp.x) + 3. * msx, 7. * (Ls.y + Lp.y) + 3. * msy, 7. * (Ls.z + Lp.z) + 3. * msz};
            R e2x = b.g_vel.y * e7.z - b.g_vel.z * e7.y;
            R e2y = b.g_vel.z * e7.x - b.g_vel.x * e7.z;
            R e2z = b.g_vel.x * e7.y - b.g_vel.y * e7.x;
            V3<R> e3s{3. * (Ls.x + Lp.x) + msx, 3. * (Ls.y + Lp.y) + msy, 3. * (Ls.z + Lp.z) + msz};
            R e3x = 3. * b.g_radvel * (nn.y * e3s.z - nn.z * e3s.y);
            R e3y = 3. * b.g_radvel * (nn.z * e3s.x - nn.x * e3s.z);
            R e3z = 3. * b.g_radvel * (nn.x * e3s.y - nn.y * e3s.x);
            R fa = G / SPEED_OF_LIGHT_2;
            R ax = fa * (e1x - e2x + e3x);
            R ay = fa * (e1y - e2y + e3y);
            R az = fa * (e1z - e2z + e3z);
            sum.x = sum.x + b.mass / h.mass * ax;
            sum.y = sum.y + b.mass / h.mass * ay;
            sum.z = sum.z + b.mass / h.mass * az;
            b.g_acc.x = b.g_acc.x + ax; b.g_acc.y = b.g_acc.y + ay; b.g_acc.z = b.g_acc.z + az;
            // Kidder 1995 eq. 2.4a
            R mu = (h.mass * b.mass) / msum;
            V3<R> Lo{mu * (b.g_pos.y * b.g_vel.z - b.g_pos.z * b.g_vel.y),
                     mu * (b.g_pos.z * b.g_vel.x - b.g_pos.x * b.g_vel.z),
                     mu * (b.g_pos.x * b.g_vel.y - b.g_pos.y * b.g_vel.x)};
            R fm = 2. + 3. / 2. * b.mass / h.mass;
            R a1x = fm * (Lo.y * Ls.z - Lo.z * Ls.y);
            R a1y = fm * (Lo.z * Ls.x - Lo.x * Ls.z);
            R a1z = fm * (Lo.x * Ls.y - Lo.y * Ls.x);
            R a2x = Lp.y * Ls.z - Lp.z * Ls.y;
            R a2y = Lp.z * Ls.x - Lp.x * Ls.z;
            R a2z = Lp.x * Ls.y - Lp.y * Ls.x;
            R sp = nn.x * Lp.x + nn.y * Lp.y + nn.z * Lp.z;
            R a3x = 3. * sp * (nn.y * Ls.z - nn.z * Ls.y);
            R a3y = 3. * sp * (nn.z * Ls.x - nn.x * Ls.z);
            R a3z = 3. * sp * (nn.x * Ls.y - nn.y * Ls.x);
            h.g_dL.x = h.g_dL.x + fa * (a1x - a2x + a3x);
            h.g_dL.y = h.g_dL.y + fa * (a1y - a2y + a3y);
            h.g_dL.z = h.g_dL.z + fa * (a1z - a2z + a3z);
            // Kidder 1995 eq. 2.4b
            fm = 2. + 3. / 2. * h.mass / b.mass;
            R b1x = fm * (Lo.y * Lp.z - Lo.z * Lp.y);
            R b1y = fm * (Lo.z * Lp.x - Lo.x * Lp.z);
            R b1z = fm * (Lo.x * Lp.y - Lo.y * Lp.x);
            R b2x = Ls.y * Lp.z - Ls.z * Lp.y;
            R b2y = Ls.z * Lp.x - Ls.x * Lp.z;
            R b2z = Ls.x * Lp.y - Ls.y * Lp.x;
            R ss = nn.x * Ls.x + nn.y * Ls.y + nn.z * Ls.z;
            R b3x = 3. * ss * (nn.y * Lp.z - nn.z * Lp.y);
            R b3y = 3. * ss * (nn.z * Lp.x - nn.x * Lp.z);
            R b3z = 3. * ss * (nn.x * Lp.y - nn.y * Lp.x);
            b.g_dL.x = fa * (b1x - b2x + b3x);
            b.g_dL.y = fa * (b1y - b2y + b3y);
            b.g_dL.z = fa * (b1z - b2z + b3z);
        }
        h.g_acc.x = h.g_acc.x + -1.0 * sum.x; h.g_acc.y = h.g_acc.y + -1.0 * sum.y; h.g_acc.z = h.g_acc.z + -1.0 * sum.z;
    }

    // ------------------------------------------------------------------ shared Newtonian helper (general_relativity.rs:641-678)
    // `others` = non-host indices in array order (particles_left ++ particles_right)
    int others(int host, int* idx) const {
        int m = 0;
        for (int i = 0; i < n; i++) if (i != host) idx[m++] = i;
        return m;
    }
    void gr_newtonian(int host, Ignore ign, V3<R>& ha, V3<R>* a) {
        int idx[MAXP]; int m = others(host, idx);
        Body<R>& h = p[host];
        ha = h.iacc;
        for (int k = 0; k < m; k++) a[k] = p[idx[k]].iacc;
        if (ign == WHFastOne || ign == WHFastTwo) {
            int cnt = ign == WHFastOne ? 1 : m;
            if (cnt > m) cnt = m;
            for (int k = 0; k < cnt; k++) {
                Body<R>& b = p[idx[k]];
                // Q9: host inertial position minus the particle's HELIOCENTRIC GR position
                R dx = h.ipos.x - b.g_pos.x, dy = h.ipos.y - b.g_pos.y, dz = h.ipos.z - b.g_pos.z;
                R r2 = powi(dx, 2) + powi(dy, 2) + powi(dz, 2);
                R r = o_sqrt(r2);
                R prefac = G / (r2 * r);
                R pms = prefac * h.mass;
                R pmp = prefac * b.mass;
                ha.x = ha.x - pmp * dx; ha.y = ha.y - pmp * dy; ha.z = ha.z - pmp * dz;
                a[k].x = a[k].x + pms * dx; a[k].y = a[k].y + pms * dy; a[k].z = a[k].z + pms * dz;
            }
        }
    }
    // ------------------------------------------------------------------ GR Anderson1975 (general_relativity.rs:461-636)
    void gr_anderson(int host, Ignore ign) {
        int idx[MAXP]; int m = others(host, idx);
        Body<R>& h = p[host];
        V3<R> ha, a[MAXP];
        gr_newtonian(host, ign, ha, a);
        // inertial -> Jacobi (:539-602)
        V3<R> jp[MAXP], jv[MAXP], ja[MAXP];
        for (int k = 0; k < MAXP; k++) { jp[k] = {0., 0., 0.}; jv[k] = {0., 0., 0.}; ja[k] = {0., 0., 0.}; }
        R eta = h.mass;
        R sx = eta * h.ipos.x, sy = eta * h.ipos.y, sz = eta * h.ipos.z;
        R svx = eta * h.ivel.x, svy = eta * h.ivel.y, svz = eta * h.ivel.z;
        R sax = eta * ha.x, say = eta * ha.y, saz = eta * ha.z;
        for (int k = 0; k < m; k++) {
            Body<R>& b = p[idx[k]];
            if (b.g_role != PB200_ROLE_ORBITING) continue;
            R ei = 1. / eta;
            eta = eta + b.mass;
            R pme = eta * ei;
            jp[k].x = b.ipos.x - sx * ei; jp[k].y = b.ipos.y - sy * ei; jp[k].z = b.ipos.z - sz * ei;
            jv[k].x = b.ivel.x - svx * ei; jv[k].y = b.ivel.y - svy * ei; jv[k].z = b.ivel.z - svz * ei;
            ja[k].x = a[k].x - sax * ei; ja[k].y = a[k].y - say * ei; ja[k].z = a[k].z - saz * ei;
            sx = sx * pme + b.mass * jp[k].x; sy = sy * pme + b.mass * jp[k].y; sz = sz * pme + b.mass * jp[k].z;
            svx = svx * pme + b.mass * jv[k].x; svy = svy * pme + b.mass * jv[k].y; svz = svz * pme + b.mass * jv[k].z;
            sax = sax * pme + b.mass * ja[k].x; say = say * pme + b.mass * ja[k].y; saz = saz * pme + b.mass * ja[k].z;
        }
        R jacobi_star_mass = eta;
        R mu = h.mass_g;
        for (int k = 0; k < m; k++) {
            Body<R>& b = p[idx[k]];
            if (b.g_role != PB200_ROLE_ORBITING) continue;
            V3<R> vi = jv[k];
            R vi2 = powi(jv[k].x, 2) + powi(jv[k].y, 2) + powi(jv[k].z, 2);
            R ri = o_sqrt(powi(jp[k].x, 2) + powi(jp[k].y, 2) + powi(jp[k].z, 2));
            R fa = (0.5 * vi2 + 3. * mu / ri) / SPEED_OF_LIGHT_2;
            V3<R> old{0., 0., 0.};
            for (int q = 0; q < 10; q++) {
                old = vi;
                vi.x = jv[k].x / (1. - fa); vi.y = jv[k].y / (1. - fa); vi.z = jv[k].z / (1. - fa);
                vi2 = vi.x * vi.x + vi.y * vi.y + vi.z * vi.z;
                fa = (0.5 * vi2 + 3. * mu / ri) / SPEED_OF_LIGHT_2;
                R dvx = vi.x - old.x, dvy = vi.y - old.y, dvz = vi.z - old.z;
                if (o_val((dvx * dvx + dvy * dvy + dvz * dvz) / vi2) < DBL_EPSILON_2) break;
            }
            R fb = (mu / ri - 1.5 * vi2) * mu / (ri * ri * ri) / SPEED_OF_LIGHT_2;
            R rdotrdot = jp[k].x * jv[k].x + jp[k].y * jv[k].y + jp[k].z * jv[k].z;
            V3<R> vidot{ja[k].x + fb * jp[k].x, ja[k].y + fb * jp[k].y, ja[k].z + fb * jp[k].z};
            R vdotvdot = vi.x * vidot.x + vi.y * vidot.y + vi.z * vidot.z;
            R fd = (vdotvdot - 3. * mu / (ri * ri * ri) * rdotrdot) / SPEED_OF_LIGHT_2;
            ja[k].x = fb * (1. - fa) * jp[k].x - fa * ja[k].x - fd * vi.x;
            ja[k].y = fb * (1. - fa) * jp[k].y - fa * ja[k].y - fd * vi.y;
            ja[k].z = fb * (1. - fa) * jp[k].z - fa * ja[k].z - fd * vi.z;
        }
        // Jacobi -> inertial accelerations (:604-636), star Jacobi acceleration = 0
        V3<R> pa[MAXP];
        for (int k = 0; k < MAXP; k++) pa[k] = {0., 0., 0.};
        eta = jacobi_star_mass;
        R s_ax = eta * 0., s_ay = eta * 0., s_az = eta * 0.;
        for (int k = m - 1; k >= 0; k--) {
            Body<R>& b = p[idx[k]];
            if (b.g_role != PB200_ROLE_ORBITING) continue;
            R ei = 1. / eta;
            s_ax = (s_ax - b.mass * ja[k].x) * ei;
            s_ay = (s_ay - b.mass * ja[k].y) * ei;
            s_az = (s_az - b.mass * ja[k].z) * ei;
            pa[k].x = ja[k].x + s_ax; pa[k].y = ja[k].y + s_ay; pa[k].z = ja[k].z + s_az;
            eta = eta - b.mass;
            s_ax = s_ax * eta; s_ay = s_ay * eta; s_az = s_az * eta;
        }
        R mtot_i = 1. / eta;
        for (int k = 0; k < m; k++) {
            Body<R>& b = p[idx[k]];
            if (b.g_role != PB200_ROLE_ORBITING) continue;
            b.g_acc = pa[k];
        }
        h.g_acc.x = s_ax * mtot_i; h.g_acc.y = s_ay * mtot_i; h.g_acc.z = s_az * mtot_i;
    }
    // ------------------------------------------------------------------ GR Newhall1983 (general_relativity.rs:683-895)
    void gr_newhall(int host, Ignore ign) {
        int idx[MAXP]; int m = others(host, idx);
        V3<R> ha, an[MAXP];
        gr_newtonian(host, ign, ha, an);
        // body order: 0 = host, 1.. = others in array order
        int M = m + 1;
        Body<R>* q[MAXP];
        q[0] = &p[host];
        for (int k = 0; k < m; k++) q[k + 1] = &p[idx[k]];
        V3<R> newt[MAXP];
        newt[0] = ha;
        for (int k = 0; k < m; k++) newt[k + 1] = an[k];
        R rs[MAXP][MAXP];
        V3<R> drs[MAXP][MAXP];
        for (int i = 0; i < MAXP; i++) for (int j = 0; j < MAXP; j++) { rs[i][j] = 0.; drs[i][j] = {0., 0., 0.}; }
        auto enabled = [&](int i, int j) { return q[i]->g_role != PB200_ROLE_DISABLED || q[j]->g_role != PB200_ROLE_DISABLED; };
        for (int i = 0; i < M; i++)
            for (int j = 0; j < M; j++) {
                if (j == i || !enabled(i, j)) continue;
                drs[i][j].x = q[i]->ipos.x - q[j]->ipos.x;
                drs[i][j].y = q[i]->ipos.y - q[j]->ipos.y;
                drs[i][j].z = q[i]->ipos.z - q[j]->ipos.z;
                rs[i][j] = o_sqrt(powi(drs[i][j].x, 2) + powi(drs[i][j].y, 2) + powi(drs[i][j].z, 2));
            }
        V3<R> a_const[MAXP], a_new[MAXP];
        for (int i = 0; i < MAXP; i++) { a_const[i] = {0., 0., 0.}; a_new[i] = {0., 0., 0.}; }
        for (int i = 0; i < M; i++) {
            R cx = 0., cy = 0., cz = 0.;
            for (int j = 0; j < M; j++) {
                if (j == i || !enabled(i, j)) continue;
                R dxij = drs[i][j].x, dyij = drs[i][j].y, dzij = drs[i][j].z;
                R rij2 = powi(rs[i][j], 2);
                R rij3 = rij2 * rs[i][j];
                R a1 = 0.;
                for (int k = 0; k < M; k++) if (k != i) a1 = a1 + (4. / (SPEED_OF_LIGHT_2)) * G * q[k]->mass / rs[i][k];
                R a2 = 0.;
                for (int l = 0; l < M; l++) if (l != j) a2 = a2 + (1. / (SPEED_OF_LIGHT_2)) * G * q[l]->mass / rs[l][j];
                R vi2 = powi(q[i]->ivel.x, 2) + powi(q[i]->ivel.y, 2) + powi(q[i]->ivel.z, 2);
                R a3 = -vi2 / SPEED_OF_LIGHT_2;
                R vj2 = powi(q[j]->ivel.x, 2) + powi(q[j]->ivel.y, 2) + powi(q[j]->ivel.z, 2);
                R a4 = -2. * vj2 / SPEED_OF_LIGHT_2;
                R a5 = (4. / SPEED_OF_LIGHT_2) * (q[i]->ivel.x * q[j]->ivel.x + q[i]->ivel.y * q[j]->ivel.y + q[i]->ivel.z * q[j]->ivel.z);
                R a6_0 = dxij * q[j]->ivel.x + dyij * q[j]->ivel.y + dzij * q[j]->ivel.z;
                R a6 = (3. / (2. * SPEED_OF_LIGHT_2)) * powi(a6_0, 2) / rij2;
                R factor1 = a1 + a2 + a3 + a4 + a5 + a6;
                cx = cx + G * q[j]->mass * dxij * factor1 / rij3;
                cy = cy + G * q[j]->mass * dyij * factor1 / rij3;
                cz = cz + G * q[j]->mass * dzij * factor1 / rij3;
                R dvx = q[i]->ivel.x - q[j]->ivel.x, dvy = q[i]->ivel.y - q[j]->ivel.y, dvz = q[i]->ivel.z - q[j]->ivel.z;
                R factor2 = dxij * (4. * q[i]->ivel.x - 3. * q[j]->ivel.x) + dyij * (4. * q[i]->ivel.y - 3. * q[j]->ivel.y)
                          + dzij * (4. * q[i]->ivel.z - 3. * q[j]->ivel.z);
                cx = cx + G * q[j]->mass * factor2 * dvx / rij3 / SPEED_OF_LIGHT_2;
                cy = cy + G * q[j]->mass * factor2 * dvy / rij3 / SPEED_OF_LIGHT_2;
                cz = cz + G * q[j]->mass * factor2 * dvz / rij3 / SPEED_OF_LIGHT_2;
            }
            a_const[i] = {cx, cy, cz};
        }
        const double dev_limit = 1.0e-30;
        for (int k = 0; k < 10; k++) {
            V3<R> a_old[MAXP];
            for (int i = 0; i < MAXP; i++) a_old[i] = a_new[i];
            for (int i = 0; i < M; i++) {
                R nx = 0., ny = 0., nz = 0.;
                for (int j = 0; j < M; j++) {
                    if (j == i || !enabled(i, j)) continue;
                    R dxij = drs[i][j].x, dyij = drs[i][j].y, dzij = drs[i][j].z;
                    R rij = rs[i][j];
                    R rij2 = powi(rij, 2);
                    R rij3 = rij2 * rij;
                    R mj = q[j]->mass;
                    nx = nx + (G * mj * dxij / rij3) * (dxij * (newt[j].x + a_old[j].x) + dyij * (newt[j].y + a_old[j].y) + dzij * (newt[j].z + a_old[j].z)) / (2. * SPEED_OF_LIGHT_2)
                            + (7. / (2. * SPEED_OF_LIGHT_2)) * G * mj * (newt[j].x + a_old[j].x) / rij;
                    ny = ny + (G * mj * dyij / rij3) * (dxij * (newt[j].x + a_old[j].x) + dyij * (newt[j].y + a_old[j].y) + dzij * (newt[j].z + a_old[j].z)) / (2. * SPEED_OF_LIGHT_2)
                            + (7. / (2. * SPEED_OF_LIGHT_2)) * G * mj * (newt[j].y + a_old[j].y) / rij;
                    nz = nz + (G * mj * dzij / rij3) * (dxij * (newt[j].x + a_old[j].x) + dyij * (newt[j].y + a_old[j].y) + dzij * (newt[j].z + a_old[j].z)) / (2. * SPEED_OF_LIGHT_2)
                            + (7. / (2. * SPEED_OF_LIGHT_2)) * G * mj * (newt[j].z + a_old[j].z) / rij;
                }
                a_new[i].x = a_const[i].x + nx; a_new[i].y = a_const[i].y + ny; a_new[i].z = a_const[i].z + nz;
            }
            // Q8: the deviation test is inverted in the reference (:856-864)
            double maxdev = 0., dx = 0., dy = 0., dz = 0.;
            for (int i = 0; i < M; i++) {
                if (q[i]->g_role == PB200_ROLE_DISABLED) continue;
                double nx = o_val(a_new[i].x), ny = o_val(a_new[i].y), nz = o_val(a_new[i].z);
                if (std::fabs(nx) < dev_limit) dx = std::fabs(nx - o_val(a_old[i].x)) / nx;
                if (std::fabs(ny) < dev_limit) dy = std::fabs(ny - o_val(a_old[i].y)) / ny;
                if (std::fabs(nz) < dev_limit) dz = std::fabs(nz - o_val(a_old[i].z)) / nz;
                if (dx > maxdev) maxdev = dx;
                if (dy > maxdev) maxdev = dy;
                if (dz > maxdev) maxdev = dz;
            }
            if (maxdev < dev_limit) break;
        }
        for (int k = 1; k < M; k++) if (q[k]->g_role == PB200_ROLE_ORBITING) q[k]->g_acc = a_new[k];
        q[0]->g_acc = a_new[0];
    }

    // ------------------------------------------------------------------ particles/universe.rs:428-614
    void calculate_additional_effects(double time, bool evolution, bool dL, bool acc, Ignore ign) {
        initialize(dL, acc);
        calculate_spin_and_evolving_quantities(time, evolution);
        if (dL && c_wind) calculate_wind_factor();
        if (c_tides || c_flat) {
            int host = h_tides;
            if (host >= 0 && host < n) {
                if ((dL && (c_tides || c_flat)) || (acc && (c_tides || c_disk || c_flat || c_gr))) {
                    tides_pair_sigma(host);
                    if (c_tides) tides_orthogonal(host);
                    if (c_flat) flat_orthogonal(host);
                    if (acc && (c_tides || c_flat)) {
                        if (c_tides) { tides_radial(host); tides_acceleration(host); }
                        if (c_flat) { flat_radial(host); flat_acceleration(host); }
                    }
                    if (dL && (c_tides || c_flat)) {
                        if (c_tides) tides_dangular_momentum_dt(host);
                        if (c_flat) flat_dangular_momentum_dt(host);
                    }
                }
            }
        }
        if (acc && c_gr) {
            int host = h_gr;
            if (host >= 0 && host < n && p[host].g_role == PB200_ROLE_CENTRAL) {
                if (gr_impl_of_host == PB200_GR_KIDDER1995) gr_kidder(host);
                else if (gr_impl_of_host == PB200_GR_ANDERSON1975) gr_anderson(host, ign);
                else if (gr_impl_of_host == PB200_GR_NEWHALL1983) gr_newhall(host, ign);
            }
        }
        if (dL) {
            if (c_tides || c_flat || (c_gr && gr_impl == PB200_GR_KIDDER1995) || c_wind) {
                for (int i = 0; i < n; i++) {
                    Body<R>& b = p[i];
                    b.dLdt.x = b.t_dL.x + b.f_dL.x + b.g_dL.x + b.w_dL.x;
                    b.dLdt.y = b.t_dL.y + b.f_dL.y + b.g_dL.y + b.w_dL.y;
                    b.dLdt.z = b.t_dL.z + b.f_dL.z + b.g_dL.z + b.w_dL.z;
                }
            }
        }
        if (acc) {
            for (int i = 0; i < n; i++) {
                Body<R>& b = p[i];
                if (c_tides) { b.iadd.x = b.iadd.x + b.t_acc.x; b.iadd.y = b.iadd.y + b.t_acc.y; b.iadd.z = b.iadd.z + b.t_acc.z; }
                if (c_disk) { b.iadd.x = b.iadd.x + b.d_acc.x; b.iadd.y = b.iadd.y + b.d_acc.y; b.iadd.z = b.iadd.z + b.d_acc.z; }
                if (c_flat) { b.iadd.x = b.iadd.x + b.f_acc.x; b.iadd.y = b.iadd.y + b.f_acc.y; b.iadd.z = b.iadd.z + b.f_acc.z; }
                if (c_gr) { b.iadd.x = b.iadd.x + b.g_acc.x; b.iadd.y = b.iadd.y + b.g_acc.y; b.iadd.z = b.iadd.z + b.g_acc.z; }
            }
        }
    }

    Ignore ignore_terms() const { return coord == PB200_COORD_JACOBI ? WHFastOne : WHFastTwo; }

    // ------------------------------------------------------------------ integrator/whfast.rs:322-466
    int last_midpoint_iterations = 0;
    void integrate_velocity_dependent_forces(double dt_, bool integrate_spin, bool evolution) {
        R dt = dt_;
        Ignore ign = ignore_terms();
        V3<R> v_orig[MAXP], L_orig[MAXP], v_final[MAXP], L_final[MAXP], v_prev[MAXP], L_prev[MAXP], dv[MAXP], dl[MAXP];
        for (int i = 0; i < n; i++) {
            v_orig[i] = p[i].ivel; L_orig[i] = p[i].L; v_final[i] = v_orig[i]; L_final[i] = L_orig[i];
            dv[i] = {0., 0., 0.}; dl[i] = {0., 0., 0.};
        }
        bool converged = false;
        int it = 0;
        for (int i = 0; i < IMPLICIT_MIDPOINT_MAX_ITER; i++) {
            it = i + 1;
            for (int k = 0; k < n; k++) { v_prev[k] = v_final[k]; L_prev[k] = L_final[k]; }
            inertial_to_heliocentric();
            calculate_additional_effects(current_time, evolution && i == 0, integrate_spin, true, ign);
            for (int k = 0; k < n; k++) {
                dv[k].x = dt * p[k].iadd.x - verr[k].x;
                dv[k].y = dt * p[k].iadd.y - verr[k].y;
                dv[k].z = dt * p[k].iadd.z - verr[k].z;
                v_final[k].x = v_orig[k].x + dv[k].x;
                v_final[k].y = v_orig[k].y + dv[k].y;
                v_final[k].z = v_orig[k].z + dv[k].z;
                if (integrate_spin) {
                    dl[k].x = dt * p[k].dLdt.x - lerr[k].x;
                    dl[k].y = dt * p[k].dLdt.y - lerr[k].y;
                    dl[k].z = dt * p[k].dLdt.z - lerr[k].z;
                    L_final[k].x = L_orig[k].x + dl[k].x;
                    L_final[k].y = L_orig[k].y + dl[k].y;
                    L_final[k].z = L_orig[k].z + dl[k].z;
                }
            }
            if (i >= IMPLICIT_MIDPOINT_MIN_ITER - 1) {
                // whfast.rs:424-451
                R fv2 = 0., dv2 = 0., fl2 = 0., dl2 = 0.;
                for (int k = 0; k < n; k++) {
                    R ddx = v_final[k].x - v_prev[k].x, ddy = v_final[k].y - v_prev[k].y, ddz = v_final[k].z - v_prev[k].z;
                    dv2 = dv2 + (powi(ddx, 2) + powi(ddy, 2) + powi(ddz, 2));
                    fv2 = fv2 + (powi(v_final[k].x, 2) + powi(v_final[k].y, 2) + powi(v_final[k].z, 2));
                    if (integrate_spin) {
                        R sx = L_final[k].x - L_prev[k].x, sy = L_final[k].y - L_prev[k].y, sz = L_final[k].z - L_prev[k].z;
                        dl2 = dl2 + (powi(sx, 2) + powi(sy, 2) + powi(sz, 2));
                        fl2 = fl2 + (powi(L_final[k].x, 2) + powi(L_final[k].y, 2) + powi(L_final[k].z, 2));
                    }
                }
                bool ok;
                if (integrate_spin) ok = o_val(dv2 / fv2) < DBL_EPSILON_2 && o_val(dl2 / fl2) < DBL_EPSILON_2;
                else ok = o_val(dv2 / fv2) < DBL_EPSILON_2;
                if (ok) { converged = true; break; }
            }
            // whfast.rs:453-466
            for (int k = 0; k < n; k++) {
                p[k].ivel.x = 0.5 * (v_orig[k].x + v_final[k].x);
                p[k].ivel.y = 0.5 * (v_orig[k].y + v_final[k].y);
                p[k].ivel.z = 0.5 * (v_orig[k].z + v_final[k].z);
                if (integrate_spin) {
                    p[k].L.x = 0.5 * (L_orig[k].x + L_final[k].x);
                    p[k].L.y = 0.5 * (L_orig[k].y + L_final[k].y);
                    p[k].L.z = 0.5 * (L_orig[k].z + L_final[k].z);
                }
            }
        }
        last_midpoint_iterations = it;
        if (!converged) warnings |= PB200_WARN_MIDPOINT_NOT_CONVERGED;
        for (int k = 0; k < n; k++) {
            p[k].ivel = v_final[k];
            verr[k].x = (p[k].ivel.x - v_orig[k].x) - dv[k].x;
            verr[k].y = (p[k].ivel.y - v_orig[k].y) - dv[k].y;
            verr[k].z = (p[k].ivel.z - v_orig[k].z) - dv[k].z;
            if (integrate_spin) {
                p[k].L = L_final[k];
                lerr[k].x = (p[k].L.x - L_orig[k].x) - dl[k].x;
                lerr[k].y = (p[k].L.y - L_orig[k].y) - dl[k].y;
                lerr[k].z = (p[k].L.z - L_orig[k].z) - dl[k].z;
            }
        }
    }

    // ------------------------------------------------------------------ coordinate transforms (whfast.rs:881-1155)
    void inertial_to_jacobi_posvel() {
        int idx[MAXP]; int m = others(h_mm, idx);
        Body<R>& star = p[h_mm];
        R eta = star.mass;
        R sx = eta * star.ipos.x, sy = eta * star.ipos.y, sz = eta * star.ipos.z;
        R svx = eta * star.ivel.x, svy = eta * star.ivel.y, svz = eta * star.ivel.z;
        for (int k = 0; k < m; k++) {
            Body<R>& b = p[idx[k]]; Alt<R>& a = alt[idx[k]];
            R ei = 1. / eta;
            eta = eta + b.mass;
            R pme = eta * ei;
            a.mass = b.mass; a.mass_g = b.mass_g;
            a.pos.x = b.ipos.x - sx * ei; a.pos.y = b.ipos.y - sy * ei; a.pos.z = b.ipos.z - sz * ei;
            a.vel.x = b.ivel.x - svx * ei; a.vel.y = b.ivel.y - svy * ei; a.vel.z = b.ivel.z - svz * ei;
            sx = sx * pme + b.mass * a.pos.x; sy = sy * pme + b.mass * a.pos.y; sz = sz * pme + b.mass * a.pos.z;
            svx = svx * pme + b.mass * a.vel.x; svy = svy * pme + b.mass * a.vel.y; svz = svz * pme + b.mass * a.vel.z;
        }
        R mtot = eta, mi = 1. / mtot;
        Alt<R>& s = alt[h_mm];
        s.mass = mtot;
        s.pos.x = sx * mi; s.pos.y = sy * mi; s.pos.z = sz * mi;
        s.vel.x = svx * mi; s.vel.y = svy * mi; s.vel.z = svz * mi;
    }
    void inertial_to_jacobi_acc() {
        int idx[MAXP]; int m = others(h_mm, idx);
        Body<R>& star = p[h_mm];
        R eta = star.mass;
        R sax = eta * star.iacc.x, say = eta * star.iacc.y, saz = eta * star.iacc.z;
        for (int k = 0; k < m; k++) {
            Body<R>& b = p[idx[k]]; Alt<R>& a = alt[idx[k]];
            R ei = 1. / eta;
            eta = eta + b.mass;
            R pme = eta * ei;
            a.acc.x = b.iacc.x - sax * ei; a.acc.y = b.iacc.y - say * ei; a.acc.z = b.iacc.z - saz * ei;
            sax = sax * pme + b.mass * a.acc.x; say = say * pme + b.mass * a.acc.y; saz = saz * pme + b.mass * a.acc.z;
        }
        R mi = 1. / eta;
        Alt<R>& s = alt[h_mm];
        s.acc.x = sax * mi; s.acc.y = say * mi; s.acc.z = saz * mi;
    }
    void inertial_to_dh_posvel() {
        int idx[MAXP]; int m = others(h_mm, idx);
        Body<R>& star = p[h_mm];
        Alt<R>& s = alt[h_mm];
        s.pos = {0., 0., 0.}; s.vel = {0., 0., 0.}; s.mass = 0.; s.mass_g = 0.;
        // iter::once(star).chain(left).chain(right)
        for (int k = -1; k < m; k++) {
            Body<R>& b = k < 0 ? star : p[idx[k]];
            s.pos.x = s.pos.x + b.ipos.x * b.mass; s.pos.y = s.pos.y + b.ipos.y * b.mass; s.pos.z = s.pos.z + b.ipos.z * b.mass;
            s.vel.x = s.vel.x + b.ivel.x * b.mass; s.vel.y = s.vel.y + b.ivel.y * b.mass; s.vel.z = s.vel.z + b.ivel.z * b.mass;
            s.mass = s.mass + b.mass; s.mass_g = s.mass_g + b.mass_g;
        }
        R mtot = s.mass;
        s.pos.x = s.pos.x / mtot; s.pos.y = s.pos.y / mtot; s.pos.z = s.pos.z / mtot;
        s.vel.x = s.vel.x / mtot; s.vel.y = s.vel.y / mtot; s.vel.z = s.vel.z / mtot;
        for (int k = 0; k < m; k++) {
            Body<R>& b = p[idx[k]]; Alt<R>& a = alt[idx[k]];
            a.pos.x = b.ipos.x - star.ipos.x; a.pos.y = b.ipos.y - star.ipos.y; a.pos.z = b.ipos.z - star.ipos.z;
            a.vel.x = b.ivel.x - s.vel.x; a.vel.y = b.ivel.y - s.vel.y; a.vel.z = b.ivel.z - s.vel.z;
            a.mass = b.mass; a.mass_g = b.mass_g;
            if (coord == PB200_COORD_WHDS) {
                R f = (star.mass + b.mass) / star.mass;
                a.vel.x = a.vel.x * f; a.vel.y = a.vel.y * f; a.vel.z = a.vel.z * f;
            }
        }
    }
    void inertial_to_alternative_posvel() {
        if (coord == PB200_COORD_JACOBI) inertial_to_jacobi_posvel(); else inertial_to_dh_posvel();
    }
    void jacobi_to_inertial_posvel() {
        int idx[MAXP]; int m = others(h_mm, idx);
        Alt<R>& s = alt[h_mm];
        R eta = s.mass;
        R sx = eta * s.pos.x, sy = eta * s.pos.y, sz = eta * s.pos.z;
        R svx = eta * s.vel.x, svy = eta * s.vel.y, svz = eta * s.vel.z;
        for (int k = m - 1; k >= 0; k--) {
            Body<R>& b = p[idx[k]]; Alt<R>& a = alt[idx[k]];
            R ei = 1. / eta;
            sx = (sx - b.mass * a.pos.x) * ei; sy = (sy - b.mass * a.pos.y) * ei; sz = (sz - b.mass * a.pos.z) * ei;
            svx = (svx - b.mass * a.vel.x) * ei; svy = (svy - b.mass * a.vel.y) * ei; svz = (svz - b.mass * a.vel.z) * ei;
            b.ipos.x = a.pos.x + sx; b.ipos.y = a.pos.y + sy; b.ipos.z = a.pos.z + sz;
            b.ivel.x = a.vel.x + svx; b.ivel.y = a.vel.y + svy; b.ivel.z = a.vel.z + svz;
            eta = eta - b.mass;
            sx = sx * eta; sy = sy * eta; sz = sz * eta; svx = svx * eta; svy = svy * eta; svz = svz * eta;
        }
        R mi = 1. / eta;
        Body<R>& star = p[h_mm];
        star.ipos.x = sx * mi; star.ipos.y = sy * mi; star.ipos.z = sz * mi;
        star.ivel.x = svx * mi; star.ivel.y = svy * mi; star.ivel.z = svz * mi;
    }
    void dh_to_inertial_posvel() {
        int idx[MAXP]; int m = others(h_mm, idx);
        Alt<R>& s = alt[h_mm];
        Body<R>& star = p[h_mm];
        // positions :1128-1155
        R mtot = s.mass;
        V3<R> np = s.pos;
        for (int k = 0; k < m; k++) {
            Body<R>& b = p[idx[k]]; Alt<R>& a = alt[idx[k]];
            np.x = np.x - a.pos.x * b.mass / mtot; np.y = np.y - a.pos.y * b.mass / mtot; np.z = np.z - a.pos.z * b.mass / mtot;
        }
        star.ipos = np;
        for (int k = 0; k < m; k++) {
            Body<R>& b = p[idx[k]]; Alt<R>& a = alt[idx[k]];
            b.ipos.x = a.pos.x + star.ipos.x; b.ipos.y = a.pos.y + star.ipos.y; b.ipos.z = a.pos.z + star.ipos.z;
        }
        // velocities :1090-1126
        R m0 = star.mass;
        V3<R> nv = s.vel;
        for (int k = 0; k < m; k++) {
            Body<R>& b = p[idx[k]]; Alt<R>& a = alt[idx[k]];
            if (coord == PB200_COORD_WHDS) {
                R f = (m0 + b.mass) / m0;
                b.ivel.x = a.vel.x / f + s.vel.x; b.ivel.y = a.vel.y / f + s.vel.y; b.ivel.z = a.vel.z / f + s.vel.z;
            } else {
                b.ivel.x = a.vel.x + s.vel.x; b.ivel.y = a.vel.y + s.vel.y; b.ivel.z = a.vel.z + s.vel.z;
            }
            R f;
            if (coord == PB200_COORD_WHDS) f = b.mass / (m0 + b.mass); else f = b.mass / m0;
            nv.x = nv.x - a.vel.x * f; nv.y = nv.y - a.vel.y * f; nv.z = nv.z - a.vel.z * f;
        }
        star.ivel = nv;
    }
    void alternative_to_inertial_posvel() {
        if (coord == PB200_COORD_JACOBI) jacobi_to_inertial_posvel(); else dh_to_inertial_posvel();
    }

    // ------------------------------------------------------------------ operators (whfast.rs:495-672)
    void jump_step(double dt_) {
        R dt = dt_;
        if (coord == PB200_COORD_JACOBI) return;
        int idx[MAXP]; int m = others(h_mm, idx);
        R m0 = p[h_mm].mass;
        R px = 0., py = 0., pz = 0.;
        if (coord == PB200_COORD_DEMOCRATIC_HELIOCENTRIC) {
            for (int k = 0; k < m; k++) {
                Body<R>& b = p[idx[k]]; Alt<R>& a = alt[idx[k]];
                px = px + b.mass * a.vel.x; py = py + b.mass * a.vel.y; pz = pz + b.mass * a.vel.z;
            }
            for (int k = 0; k < m; k++) {
                Alt<R>& a = alt[idx[k]];
                a.pos.x = a.pos.x + dt * px / m0; a.pos.y = a.pos.y + dt * py / m0; a.pos.z = a.pos.z + dt * pz / m0;
            }
        } else {
            for (int k = 0; k < m; k++) {
                Body<R>& b = p[idx[k]]; Alt<R>& a = alt[idx[k]];
                R f = m0 + b.mass;
                px = px + b.mass * a.vel.x / f; py = py + b.mass * a.vel.y / f; pz = pz + b.mass * a.vel.z / f;
            }
            for (int k = 0; k < m; k++) {
                Body<R>& b = p[idx[k]]; Alt<R>& a = alt[idx[k]];
                R f = m0 + b.mass;
                a.pos.x = a.pos.x + dt * (px - (b.mass * a.vel.x / f));
                a.pos.y = a.pos.y + dt * (py - (b.mass * a.vel.y / f));
                a.pos.z = a.pos.z + dt * (pz - (b.mass * a.vel.z / f));
            }
        }
    }
    void interaction_step(double dt_) {
        R dt = dt_;
        int idx[MAXP]; int m = others(h_mm, idx);
        if (coord == PB200_COORD_JACOBI) {
            inertial_to_jacobi_acc();
            double softening = 1e-12;
            R eta = p[h_mm].mass;
            for (int k = 0; k < m; k++) {
                Alt<R>& a = alt[idx[k]];
                eta = eta + a.mass;
                a.vel.x = a.vel.x + dt * a.acc.x; a.vel.y = a.vel.y + dt * a.acc.y; a.vel.z = a.vel.z + dt * a.acc.z;
                if (k > 0) {
                    R rj2i = 1. / (powi(a.pos.x, 2) + powi(a.pos.y, 2) + powi(a.pos.z, 2) + softening);
                    R rji = o_sqrt(rj2i);
                    R rj3im = rji * rj2i * G * eta;
                    R prefac = dt * rj3im;
                    a.vel.x = a.vel.x + prefac * a.pos.x; a.vel.y = a.vel.y + prefac * a.pos.y; a.vel.z = a.vel.z + prefac * a.pos.z;
                }
            }
        } else if (coord == PB200_COORD_DEMOCRATIC_HELIOCENTRIC) {
            for (int k = 0; k < m; k++) {
                Body<R>& b = p[idx[k]]; Alt<R>& a = alt[idx[k]];
                a.vel.x = a.vel.x + dt * b.iacc.x; a.vel.y = a.vel.y + dt * b.iacc.y; a.vel.z = a.vel.z + dt * b.iacc.z;
            }
        } else {
            R m0 = p[h_mm].mass;
            for (int k = 0; k < m; k++) {
                Body<R>& b = p[idx[k]]; Alt<R>& a = alt[idx[k]];
                R f = m0 + b.mass;
                a.vel.x = a.vel.x + dt * f * b.iacc.x / m0; a.vel.y = a.vel.y + dt * f * b.iacc.y / m0; a.vel.z = a.vel.z + dt * f * b.iacc.z / m0;
            }
        }
    }
    // whfast.rs:844-876
    static void stumpff_cs3(R z, R* cs) {
        static const double invfactorial[35] = {1., 1., 1. / 2., 1. / 6., 1. / 24., 1. / 120., 1. / 720., 1. / 5040., 1. / 40320., 1. / 362880., 1. / 3628800., 1. / 39916800., 1. / 479001600., 1. / 6227020800., 1. / 87178291200., 1. / 1307674368000., 1. / 20922789888000., 1. / 355687428096000., 1. / 6402373705728000., 1. / 121645100408832000., 1. / 2432902008176640000., 1. / 51090942171709440000., 1. / 1124000727777607680000., 1. / 25852016738884976640000., 1. / 620448401733239439360000., 1. / 15511210043330985984000000., 1. / 403291461126605635584000000., 1. / 10888869450418352160768000000., 1. / 304888344611713860501504000000., 1. / 8841761993739701954543616000000., 1. / 265252859812191058636308480000000., 1. / 8222838654177922817725562880000000., 1. / 263130836933693530167218012160000000., 1. / 8683317618811886495518194401280000000., 1. / 295232799039604140847618609643520000000.};
        int nn = 0;
        // DEVIATION D3: the reference's `while z.abs() > 0.1` never ends for z = +-inf (a blown-up state); bounded here — any
        // finite double is below 0.1 after 515 divisions by 4, so finite inputs are unaffected.
        while (o_val(o_abs(z)) > 0.1 && nn < 600) { z = z / 4.; nn++; }
        const int nmax = 13;
        R c_odd = invfactorial[nmax];
        R c_even = invfactorial[nmax - 1];
        for (int np = nmax - 2; np >= 3; np -= 2) {
            c_odd = invfactorial[np] - z * c_odd;
            c_even = invfactorial[np - 1] - z * c_even;
        }
        cs[3] = c_odd; cs[2] = c_even;
        cs[1] = invfactorial[1] - z * c_odd;
        cs[0] = invfactorial[0] - z * c_even;
        for (; nn > 0; nn--) {
            cs[3] = (cs[2] + cs[0] * cs[3]) * 0.25;
            cs[2] = cs[1] * cs[1] * 0.5;
            cs[1] = cs[0] * cs[1];
            cs[0] = 2. * cs[0] * cs[0] - 1.;
        }
        cs[4] = 0.; cs[5] = 0.;
    }
    // whfast.rs:835-842
    static void stiefel_gs3(R beta, R x, R* gs) {
        R x2 = powi(x, 2);
        stumpff_cs3(beta * x2, gs);
        gs[1] = gs[1] * x; gs[2] = gs[2] * x2; gs[3] = gs[3] * (x2 * x);
    }
    // whfast.rs:676-833
    uint64_t kepler_stumpff_calls = 0;
    // which branch of the solver each call took (test instrumentation): [0] Newton converged, [1] quartic solver entered,
    // [2] bisection fallback entered, [3] hyperbolic orbit (beta <= 0)
    uint64_t kepler_branches[4] = {0, 0, 0, 0};
    void kepler_individual_step(int i, R mass_g, double dt_) {
        R dt = dt_;
        V3<R> p1 = alt[i].pos, v1 = alt[i].vel;
        R r0 = o_sqrt(powi(p1.x, 2) + powi(p1.y, 2) + powi(p1.z, 2));
        R r0i = 1. / r0;
        R v2 = powi(v1.x, 2) + powi(v1.y, 2) + powi(v1.z, 2);
        R beta = 2. * mass_g * r0i - v2;
        R eta0 = p1.x * v1.x + p1.y * v1.y + p1.z * v1.z;
        R zeta0 = mass_g - beta * r0;
        R x, gs[6];
        R invperiod = 0.;
        R x_per_period;
        bool xpp_nan = false;
        if (o_val(beta) > 0.) {
            R sqrt_beta = o_sqrt(beta);
            invperiod = sqrt_beta * beta / (2. * C_PI * mass_g);
            x_per_period = 2. * C_PI / sqrt_beta;
            if (o_val(o_abs(dt) * invperiod) > 1. && timestep_warning == 0) {
                timestep_warning += 1;
                warnings |= PB200_WARN_TIMESTEP_GT_PERIOD;
            }
            R dtr0i = dt * r0i;
            x = dtr0i * (1. - dtr0i * eta0 * 0.5 * r0i);
        } else {
            x = 0.;
            x_per_period = NAN;
            xpp_nan = true;
            kepler_branches[3]++;
        }
        int converged = 0;
        R old_x = x;
        stiefel_gs3(beta, x, gs); kepler_stumpff_calls++;
        R e1 = eta0 * gs[1] + zeta0 * gs[2];
        R ri = 1. / (r0 + e1);
        x = ri * (x * e1 - eta0 * gs[2] - zeta0 * gs[3] + dt);
        if (!xpp_nan && o_val(o_abs(x - old_x)) > o_val(0.01 * x_per_period)) {
            kepler_branches[1]++;
            x = beta * dt / mass_g;
            R prev_x[WHFAST_NMAX_QUART + 1];
            for (int k = 0; k <= WHFAST_NMAX_QUART; k++) prev_x[k] = 0.;
            for (int n_lag = 1; n_lag < WHFAST_NMAX_QUART; n_lag++) {
                stiefel_gs3(beta, x, gs); kepler_stumpff_calls++;
                R f = r0 * x + eta0 * gs[2] + zeta0 * gs[3] - dt;
                R fp = r0 + eta0 * gs[1] + zeta0 * gs[2];
                R fpp = eta0 * gs[0] + zeta0 * gs[1];
                R denom = fp + o_sqrt(o_abs(16. * fp * fp - 20. * f * fpp));
                x = (x * denom - 5. * f) / denom;
                bool hit = false;
                for (int k = 1; k < n_lag; k++) if (o_val(x) == o_val(prev_x[k])) { hit = true; break; }
                if (hit) { converged = 1; break; }
                prev_x[n_lag] = x;
            }
            R e = eta0 * gs[1] + zeta0 * gs[2];
            ri = 1. / (r0 + e);
        } else {
            R old_x2;
            for (int k = 1; k < WHFAST_NMAX_NEWT; k++) {
                old_x2 = old_x;
                old_x = x;
                stiefel_gs3(beta, x, gs); kepler_stumpff_calls++;
                R e = eta0 * gs[1] + zeta0 * gs[2];
                ri = 1. / (r0 + e);
                x = ri * (x * e - eta0 * gs[2] - zeta0 * gs[3] + dt);
                if (o_val(x) == o_val(old_x) || o_val(x) == o_val(old_x2)) { converged = 1; kepler_branches[0]++; break; }
            }
        }
        if (converged == 0) {
            kepler_branches[2]++;
            R x_min, x_max;
            if (o_val(beta) > 0.) {
                x_min = x_per_period * o_floor(dt * invperiod);
                x_max = x_min + x_per_period;
            } else {
                R h2 = r0 * r0 * v2 - eta0 * eta0;
                R q = h2 / mass_g / (1. + o_sqrt(1. - h2 * beta / (mass_g * mass_g)));
                R vq = o_sqrt(h2) / q;
                x_min = 1. / (vq + r0 / dt);
                x_max = dt / q;
            }
            x = (x_max + x_min) / 2.;
            // DEVIATION D2: the reference's `loop {}` never ends on NaN input; bounded here (2^-200 < 1e-15 long before).
            for (int guard = 0; guard < 200; guard++) {
                stiefel_gs3(beta, x, gs); kepler_stumpff_calls++;
                R s = r0 * x + eta0 * gs[2] + zeta0 * gs[3] - dt;
                if (o_val(s) >= 0.) x_max = x; else x_min = x;
                x = (x_max + x_min) / 2.;
                if (o_val(o_abs(x_max - x_min) / x_max) <= 1e-15) break;
            }
            R e = eta0 * gs[1] + zeta0 * gs[2];
            ri = 1. / (r0 + e);
        }
        if (o_isnan(o_val(ri))) { ri = 0.; gs[1] = 0.; gs[2] = 0.; gs[3] = 0.; }
        R f = -mass_g * gs[2] * r0i;
        R g = dt - mass_g * gs[3];
        R fd = -mass_g * gs[1] * r0i * ri;
        R gd = -mass_g * gs[2] * ri;
        Alt<R>& a = alt[i];
        a.pos.x = a.pos.x + (f * p1.x + g * v1.x);
        a.pos.y = a.pos.y + (f * p1.y + g * v1.y);
        a.pos.z = a.pos.z + (f * p1.z + g * v1.z);
        a.vel.x = a.vel.x + (fd * p1.x + gd * v1.x);
        a.vel.y = a.vel.y + (fd * p1.y + gd * v1.y);
        a.vel.z = a.vel.z + (fd * p1.z + gd * v1.z);
    }
    // whfast.rs:628-672
    void kepler_steps(double dt_) {
        R star_mg = p[h_mm].mass_g;
        if (coord == PB200_COORD_JACOBI) {
            R mg = star_mg;
            for (int i = 0; i < n; i++) { if (i == h_mm) continue; mg = mg + alt[i].mass_g; kepler_individual_step(i, mg, dt_); }
        } else if (coord == PB200_COORD_DEMOCRATIC_HELIOCENTRIC) {
            for (int i = 0; i < n; i++) { if (i == h_mm) continue; kepler_individual_step(i, star_mg, dt_); }
        } else {
            for (int i = 0; i < n; i++) { if (i == h_mm) continue; kepler_individual_step(i, star_mg + alt[i].mass_g, dt_); }
        }
        R dt = dt_;
        Alt<R>& s = alt[h_mm];
        s.pos.x = s.pos.x + dt * s.vel.x; s.pos.y = s.pos.y + dt * s.vel.y; s.pos.z = s.pos.z + dt * s.vel.z;
    }

    // ------------------------------------------------------------------ historic record (output.rs:119-163)
    void write_historic_record() {
        size_t off = history.size();
        history.resize(off + (size_t)n * PB200_HISTORIC_RECORD_BYTES);
        uint8_t* w = history.data() + off;
        auto put_f64 = [&](double v) { std::memcpy(w, &v, 8); w += 8; };
        for (int i = 0; i < n; i++) {
            Body<R>& b = p[i];
            put_f64(current_time); put_f64(time_step);
            int32_t id = b.id; std::memcpy(w, &id, 4); w += 4;
            put_f64(o_val(b.ipos.x)); put_f64(o_val(b.ipos.y)); put_f64(o_val(b.ipos.z));
            put_f64(o_val(b.spin.x)); put_f64(o_val(b.spin.y)); put_f64(o_val(b.spin.z));
            put_f64(o_val(b.ivel.x)); put_f64(o_val(b.ivel.y)); put_f64(o_val(b.ivel.z));
            put_f64(o_val(b.mass)); put_f64(o_val(b.radius)); put_f64(o_val(b.rg2));
            put_f64(b.t_role != PB200_ROLE_DISABLED ? o_val(b.t_k2) : 0.);
            put_f64(o_val(b.t_sigma)); put_f64(o_val(b.t_lag)); put_f64(o_val(b.t_denergy));
            put_f64(0.);  // disk migration_timescale (disk is out of scope, always 0)
        }
    }

    // ------------------------------------------------------------------ whfast.rs:235-305; returns true while the system keeps running
    bool iterate() {
        if (status != PB200_STATUS_OK) return false;
        bool first = last_historic_snapshot_time < 0.;
        bool due = last_historic_snapshot_time + historic_snapshot_period <= current_time;
        if (first || due) {
            inertial_to_heliocentric();
            calculate_spin_and_evolving_quantities(current_time, true);
            if (c_tides) tides_denergy_dt(h_tides);
            write_historic_record();
            if (!first) last_historic_snapshot_time += historic_snapshot_period; else last_historic_snapshot_time = 0.;
            n_historic_snapshots += 1;
        }
        Ignore ign = ignore_terms();
        bool gr_spin = c_gr && gr_impl == PB200_GR_KIDDER1995;
        bool integrate_spin = c_tides || c_flat || c_evo || gr_spin;
        integrate_velocity_dependent_forces(half_time_step, integrate_spin, true);
        inertial_to_alternative_posvel();
        kepler_steps(half_time_step);
        jump_step(half_time_step);
        alternative_to_inertial_posvel();
        gravity_calculate_acceleration(ign);
        if (status != PB200_STATUS_OK) return false;  // the reference panics inside gravity
        interaction_step(time_step);
        jump_step(half_time_step);
        kepler_steps(half_time_step);
        alternative_to_inertial_posvel();
        integrate_velocity_dependent_forces(half_time_step, integrate_spin, false);
        current_time += time_step;
        current_iteration += 1;
        if (current_time + time_step > time_limit) { status = PB200_STATUS_COMPLETED; event_iteration = current_iteration; return false; }
        return true;
    }

    // universe.rs:625-658 (after a heliocentric refresh)
    void summary(double& energy, double& angmom) {
        inertial_to_heliocentric();
        R e_kin = 0., e_pot = 0.;
        for (int i = 0; i < n; i++)
            e_kin = e_kin + 0.5 * p[i].mass * (powi(p[i].hvel.x, 2) + powi(p[i].hvel.y, 2) + powi(p[i].hvel.z, 2));
        for (int i = 0; i < n; i++)
            for (int j = i + 1; j < n; j++) {
                R dx = p[i].hpos.x - p[j].hpos.x, dy = p[i].hpos.y - p[j].hpos.y, dz = p[i].hpos.z - p[j].hpos.z;
                e_pot = e_pot - p[j].mass_g * p[i].mass / o_sqrt(powi(dx, 2) + powi(dy, 2) + powi(dz, 2));
            }
        energy = o_val(e_kin + e_pot + 0.);
        V3<R> Lt{0., 0., 0.};
        for (int i = 0; i < n; i++) {
            Lt.x = Lt.x + p[i].mass * (p[i].hpos.y * p[i].hvel.z - p[i].hpos.z * p[i].hvel.y);
            Lt.y = Lt.y + p[i].mass * (p[i].hpos.z * p[i].hvel.x - p[i].hpos.x * p[i].hvel.z);
            Lt.z = Lt.z + p[i].mass * (p[i].hpos.x * p[i].hvel.y - p[i].hpos.y * p[i].hvel.x);
        }
        angmom = o_val(o_sqrt(powi(Lt.x, 2) + powi(Lt.y, 2) + powi(Lt.z, 2)));
    }

    // ------------------------------------------------------------------ image <-> flat case
    void load(const pb200_case_t& c, const pb200_table_t* tables, size_t n_tables) {
        time_step = c.time_step; half_time_step = c.half_time_step;
        initial_time = c.initial_time; time_limit = c.time_limit; current_time = c.current_time;
        recovery_snapshot_period = c.recovery_snapshot_period; historic_snapshot_period = c.historic_snapshot_period;
        last_recovery_snapshot_time = c.last_recovery_snapshot_time; last_historic_snapshot_time = c.last_historic_snapshot_time;
        current_iteration = c.current_iteration; n_historic_snapshots = c.n_historic_snapshots; timestep_warning = c.timestep_warning;
        coord = c.coordinates_type; n = c.n_particles;
        c_tides = c.consider_tides; c_flat = c.consider_rotational_flattening; c_gr = c.consider_general_relativity;
        c_disk = c.consider_disk; c_wind = c.consider_wind; c_evo = c.consider_evolution;
        gr_impl = c.general_relativity_implementation;
        h_mm = c.host_most_massive; h_tides = c.host_tides; h_flat = c.host_rotational_flattening; h_gr = c.host_general_relativity;
        // HostMostMassive flags as computed by find_indices (universe.rs:981-988)
        mm_tides = c_tides && h_mm == h_tides;
        mm_flat = c_flat && h_mm == h_flat;
        mm_gr = c_gr && h_mm == h_gr;
        mm_all = (mm_tides || !c_tides) && (mm_flat || !c_flat) && (mm_gr || !c_gr);
        // universe.rs:959-963: flattening without tides borrows the tides host slot
        evolvers.assign(MAXP, Table());
        gr_impl_of_host = PB200_GR_DISABLED;
        for (int i = 0; i < n; i++) {
            const pb200_body_t& s = c.bodies[i];
            Body<R>& b = p[i];
            b = Body<R>();
            b.id = s.id;
            b.mass = s.mass; b.mass_g = s.mass_g; b.radius = s.radius; b.rg2 = s.radius_of_gyration_2; b.moi = s.moment_of_inertia;
            b.ipos = {s.inertial_position[0], s.inertial_position[1], s.inertial_position[2]};
            b.ivel = {s.inertial_velocity[0], s.inertial_velocity[1], s.inertial_velocity[2]};
            b.iacc = {s.inertial_acceleration[0], s.inertial_acceleration[1], s.inertial_acceleration[2]};
            b.hpos = {s.heliocentric_position[0], s.heliocentric_position[1], s.heliocentric_position[2]};
            b.hvel = {s.heliocentric_velocity[0], s.heliocentric_velocity[1], s.heliocentric_velocity[2]};
            b.spin = {s.spin[0], s.spin[1], s.spin[2]};
            b.norm_spin2 = (powi(b.spin.x, 2)) + (powi(b.spin.y, 2)) + (powi(b.spin.z, 2));
            b.L = {s.angular_momentum[0], s.angular_momentum[1], s.angular_momentum[2]};
            b.t_role = s.tides_role; b.t_dissipation_factor = s.tides_dissipation_factor;
            b.t_dissipation_factor_scale = s.tides_dissipation_factor_scale; b.t_k2 = s.tides_love_number;
            b.t_sigma = s.tides_scaled_dissipation_factor; b.t_lag = s.tides_lag_angle; b.t_denergy = s.tides_denergy_dt;
            b.f_role = s.flattening_role; b.f_k2 = s.flattening_love_number;
            b.g_role = s.general_relativity_role; b.g_factor = s.general_relativity_factor;
            b.w_role = s.wind_role; b.w_k = s.wind_k_factor; b.w_sat = s.wind_rotation_saturation;
            b.w_sat2 = powi(b.w_sat, 2);   // wind.rs:52, recomputed from the input (the image stores both)
            b.evo_type = s.evolution_type; b.evo_param = s.evolution_parameter; b.evo_table = s.evolution_table;
            if (s.evolution_table >= 0 && (size_t)s.evolution_table < n_tables) {
                const pb200_table_t& t = tables[s.evolution_table];
                Table& e = evolvers[i];
                e.time.assign(t.time, t.time + t.n_rows);
                if (t.radius) e.radius.assign(t.radius, t.radius + t.n_rows);
                if (t.radius_of_gyration_2) e.rg2.assign(t.radius_of_gyration_2, t.radius_of_gyration_2 + t.n_rows);
                if (t.love_number) e.love.assign(t.love_number, t.love_number + t.n_rows);
                if (t.inverse_tidal_q_factor) e.qinv.assign(t.inverse_tidal_q_factor, t.inverse_tidal_q_factor + t.n_rows);
                e.left_index = (size_t)s.evolution_left_index;
            }
            verr[i] = {c.inertial_velocity_errors[i][0], c.inertial_velocity_errors[i][1], c.inertial_velocity_errors[i][2]};
            lerr[i] = {c.particle_angular_momentum_errors[i][0], c.particle_angular_momentum_errors[i][1], c.particle_angular_momentum_errors[i][2]};
        }
        if (c_gr && h_gr >= 0 && h_gr < n && p[h_gr].g_role == PB200_ROLE_CENTRAL) gr_impl_of_host = gr_impl;
        std::memcpy(roche, c.roche_radiuses, sizeof(roche));
        pair_clear();
        for (int k = 0; k < MAXP * MAXP; k++) {
            double v = c.pair_dependent_scaled_dissipation_factor[k];
            if (v == v) { pair_sigma[k] = v; pair_sigma_set[k] = true; }   // NaN = key absent
        }
        status = PB200_STATUS_OK; warnings = 0; event_iteration = 0;
    }
    void store(pb200_case_t& c) const {
        c.time_step = time_step; c.half_time_step = half_time_step; c.initial_time = initial_time; c.time_limit = time_limit;
        c.current_time = current_time; c.recovery_snapshot_period = recovery_snapshot_period; c.historic_snapshot_period = historic_snapshot_period;
        c.last_recovery_snapshot_time = last_recovery_snapshot_time; c.last_historic_snapshot_time = last_historic_snapshot_time;
        c.current_iteration = current_iteration; c.n_historic_snapshots = n_historic_snapshots; c.timestep_warning = timestep_warning;
        for (int i = 0; i < n; i++) {
            pb200_body_t& s = c.bodies[i];
            const Body<R>& b = p[i];
            s.radius = o_val(b.radius); s.radius_of_gyration_2 = o_val(b.rg2); s.moment_of_inertia = o_val(b.moi);
            s.inertial_position[0] = o_val(b.ipos.x); s.inertial_position[1] = o_val(b.ipos.y); s.inertial_position[2] = o_val(b.ipos.z);
            s.inertial_velocity[0] = o_val(b.ivel.x); s.inertial_velocity[1] = o_val(b.ivel.y); s.inertial_velocity[2] = o_val(b.ivel.z);
            s.inertial_acceleration[0] = o_val(b.iacc.x); s.inertial_acceleration[1] = o_val(b.iacc.y); s.inertial_acceleration[2] = o_val(b.iacc.z);
            s.heliocentric_position[0] = o_val(b.hpos.x); s.heliocentric_position[1] = o_val(b.hpos.y); s.heliocentric_position[2] = o_val(b.hpos.z);
            s.heliocentric_velocity[0] = o_val(b.hvel.x); s.heliocentric_velocity[1] = o_val(b.hvel.y); s.heliocentric_velocity[2] = o_val(b.hvel.z);
            s.spin[0] = o_val(b.spin.x); s.spin[1] = o_val(b.spin.y); s.spin[2] = o_val(b.spin.z);
            s.angular_momentum[0] = o_val(b.L.x); s.angular_momentum[1] = o_val(b.L.y); s.angular_momentum[2] = o_val(b.L.z);
            s.tides_lag_angle = o_val(b.t_lag); s.tides_denergy_dt = o_val(b.t_denergy);
            s.general_relativity_factor = o_val(b.g_factor);
            s.evolution_left_index = (int32_t)evolvers[i].left_index;
            c.inertial_velocity_errors[i][0] = o_val(verr[i].x); c.inertial_velocity_errors[i][1] = o_val(verr[i].y); c.inertial_velocity_errors[i][2] = o_val(verr[i].z);
            c.particle_angular_momentum_errors[i][0] = o_val(lerr[i].x); c.particle_angular_momentum_errors[i][1] = o_val(lerr[i].y); c.particle_angular_momentum_errors[i][2] = o_val(lerr[i].z);
        }
        std::memcpy(c.roche_radiuses, roche, sizeof(roche));
        for (int k = 0; k < MAXP * MAXP; k++) c.pair_dependent_scaled_dissipation_factor[k] = pair_sigma_set[k] ? pair_sigma[k] : std::nan("");
    }
};

}  // namespace pb200_oracle
