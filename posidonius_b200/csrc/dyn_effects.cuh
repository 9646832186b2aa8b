// dyn_effects.cuh — stellar wind and dynamical tides, the two effects that only the run-time geometry build carries.
//
//   wind::calculate_wind_factor                              effects/wind.rs:72-91
//   calculate_particles_spin_dependent_evolving_quantities   effects/evolution.rs:548-567   (lag angle)
//   calculate_pair_dependent_scaled_dissipation_factors      effects/tides/constant_time_lag.rs:20-165
//   tools::calculate_perihelion_distance_and_eccentricity    tools.rs:251-285
//
// Strict arithmetic in the reference's association order (transcribed from the CPU oracle, which reproduces the
// reference goldens of the solar-like fixtures bit for bit). The one operation that is not IEEE-exact is
// powf(-1.5): the reference calls libm's pow, here it is CUDA's pow (<= 1 ulp apart); the pair-dependent sigma it feeds
// scales a stellar tide of relative size < 1e-9, far below the golden tolerance.
// Included inside namespace PB_NS; no include guard on purpose (one copy per geometry build).

namespace PB_NS {
using namespace pb200;

// effects/evolution.rs:548-567 for this lane's body; w2 = |spin|^2 of the fresh calculate_spin
__device__ __forceinline__ void update_lag_angle(const KParams& P, const Roles& ro, int b, size_t sys, double t, sd w2, bool commit) {
    if (!ro.valid || !commit) return;
    double lag = 0.;
    if ((P.dyn_evo >> b) & 1u) {
        const int ti = P.evo_table[b];
        if (ti >= 0 && P.tables[ti].qinv) {
            const DevTable& T = P.tables[ti];
            const int k = table_upper(T.time, T.n_rows, t);
            const sd qinv = sd(table_interp(T.time, T.qinv, T.n_rows, k, t));
            const sd eps2 = w2 / sd(kSunDynFreq2);
            lag = (sd(3.0) * eps2 * qinv / sd(4.0)).v;
        }
    }
    P.lag[(size_t)b * (size_t)P.n_sys + sys] = lag;
}

// tools.rs:251-285
__device__ __forceinline__ void perihelion_and_eccentricity(sd gm, S3 r, S3 v, sd& q, sd& e) {
    const sd hx = r.y * v.z - r.z * v.y, hy = r.z * v.x - r.x * v.z, hz = r.x * v.y - r.y * v.x;
    const sd h2 = hx * hx + hy * hy + hz * hz;   // powf(2.) is folded to a product by LLVM
    const sd v2 = v.x * v.x + v.y * v.y + v.z * v.z;
    const sd rr = ssqrt_ieee(r.x * r.x + r.y * r.y + r.z * r.z);
    const sd s = h2 / gm;
    const sd temp = sd(1.) + s * (v2 / gm - sd(2.) / rr);
    e = temp.v <= 0. ? sd(0.) : ssqrt_ieee(temp);
    q = s / (sd(1.) + e);
}

// constant_time_lag.rs:20-165 for the pair (host, this lane's body): the scaled dissipation factors that
// calculate_orthogonal/radial_component_of_the_tidal_force will read (get_pair_dependent_scaled_dissipation_factor_or_else).
// hr, hv: tidal (heliocentric) coordinates; w2, wh2: |spin|^2 of this body and of the host.
__device__ __forceinline__ void pair_dependent_sigmas(const KParams& P, const Roles& ro, const Cold& cold, int hl, int b, size_t sys,
                                                      S3 hr, S3 hv, sd w2, sd wh2, sd& sig_h, sd& sig_p) {
    const size_t ns = (size_t)P.n_sys;
    const size_t i = (size_t)b * ns + sys;
    const bool host_dyn = (P.dyn_evo >> PB_HOST(P)) & 1u;
    const bool body_dyn = ro.t_on && ((P.dyn_evo >> b) & 1u);
    double base = 0., lag = 0., diss = 0., scale = 0., stale = __longlong_as_double(0x7ff8000000000000LL);
    if (ro.valid) { base = P.sigma[i]; lag = ldm(P.lag + i); diss = P.diss[i]; scale = P.diss_scale[i]; stale = ldm(P.pair_p + i); }
    const sd base_h = sd(shfl(base, hl)), lag_h = sd(shfl(lag, hl));
    // (0, 0) unless the host is TidesEffect::CentralBody (constant_time_lag.rs:33-41)
    const sd diss_h = sd(P.tides_host_central ? shfl(diss, hl) : 0.), scale_h = sd(P.tides_host_central ? shfl(scale, hl) : 0.);
    const sd R = sd(cold.get(K_R)), Rh = sd(shfl(R.v, hl));
    // mu_host + mu (cold path: the gravitational masses stay in global memory)
    const sd gm = sd(P.mass_g[(size_t)PB_HOST(P) * ns + (ro.valid ? sys : 0)]) + sd(ro.valid ? P.mass_g[i] : 1.);
    sd q, e;
    perihelion_and_eccentricity(gm, hr, hv, q, e);
    const sd mean_motion = ssqrt_ieee(gm) * sd(pow((q / (sd(1.0) - e)).v, -1.5));
    bool set_h = false;
    sd val_h = sd(0.);
    if (host_dyn) {
        const sd wn = ssqrt_ieee(wh2);
        sd half = sabs(wn - mean_motion);
        if (half.v < wn.v) {
            if (half.v < kSmoothDynTide) half = sd(kSmoothDynTide);
            const sd inv = sd(1.) / half;
            const sd Rh2 = Rh * Rh, Rh4 = Rh2 * Rh2;
            val_h = scale_h * (sd(2.0) * sd(kK2) / (sd(3.0) * (Rh * Rh4)) * lag_h * inv + diss_h);
            set_h = true;
        }
    }
    if (body_dyn) {
        const sd wn = ssqrt_ieee(w2);
        sd half = sabs(wn - mean_motion);
        if (half.v < wn.v) {
            if (half.v < kSmoothDynTide) half = sd(kSmoothDynTide);
            const sd inv = sd(1.) / half;
            const sd R2 = R * R, R4 = R2 * R2;
            stale = (sd(scale) * (sd(2.0) * sd(kK2) / (sd(3.0) * (R * R4)) * sd(lag) * inv + sd(diss))).v;
        } else {
            set_h = false;   // Q7: the equilibrium branch removes the (host, particle) key, not (particle, host)
        }
    }
    sig_h = (host_dyn && set_h) ? val_h : base_h;
    sig_p = (body_dyn && stale == stale) ? sd(stale) : sd(base);
    if (ro.t_on) {
        // map image for recovery snapshots; an entry of a non-dynamical host is never written by the reference
        if (host_dyn) P.pair_h[i] = set_h ? val_h.v : __longlong_as_double(0x7ff8000000000000LL);
        else if (body_dyn && !set_h) P.pair_h[i] = __longlong_as_double(0x7ff8000000000000LL);
        if (body_dyn) P.pair_p[i] = stale;
    }
}

// wind.rs:72-91: dL/dt of this lane's body (zero unless WindEffect::Interaction); s, w2 = fresh spin and |spin|^2
__device__ __forceinline__ S3 wind_dangular_momentum_dt(const KParams& P, const Roles& ro, const Cold& cold, int b, size_t sys, S3 s, sd w2) {
    const bool on = ro.valid && ((P.wind_on >> b) & 1u);
    double k = 0., sat = 1.;
    if (on) { const size_t i = (size_t)b * (size_t)P.n_sys + sys; k = P.wind_k[i]; sat = P.wind_sat[i]; }
    const sd threshold = ssqrt_ieee(w2);
    const sd factor = threshold.v >= sat ? sd(sat) * sd(sat) : w2;   // rotation_saturation_2 = powi(2) (wind.rs:52)
    const sd root = ssqrt_ieee(sd(cold.get(K_R)) / sd(kRSun) * sd(1.) / sd(cold.get(K_M)));
    const sd mk = sd(-1.) * sd(k);
    S3 out = s3(mk * s.x * factor * root, mk * s.y * factor * root, mk * s.z * factor * root);
    if (!on) out = s3(sd(0.), sd(0.), sd(0.));
    return out;
}

}  // namespace PB_NS
