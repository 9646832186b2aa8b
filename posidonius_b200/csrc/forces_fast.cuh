// forces_fast.cuh — PB200_ARITH_FAST perturbation forces (tides, flattening, GR Kidder1995) and their per-system constants.
// Included once per geometry specialisation (see pb200_api.cu), inside namespace PB_NS; no include guard on purpose.
#include "whfast_kernel.cuh"
#include "cold_slots.cuh"
#include "dyn_effects.cuh"

namespace PB_NS {
using namespace pb200;

extern __shared__ __align__(128) double pb_smem[];   // 64-byte aligned groups of 8 columns: dist_ld / dist_st flip address bits 3-5

// ---- Distributed ordered sums (8 bodies, host 0). The reference accumulates its sums over the bodies serially, so the
// association order is fixed and every lane of the group used to walk all seven terms of all three components of every
// sum itself (bit-identical copies, 21 LDS + 21 DADD per vector sum and lane). The SCALAR sums are independent of each
// other, though: lane c of the group accumulates scalar c over the bodies in index order — same additions, same order,
// same bits — applies the scalar's division, and leaves the result in one slot that the group reads back. Body k's term
// of scalar c sits in column (k ^ c) of slot base + c, so that at every step of the walk the reducing lanes touch
// different banks (a plain [scalar][body] layout would be a six-way bank conflict).
#define PB_DIST (PB_FIXED_N == 8)
// A "row" is the shared-memory BYTE address of this thread's column in a slot; body (b ^ k)'s column is row ^ (k << 3): one
// LOP3 per access (the dynamic shared memory is 128-byte aligned, a group's eight columns are one aligned 64-byte line).
// The addresses are loop-invariant; left alone, the compiler hoists all of them out of the step loop into ~30 registers
// (and spills). An empty volatile asm makes the thread's own address opaque where it is used.
__device__ __forceinline__ unsigned dist_self(const Cold& cold) {
    unsigned a = (unsigned)__cvta_generic_to_shared((const void*)cold.base);
    asm volatile("" : "+r"(a));
    return a;
}
__device__ __forceinline__ double dist_ld(unsigned row, int k) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(row ^ (unsigned)(k << 3)) : "memory");
    return v;
}
__device__ __forceinline__ void dist_st(unsigned row, int k, double v) {
    asm volatile("st.shared.f64 [%0], %1;" : : "r"(row ^ (unsigned)(k << 3)), "d"(v) : "memory");
}
__device__ __forceinline__ void dist_put(const Cold& cold, int base, int c, double v) { dist_st(dist_self(cold) + (unsigned)((base + c) * PB_BLOCK * 8), c, v); }
__device__ __forceinline__ void dist_put3(const Cold& cold, int base, int c0, V3 v) { dist_put(cold, base, c0, v.x); dist_put(cold, base, c0 + 1, v.y); dist_put(cold, base, c0 + 2, v.z); }
// the row of the scalar this lane reduces (lanes beyond the last scalar repeat it; their results are never read)
__device__ __forceinline__ unsigned dist_row(const Cold& cold, int base, int b, int n_scalars) { return dist_self(cold) + (unsigned)((base + (b < n_scalars ? b : n_scalars - 1)) * PB_BLOCK * 8); }
// acc (+/-)= term of body 1, 2, ... 7 in index order
template <bool SUB>
__device__ __forceinline__ sd dist_walk(unsigned row, sd acc) {
#pragma unroll
    for (int k = 1; k < 8; k++) { const sd t = sd(dist_ld(row, k)); acc = SUB ? acc - t : acc + t; }
    return acc;
}

// Derives the force constants from masses, radii and dissipation parameters (cold path: launch start and whenever a
// radius evolves). One routine serves every arithmetic mode. What the exact forces (exact_effects.cuh) read is computed
// with `sd` in the reference's association order: the base quantities (sigma, k2, k2f, R^5, R^10 with powi as LLVM expands
// it — strict_effects' Z_* constants are rebuilt from them per evaluation), the tidal numerators, 1 / m, the GR factor
// polynomials, mass_factor (M + m) and the reduced mass. The remaining constants are the fast forces' folded products.
__device__ __forceinline__ void make_consts(const KParams& P, const Roles& ro, const Cold& cold, int hl, int b, size_t sys) {
    double sigma = 0., k2t = 0., k2f = 0., mg = 1.;
    if (ro.valid) {
        const size_t i = (size_t)b * (size_t)P.n_sys + sys;
        sigma = P.sigma[i]; k2t = P.k2t[i]; k2f = P.k2f[i]; mg = P.mass_g[i];
    }
    const double m = cold.get(K_M), R = cold.get(K_R), I = cold.get(K_I);
    const sd R_s = sd(R), R2_s = R_s * R_s, R4_s = R2_s * R2_s, R8_s = R4_s * R4_s;
    const double R5 = (R_s * R4_s).v, R10 = (R2_s * R8_s).v;      // powi(5) = x * x^4, powi(10) = x^2 * x^8 (Q10)
    cold.set(K_SIG, sigma); cold.set(K_K2T, k2t); cold.set(K_K2F, k2f); cold.set(K_R5, R5); cold.set(K_R10, R10);
    const double M = shfl(m, hl), Mg = shfl(mg, hl), Ih = shfl(I, hl);
    const double Rh5 = shfl(R5, hl), Rh10 = shfl(R10, hl);
    const double sig_h = shfl(sigma, hl), k2t_h = shfl(k2t, hl), k2f_h = shfl(k2f, hl);
    const sd m_s = sd(m), M_s = sd(M);
    const sd m2_s = m_s * m_s, M2_s = M_s * M_s;
    const double m2 = m2_s.v, M2 = M2_s.v;
    cold.set(C_INVI, 1. / I);
    // Role gates are folded into the constants: a lane that is not an OrbitingBody of an effect (host slot, padding,
    // Disabled role) carries zeros, so the force code needs no per-lane branches or selects.
    const double gt = ro.t_on ? 1. : 0., gf = ro.f_on ? 1. : 0., gg = ro.g_on ? 1. : 0.;
    const double gts = P.tides_host_central ? gt : 0., gfs = P.flat_host_central ? gf : 0.;
    // the tidal numerators in the reference's order (the gates are exact factors 0 / 1)
    cold.set2(C_AS, 0, gts * (sd(4.5) * m2_s * sd(Rh10) * sd(sig_h)).v,      // 4.5 m^2 R*^10 sigma*   (constant_time_lag.rs:232-234)
                       gt * (sd(4.5) * M2_s * sd(R10) * sd(sigma)).v);        // 4.5 M^2 R^10 sigma     (constant_time_lag.rs:243-245)
    const double c_bk = gt * (3.0 * kK2 * (m2 * Rh5 * k2t_h + M2 * R5 * k2t));   // 3 K2 (m^2 R*^5 k2* + M^2 R^5 k2) (:283-285)
    cold.set2(C_AS, 1, gfs * (m * k2f_h * Rh5),                     // flattening: m k2f* R*^5 (oblate_spheroid.rs:37)
                       gf * (M * k2f * R5));                        //             M k2f R^5   (oblate_spheroid.rs:42)
    cold.set(C_INVM, 1. / m);
    const sd mgs_s = sd(Mg) + sd(mg);
    const double mgs = mgs_s.v;
    cold.set2(C_AS, 5, c_bk, gg * mgs);                            // gated: A = mgs / (r^2 c^2) vanishes for non-GR lanes
    // 1.5PN spin-orbit terms (general_relativity.rs:300-456) in terms of the spins (L = I w), G / c^2 and the role gate folded in:
    //   mass_factor * (Lp / m - Ls / M) = Z - S with S = Ls + Lp and Z = (M / m) Lp + (m / M) Ls, so the three vectors of the
    //   acceleration are 2S + msf = S + Z, 3S + msf = 2S + Z and 7S + 3 msf = 4S + 3Z
    const double fa = gg * (kG * kInvC2);
    const sd msum_s = M_s + m_s, mdiff_s = M_s - m_s;
    const sd mured_s = (M_s * m_s) / msum_s;                      // general_relativity.rs:383
    const double mured = mured_s.v;
    cold.set(Z_MURED, mured);
    cold.set(Z_MFM, (mdiff_s / msum_s * msum_s).v);               // mass_factor * star_planet_mass (:321, 336)
    cold.set2(C_AS, 2, I * (M / m), Ih * (m / M));                // Z = zp w_p + zh w_s
    cold.set2(C_AS, 3, fa * I * ((2. + 1.5 * M / m) * mured),     // :419  dLp/dt: (2 + 3 M / 2m) Lorb x Lp
                       fa * Ih * ((2. + 1.5 * m / M) * mured));   // :390  dLs/dt: (2 + 3 m / 2M) Lorb x Ls
    cold.set2(C_AS, 4, fa * m,                                    // force = m * acceleration: the host gets -F / M, the planet F / m
                       fa * I * Ih);                              // Lp x Ls and the 3 (n.L)(n x L) terms
    // polynomials in the GR factor f of the 1PN / 2PN terms (general_relativity.rs:98, 197-205, 256-268): per-system
    // constants, every one a leading sub-expression of the reference's products, so the exact forces read them too
    const sd f = sd(Mg) * sd(mg) / (mgs_s * mgs_s), f2 = f * f;
    // 13 polynomial coefficients: six pair cells and a single
    cold.set2(G_0, 0, (sd(1.0) + sd(3.0) * f).v, (sd(2.0) * (sd(2.0) + f)).v);
    cold.set2(G_0, 1, (sd(1.5) * f).v, (sd(2.0) * (sd(2.0) - f)).v);
    cold.set2(G_0, 2, (sd(0.75) * (sd(12.0) + sd(29.0) * f)).v, (f * (sd(3.0) - sd(4.0) * f)).v);
    cold.set2(G_0, 3, (sd(1.875) * f * (sd(1.0) - sd(3.0) * f)).v, (sd(1.5) * f * (sd(3.0) - sd(4.0) * f)).v);
    cold.set2(G_0, 4, (sd(0.5) * f * (sd(13.0) - sd(4.0) * f)).v, (sd(2.0) + sd(25.0) * f + sd(2.0) * f2).v);
    cold.set2(G_0, 5, (f * (sd(15.0) + sd(4.0) * f)).v, (sd(4.0) + sd(41.0) * f + sd(8.0) * f2).v);
    cold.set(G_0 + 12, (sd(3.0) * f * (sd(3.0) + sd(2.0) * f)).v);
    __syncwarp();   // the host's column (1/M, inertia, base quantities) is read by the other lanes
}

// ---------------------------------------------------------------------------------------------
// Universe::calculate_additional_effects for the lane's body at (hr, hv) with the current L
// (universe.rs:428-614). Returns the inertial additional acceleration and dL/dt of THIS body;
// host-lane values are the group reductions.
// HYB (hybrid arithmetic, see midpoint()): this evaluation is one of the two uncommitted iterates that precede an exact
// evaluation. The spin is then the correctly rounded L / I and the carried r . w products are rounded like the
// reference's, because the exact evaluation that follows reads both (Q3: spins of the previous evaluation).
template <int GR, bool HYB>
__device__ __forceinline__ void additional_effects(const KParams& P, const Roles& ro, const Cold& cold, int hl, int b, size_t sys,
                                                   double t, bool evolve_now, Lane& q, V3 hr, double inv_d, V3 hv, V3& a_out,
                                                   V3& dl_out, bool tide_save) {
    const int W = PB_W(P);
    // Q3: r.omega uses the spins of the previous evaluation (universe.rs:429-430)
    // The two dot products are carried from the previous evaluation (by-products of its spin-orbit terms: no reload of the
    // host's previous spin); the midpoint sets them when the position has changed.
    const double rs_s = q.rs_s, rs_p = q.rs_p;
    // calculate_spin (particles/common.rs:3-15)
    if (HYB) q.s = plain(strict(q.L) / make_rcp(sd(cold.get(K_I))));
    else q.s = cold.get(C_INVI) * q.L;
    const double w2 = HYB ? sdot(strict(q.s), strict(q.s)).v : dot(q.s, q.s);
    // (no barrier before these stores: the group's last reads of E_S / M_6 lie before the previous evaluation's exchange
    // barriers)
    cold.set3(E_S, q.s); cold.set(M_6, w2);
    __syncwarp();
    V3 sh = cold.getk3(PB_HOST(P), E_S);
    double wh2 = cold.getk(PB_HOST(P), M_6);
    // for the next evaluation (Q3) and the 3 (n.L)(n x L) terms below
    if (HYB) { q.rs_s = sdot(strict(hr), strict(sh)).v; q.rs_p = sdot(strict(hr), strict(q.s)).v; }
    else { q.rs_s = dot(hr, sh); q.rs_p = dot(hr, q.s); }
#if !PB_FIXED_N
    // lag angle of the dynamical-tide models, once per step like the other evolving quantities (evolution.rs:548-567)
    if (evolve_now && (PB_FLAGS(P) & FLAG_DYN) && (PB_FLAGS(P) & FLAG_EVO)) { update_lag_angle(P, ro, b, sys, t, sd(w2), true); __syncwarp(); }
#endif
    // Everything below is assembled as ONE force on the planet, F = Kr r + Kv v + (vector terms): the planet's acceleration is
    // F / m, the host's -F / M (Q5: no indirect term), and the torques as coefficient sums over a small vector basis
    // (r, r x v, r x w, (r x v) x w, w_p x w_s) - the effects share the cross products instead of rebuilding them.
    const double inv_d2 = inv_d * inv_d, inv_d4 = inv_d2 * inv_d2;
    const double radvel = dot(hr, hv) * inv_d;
    const V3 rxv = cross(hr, hv);
    const V3 cs = cross(hr, sh), cp = cross(hr, q.s);     // r x w_host, r x w_planet (fresh spins)
    const double inv_m = cold.get(C_INVM), inv_M = cold.getk(PB_HOST(P), C_INVM);
    const double2 bk_mgs = cold.get2(C_AS, 5);   // (3 K2 (...), gated mu_s + mu_p): tides / GR
    double Kr = 0., Kv = 0.;          // coefficients of r and v in F
    double Pcp = 0., Hcs = 0.;        // coefficient of r x w_planet in dLp/dt, of r x w_host in the host's dL/dt
    V3 F = v3(0., 0., 0.), dl_p = v3(0., 0., 0.), dl_h = v3(0., 0., 0.);
    if (PB_FLAGS(P) & FLAG_TIDES) {
        // constant_time_lag.rs:206-332, tides/common.rs:223-345
        const double inv_d6 = inv_d4 * inv_d2, inv_d7 = inv_d6 * inv_d;
        const double2 as_ap = cold.get2(C_AS, 0);
        double FodS = as_ap.x * inv_d6;      // F_orth * r
        double FodP = as_ap.y * inv_d6;
#if !PB_FIXED_N
        if (PB_FLAGS(P) & FLAG_DYN) {
            sd sig_h, sig_p;
            pair_dependent_sigmas(P, ro, cold, hl, b, sys, strict(hr), strict(hv), sd(w2), sd(wh2), sig_h, sig_p);
            // the same numerators without the constant sigma (cold path: run-time geometry build only)
            const double m_ = cold.get(K_M), M_ = cold.getk(PB_HOST(P), K_M);
            const double gt_ = ro.t_on ? 1. : 0., gts_ = P.tides_host_central ? gt_ : 0.;
            FodS = gts_ * (4.5 * m_ * m_ * cold.getk(PB_HOST(P), K_R10)) * sig_h.v * inv_d6;
            FodP = gt_ * (4.5 * M_ * M_ * cold.get(K_R10)) * sig_p.v * inv_d6;
        }
#endif
        const double Fos = FodS * inv_d, Fop = FodP * inv_d;
        const double ks = Fos * inv_d, kp = Fop * inv_d;
        // radial: conservative + dissipative (-13.5 vr/r^8 (...) = -3 vr/r (Fos + Fop)), plus the radial part of the orthogonal one
        Kr = -inv_d * (bk_mgs.x * inv_d7 + (2.0 * inv_d) * ((Fos + Fop) * radvel));
        Kv = -(ks + kp);
        // F_orth (w x r - v) = -F_orth (r x w) - F_orth v
        F = v3(-(ks * cs.x + kp * cp.x), -(ks * cs.y + kp * cp.y), -(ks * cs.z + kp * cp.z));
        // torques (eqs 8-9 Bolmont+2015): N = Forth (d w - (r.w) r/d - (r x v)/d); dL/dt = -N
        const double krp = kp * rs_p, krs = ks * rs_s;
        dl_p = v3(krp * hr.x + kp * rxv.x - FodP * q.s.x, krp * hr.y + kp * rxv.y - FodP * q.s.y, krp * hr.z + kp * rxv.z - FodP * q.s.z);
        dl_h = v3(krs * hr.x + ks * rxv.x - FodS * sh.x, krs * hr.y + ks * rxv.y - FodS * sh.y, krs * hr.z + ks * rxv.z - FodS * sh.z);
        if (tide_save && ro.valid) {
            // internals that calculate_denergy_dt (tides/common.rs:263-279) will read at the next snapshot; warp-uniform
            // branch, taken on the last step before a snapshot or the end of a launch only
            const size_t ns = (size_t)P.n_sys;
            double* ts = P.tide_scratch + (size_t)b * ns + sys;
            const size_t cs_ = (size_t)PB_N(P) * ns;
            const double d = dot(hr, hr) * inv_d;
            ts[0 * cs_] = hr.x; ts[1 * cs_] = hr.y; ts[2 * cs_] = hr.z;
            ts[3 * cs_] = hv.x; ts[4 * cs_] = hv.y; ts[5 * cs_] = hv.z;
            ts[6 * cs_] = d; ts[7 * cs_] = radvel; ts[8 * cs_] = Fop;
            ts[9 * cs_] = -3.0 * Fop * radvel * inv_d;  // dissipative radial part with the star as a point mass
            ts[10 * cs_] = dl_p.x; ts[11 * cs_] = dl_p.y; ts[12 * cs_] = dl_p.z;
        }
    }
    if (PB_FLAGS(P) & FLAG_FLAT) {
        // oblate_spheroid.rs:12-97, rotational_flattening/common.rs:165-237
        const double inv_d5 = inv_d4 * inv_d;
        const double2 ks_kp = cold.get2(C_AS, 1);
        const double KsRs = ks_kp.x * rs_s, KpRp = ks_kp.y * rs_p;
        const double Fos = -KsRs * inv_d5;
        const double Fop = -KpRp * inv_d5;
        const double q1 = ks_kp.x * wh2 + ks_kp.y * w2;
        const double q2 = KsRs * rs_s + KpRp * rs_p;
        Kr += inv_d5 * ((2.5 * inv_d2) * q2 - 0.5 * q1);
        F = v3(F.x + Fop * q.s.x + Fos * sh.x, F.y + Fop * q.s.y + Fos * sh.y, F.z + Fop * q.s.z + Fos * sh.z);
        // torques F_orth (r x w); dL/dt = -N
        Pcp = -Fop;
        Hcs = -Fos;
    }
    if (GR == PB200_GR_KIDDER1995 && (PB_FLAGS(P) & FLAG_GR)) {
        // general_relativity.rs:177-456
        const double v2 = dot(hv, hv);
        const double mgs = bk_mgs.y;
        const double A = mgs * inv_d2 * kInvC2;
        const double u = mgs * inv_d;
        const double rv2 = radvel * radvel;
        // 1PN; the orthoradial term divides by |v| and multiplies by |v|: cancelled. Coefficients G_k: make_consts.
        const double2 g01 = cold.get2(G_0, 0), g23 = cold.get2(G_0, 1), g45 = cold.get2(G_0, 2), g67 = cold.get2(G_0, 3), g89 = cold.get2(G_0, 4), gab = cold.get2(G_0, 5);
        double rad = -A * (g01.x * v2 - g01.y * u - g23.x * rv2);
        double orth = A * g23.y * radvel;
        rad += -A * (g45.x * (u * u) + g45.y * (v2 * v2) + g67.x * (rv2 * rv2) - g67.y * rv2 * v2 - g89.x * u * v2 - g89.y * u * rv2);
        orth += 0.5 * A * radvel * (gab.x * v2 - gab.y * u - cold.get(G_0 + 12) * rv2);
        const double m = cold.get(K_M);
        Kr += m * (rad * inv_d);
        Kv += m * orth;
        // 1.5PN spin-orbit (:300-456) with n = r / d, n x v = (r x v) / d and the spin combinations of make_consts
        const double2 z_ph = cold.get2(C_AS, 2), d_ps = cold.get2(C_AS, 3), mfa_sxs = cold.get2(C_AS, 4);
        const double Ip = cold.get(K_I), Ih = cold.getk(PB_HOST(P), K_I), zp = z_ph.x, zh = z_ph.y;
        const V3 S = v3(Ip * q.s.x + Ih * sh.x, Ip * q.s.y + Ih * sh.y, Ip * q.s.z + Ih * sh.z);
        const V3 Z = v3(zp * q.s.x + zh * sh.x, zp * q.s.y + zh * sh.y, zp * q.s.z + zh * sh.z);
        const V3 A1 = S + Z;                                                   // 2S + msf
        const V3 A3 = A1 + S;                                                  // 3S + msf
        const V3 A7 = v3(3. * A1.x + S.x, 3. * A1.y + S.y, 3. * A1.z + S.z);   // 7S + 3 msf
        const double mfa = mfa_sxs.x;
        const double s1 = 6. * mfa * inv_d2, s3 = 3. * mfa * (radvel * inv_d);
        const V3 e2 = cross(hv, A7), e3 = cross(hr, A3);
        F = v3(F.x + s1 * (hr.x * rxv.x * A1.x) - mfa * e2.x + s3 * e3.x,
               F.y + s1 * (hr.y * rxv.y * A1.y) - mfa * e2.y + s3 * e3.y,
               F.z + s1 * (hr.z * rxv.z * A1.z) - mfa * e2.z + s3 * e3.z);
        // Kidder 1995 eqs 2.4a, 2.4b: dLs/dt = fms Lo x Ls - Lp x Ls + 3 (n.Lp) n x Ls, dLp/dt = fmp Lo x Lp + Lp x Ls + 3 (n.Ls) n x Lp
        const double sxs_k = mfa_sxs.y;
        const double c3 = 3. * sxs_k * inv_d2;
        Pcp += c3 * q.rs_s;
        Hcs += c3 * q.rs_p;
        const V3 wxw = cross(q.s, sh), jp = cross(rxv, q.s), js = cross(rxv, sh);
        const double dp1 = d_ps.x, ds1 = d_ps.y;
        dl_p = v3(dl_p.x + dp1 * jp.x + sxs_k * wxw.x, dl_p.y + dp1 * jp.y + sxs_k * wxw.y, dl_p.z + dp1 * jp.z + sxs_k * wxw.z);
        dl_h = v3(dl_h.x + ds1 * js.x - sxs_k * wxw.x, dl_h.y + ds1 * js.y - sxs_k * wxw.y, dl_h.z + ds1 * js.z - sxs_k * wxw.z);
    }
    F = v3(F.x + Kr * hr.x + Kv * hv.x, F.y + Kr * hr.y + Kv * hv.y, F.z + Kr * hr.z + Kv * hv.z);
    dl_p = v3(dl_p.x + Pcp * cp.x, dl_p.y + Pcp * cp.y, dl_p.z + Pcp * cp.z);
    dl_h = v3(dl_h.x + Hcs * cs.x, dl_h.y + Hcs * cs.y, dl_h.z + Hcs * cs.z);
    const V3 a_p = inv_m * F;
    const V3 a_h = (-inv_M) * F;
    // lanes that are not orbiting bodies carry zero constants (make_consts): their terms vanish; reduce onto the host
    // Transposed reduction through the exchange columns: every lane leaves its six contributions, lane c adds component
    // c over the group (columns visited in rotated order: conflict-free banks), the host lane collects the six totals.
    // (the totals of the previous evaluation and the host velocity in M_0 were read before the spin-exchange barrier above)
    cold.set3(M_0, a_h); cold.set3(M_3, dl_h);
    __syncwarp();
#if PB_FIXED_N == 8
    if (b < 6) {
        // lane c adds component c: columns visited in XOR order (body b ^ j at step j: conflict-free banks, one LOP3 per address)
        const unsigned row = dist_self(cold) + (unsigned)((M_0 + b) * PB_BLOCK * 8);
        const double x0 = dist_ld(row, 0), x1 = dist_ld(row, 1), x2 = dist_ld(row, 2), x3 = dist_ld(row, 3),
                     x4 = dist_ld(row, 4), x5 = dist_ld(row, 5), x6 = dist_ld(row, 6), x7 = dist_ld(row, 7);
        dist_st(row, 0, ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7)));   // only this lane reads column b of slot M_0 + b: no hazard
    }
#else
    for (int c = b; c < 6; c += W) {
        double t = 0.;
        for (int j = 0; j < W; j++) t += cold.getk((j + b) & (W - 1), M_0 + c);
        cold.set(M_0 + c, t);   // only this lane reads column b of slot M_0 + c: no hazard
    }
#endif
    __syncwarp();
    a_out = a_p; dl_out = dl_p;
    if (ro.host) {
        a_out = v3(cold.getk(0, M_0), cold.getk(1 & (W - 1), M_1), cold.getk(2 & (W - 1), M_2));
        dl_out = v3(cold.getk(3 & (W - 1), M_3), cold.getk(4 & (W - 1), M_4), cold.getk(5 & (W - 1), M_5));
    }
#if !PB_FIXED_N
    if (PB_FLAGS(P) & FLAG_WIND) dl_out = dl_out + plain(wind_dangular_momentum_dt(P, ro, cold, b, sys, strict(q.s), sd(w2)));
#endif
}


}  // namespace PB_NS
