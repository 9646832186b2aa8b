// forces_fast.cuh — PB200_ARITH_FAST perturbation forces (tides, flattening, GR Kidder1995) and their per-system constants.
// Included once per geometry specialisation (see pb200_api.cu), inside namespace PB_NS; no include guard on purpose.
#include "whfast_kernel.cuh"
#include "dyn_effects.cuh"

namespace PB_NS {
using namespace pb200;

// Derives the force constants from masses, radii and dissipation parameters (cold path: launch start and whenever a
// radius evolves). sigma / k2 are fetched from global memory here, they are not kept on chip.
__device__ __forceinline__ void make_consts(const KParams& P, const Roles& ro, const Cold& cold, int hl, int b, size_t sys) {
    double sigma = 0., k2t = 0., k2f = 0.;
    if (ro.valid) {
        const size_t i = (size_t)b * (size_t)P.n_sys + sys;
        sigma = P.sigma[i]; k2t = P.k2t[i]; k2f = P.k2f[i];
    }
    const double m = cold.get(K_M), mg = cold.get(K_MG), R = cold.get(K_R), I = cold.get(K_I);
    const double M = shfl(m, hl), Mg = shfl(mg, hl), Ih = shfl(I, hl);
    const double Rh5 = pow5(shfl(R, hl));
    const double R5 = pow5(R);
    const double sig_h = shfl(sigma, hl), k2t_h = shfl(k2t, hl), k2f_h = shfl(k2f, hl);
    const double m2 = m * m, M2 = M * M;
    cold.set(C_INVI, 1. / I);
    // Role gates are folded into the constants: a lane that is not an OrbitingBody of an effect (host slot, padding,
    // Disabled role) carries zeros, so the force code needs no per-lane branches or selects.
    const double gt = ro.t_on ? 1. : 0., gf = ro.f_on ? 1. : 0., gg = ro.g_on ? 1. : 0.;
    const double gts = P.tides_host_central ? gt : 0., gfs = P.flat_host_central ? gf : 0.;
    cold.set(C_AS, gts * (4.5 * m2 * (Rh5 * Rh5) * sig_h));         // 4.5 m^2 R*^10 sigma*   (constant_time_lag.rs:232-234)
    cold.set(C_AP, gt * (4.5 * M2 * (R5 * R5) * sigma));            // 4.5 M^2 R^10 sigma     (constant_time_lag.rs:243-245)
    cold.set(C_BK, gt * (3.0 * kK2 * (m2 * Rh5 * k2t_h + M2 * R5 * k2t))); // 3 K2 (m^2 R*^5 k2* + M^2 R^5 k2) (:283-285)
#if !PB_FIXED_N
    // dynamical tides: the same constants without sigma (the pair-dependent sigma multiplies them per evaluation)
    cold.set(D_0, gts * (4.5 * m2 * (Rh5 * Rh5)));
    cold.set(D_1, gt * (4.5 * M2 * (R5 * R5)));
#endif
    cold.set(C_KS, gfs * (m * k2f_h * Rh5));                        // flattening: m k2f* R*^5 (oblate_spheroid.rs:37)
    cold.set(C_KP, gf * (M * k2f * R5));                            //             M k2f R^5   (oblate_spheroid.rs:42)
    cold.set(C_INVM, 1. / m);
    const double mgs = Mg + mg;
    cold.set(C_MGS, gg * mgs);                                     // gated: A = mgs / (r^2 c^2) vanishes for non-GR lanes
    cold.set(C_FA, gg * (kG * kInvC2));                            // G / c^2 of the 1.5PN terms, gated
    cold.set(C_GRF, Mg * mg / (mgs * mgs));                       // general_relativity.rs:98
    cold.set(C_MURED, (M * m) / (M + m));                         // general_relativity.rs:383
    cold.set(C_MD, M - m);                                        // mass_factor * star_planet_mass (:319-321, 336)
    cold.set(C_MOM, m / M);
    cold.set(C_FMS, 2. + 1.5 * m / M);                            // :390
    cold.set(C_FMP, 2. + 1.5 * M / m);                            // :419
    // polynomials in the GR factor f of the 1PN / 2PN terms (general_relativity.rs:197-205, 256-268): per-system constants
    const double f = Mg * mg / (mgs * mgs), f2 = f * f;
    cold.set(G_0, 1.0 + 3.0 * f);
    cold.set(G_0 + 1, 2.0 * (2.0 + f));
    cold.set(G_0 + 2, 1.5 * f);
    cold.set(G_0 + 3, 2.0 * (2.0 - f));
    cold.set(G_0 + 4, 0.75 * (12.0 + 29.0 * f));
    cold.set(G_0 + 5, f * (3.0 - 4.0 * f));
    cold.set(G_0 + 6, 1.875 * f * (1.0 - 3.0 * f));
    cold.set(G_0 + 7, 1.5 * f * (3.0 - 4.0 * f));
    cold.set(G_0 + 8, 0.5 * f * (13.0 - 4.0 * f));
    cold.set(G_0 + 9, 2.0 + 25.0 * f + 2.0 * f2);
    cold.set(G_0 + 10, f * (15.0 + 4.0 * f));
    cold.set(G_0 + 11, 4.0 + 41.0 * f + 8.0 * f2);
    cold.set(G_0 + 12, 3.0 * f * (3.0 + 2.0 * f));
    __syncwarp();   // the host's column (1/M, inertia) is read by the other lanes
}

// ---------------------------------------------------------------------------------------------
// Universe::calculate_additional_effects for the lane's body at (hr, hv) with the current L
// (universe.rs:428-614). Returns the inertial additional acceleration and dL/dt of THIS body;
// host-lane values are the group reductions.
template <int GR>
__device__ __forceinline__ void additional_effects(const KParams& P, const Roles& ro, const Cold& cold, int hl, int b, size_t sys,
                                                   double t, bool evolve_now, Lane& q, V3 hr, double inv_d, V3 hv, V3& a_out,
                                                   V3& dl_out, bool tide_save) {
    const int W = PB_W(P);
    // Q3: r.omega uses the spins of the previous evaluation (universe.rs:429-430)
    // (the host's spin travels through the exchange triple E_S: still the previous evaluation's here)
    V3 s_host_prev = cold.getk3(PB_HOST(P), E_S);
    double rs_s = dot(hr, s_host_prev), rs_p = dot(hr, q.s);
    // calculate_spin (particles/common.rs:3-15)
    q.s = cold.get(C_INVI) * q.L;
    double w2 = dot(q.s, q.s);
    __syncwarp();
    cold.set3(E_S, q.s); cold.set(M_6, w2);
    __syncwarp();
    V3 sh = cold.getk3(PB_HOST(P), E_S);
    double wh2 = cold.getk(PB_HOST(P), M_6);
#if !PB_FIXED_N
    // lag angle of the dynamical-tide models, once per step like the other evolving quantities (evolution.rs:548-567)
    if (evolve_now && (PB_FLAGS(P) & FLAG_DYN) && (PB_FLAGS(P) & FLAG_EVO)) { update_lag_angle(P, ro, b, sys, t, sd(w2), true); __syncwarp(); }
#endif
    double inv_d2 = inv_d * inv_d;
    double d = dot(hr, hr) * inv_d;
    double radvel = dot(hr, hv) * inv_d;
    V3 rxv = cross(hr, hv);
    V3 a_p = v3(0., 0., 0.), dl_p = v3(0., 0., 0.);       // this body's own acceleration / torque
    V3 a_h = v3(0., 0., 0.), dl_h = v3(0., 0., 0.);       // contribution to the host
    const double inv_m = cold.get(C_INVM), inv_M = cold.getk(PB_HOST(P), C_INVM);
    if (PB_FLAGS(P) & FLAG_TIDES) {
        // constant_time_lag.rs:206-332, tides/common.rs:223-345
        double inv_d4 = inv_d2 * inv_d2;
        double inv_d7 = inv_d4 * inv_d2 * inv_d;
        double Fos = cold.get(C_AS) * inv_d7;
        double Fop = cold.get(C_AP) * inv_d7;
#if !PB_FIXED_N
        if (PB_FLAGS(P) & FLAG_DYN) {
            sd sig_h, sig_p;
            pair_dependent_sigmas(P, ro, cold, hl, b, sys, strict(hr), strict(hv), sd(w2), sd(wh2), sig_h, sig_p);
            Fos = cold.get(D_0) * sig_h.v * inv_d7;
            Fop = cold.get(D_1) * sig_p.v * inv_d7;
        }
#endif
        double Fsum = Fos + Fop;
        // radial: conservative + dissipative (-13.5 vr/r^8 (...) = -3 vr/r (Fos + Fop))
        double f3 = -cold.get(C_BK) * inv_d7 - 2.0 * Fsum * radvel * inv_d;
        V3 wxr_s = cross(sh, hr), wxr_p = cross(q.s, hr);
        double k3 = f3 * inv_d, ks = Fos * inv_d, kp = Fop * inv_d;
        V3 F = v3(k3 * hr.x + ks * (wxr_s.x - hv.x) + kp * (wxr_p.x - hv.x),
                  k3 * hr.y + ks * (wxr_s.y - hv.y) + kp * (wxr_p.y - hv.y),
                  k3 * hr.z + ks * (wxr_s.z - hv.z) + kp * (wxr_p.z - hv.z));
        // torques (eqs 8-9 Bolmont+2015): N = Forth (d w - (r.w) r/d - (r x v)/d); dL/dt = -N
        V3 Np = v3(Fop * (d * q.s.x - (rs_p * hr.x + rxv.x) * inv_d), Fop * (d * q.s.y - (rs_p * hr.y + rxv.y) * inv_d),
                   Fop * (d * q.s.z - (rs_p * hr.z + rxv.z) * inv_d));
        V3 Ns = v3(Fos * (d * sh.x - (rs_s * hr.x + rxv.x) * inv_d), Fos * (d * sh.y - (rs_s * hr.y + rxv.y) * inv_d),
                   Fos * (d * sh.z - (rs_s * hr.z + rxv.z) * inv_d));
        a_p = inv_m * F;
        a_h = (-inv_M) * F;
        dl_p = v3(-Np.x, -Np.y, -Np.z);
        dl_h = v3(-Ns.x, -Ns.y, -Ns.z);
        if (tide_save && ro.valid) {
            // internals that calculate_denergy_dt (tides/common.rs:263-279) will read at the next snapshot; warp-uniform
            // branch, taken on the last step before a snapshot or the end of a launch only
            const size_t ns = (size_t)P.n_sys;
            double* ts = P.tide_scratch + (size_t)b * ns + sys;
            const size_t cs = (size_t)PB_N(P) * ns;
            ts[0 * cs] = hr.x; ts[1 * cs] = hr.y; ts[2 * cs] = hr.z;
            ts[3 * cs] = hv.x; ts[4 * cs] = hv.y; ts[5 * cs] = hv.z;
            ts[6 * cs] = d; ts[7 * cs] = radvel; ts[8 * cs] = Fop;
            ts[9 * cs] = -3.0 * Fop * radvel * inv_d;  // dissipative radial part with the star as a point mass
            ts[10 * cs] = -Np.x; ts[11 * cs] = -Np.y; ts[12 * cs] = -Np.z;
        }
    }
    if (PB_FLAGS(P) & FLAG_FLAT) {
        // oblate_spheroid.rs:12-97, rotational_flattening/common.rs:165-237
        double inv_d5 = inv_d2 * inv_d2 * inv_d;
        double inv_d7 = inv_d5 * inv_d2;
        double Ks = cold.get(C_KS);
        double Kp = cold.get(C_KP);
        double Fos = -Ks * rs_s * inv_d5;
        double Fop = -Kp * rs_p * inv_d5;
        double Frad = -0.5 * inv_d5 * (Ks * wh2 + Kp * w2) + 2.5 * inv_d7 * (Ks * rs_s * rs_s + Kp * rs_p * rs_p);
        V3 F = v3(Frad * hr.x + Fop * q.s.x + Fos * sh.x, Frad * hr.y + Fop * q.s.y + Fos * sh.y, Frad * hr.z + Fop * q.s.z + Fos * sh.z);
        V3 Np = Fop * cross(hr, q.s);
        V3 Ns = Fos * cross(hr, sh);
        a_p = a_p + inv_m * F;
        a_h = a_h - inv_M * F;
        dl_p = dl_p - Np;
        dl_h = dl_h - Ns;
    }
    if (GR == PB200_GR_KIDDER1995) {
        // general_relativity.rs:177-456
        double v2 = dot(hv, hv);
        double mgs = cold.get(C_MGS);
        double A = mgs * inv_d2 * kInvC2;
        double u = mgs * inv_d;
        double rv2 = radvel * radvel;
        // 1PN; the orthoradial term divides by |v| and multiplies by |v|: cancelled. Coefficients G_k: make_consts.
        double rad = -A * (cold.get(G_0) * v2 - cold.get(G_0 + 1) * u - cold.get(G_0 + 2) * rv2);
        double orth = A * cold.get(G_0 + 3) * radvel;
        // 2PN (Kidder 1995 eq. 2.2d)
        rad += -A * (cold.get(G_0 + 4) * (u * u) + cold.get(G_0 + 5) * (v2 * v2) + cold.get(G_0 + 6) * (rv2 * rv2)
                     - cold.get(G_0 + 7) * rv2 * v2 - cold.get(G_0 + 8) * u * v2 - cold.get(G_0 + 9) * u * rv2);
        orth += 0.5 * A * radvel * (cold.get(G_0 + 10) * v2 - cold.get(G_0 + 11) * u - cold.get(G_0 + 12) * rv2);
        double kr = rad * inv_d;
        V3 a = v3(kr * hr.x + orth * hv.x, kr * hr.y + orth * hv.y, kr * hr.z + orth * hv.z);
        // 1.5PN spin-orbit (:300-456); component-wise products exactly as the reference writes them
        V3 Ls = cold.getk(PB_HOST(P), K_I) * sh, Lp = cold.get(K_I) * q.s;
        V3 nn = inv_d * hr;
        double md = cold.get(C_MD);
        V3 msf = v3(md * (Lp.x * inv_m - Ls.x * inv_M), md * (Lp.y * inv_m - Ls.y * inv_M), md * (Lp.z * inv_m - Ls.z * inv_M));
        V3 S = Ls + Lp;
        V3 nxv = cross(nn, hv);
        V3 e1 = v3(6. * nn.x * (nxv.x * (2. * S.x + msf.x)), 6. * nn.y * (nxv.y * (2. * S.y + msf.y)), 6. * nn.z * (nxv.z * (2. * S.z + msf.z)));
        V3 e7 = v3(7. * S.x + 3. * msf.x, 7. * S.y + 3. * msf.y, 7. * S.z + 3. * msf.z);
        V3 e2 = cross(hv, e7);
        V3 e3s = v3(3. * S.x + msf.x, 3. * S.y + msf.y, 3. * S.z + msf.z);
        V3 e3 = (3. * radvel) * cross(nn, e3s);
        const double fa = cold.get(C_FA);
        a = a + fa * (e1 - e2 + e3);
        // Kidder 1995 eqs 2.4a, 2.4b
        V3 Lo = cold.get(C_MURED) * rxv;
        V3 LpxLs = cross(Lp, Ls);
        V3 ds = cold.get(C_FMS) * cross(Lo, Ls) - LpxLs + (3. * dot(nn, Lp)) * cross(nn, Ls);
        V3 dp = cold.get(C_FMP) * cross(Lo, Lp) + LpxLs + (3. * dot(nn, Ls)) * cross(nn, Lp);
        a_p = a_p + a;
        a_h = a_h - cold.get(C_MOM) * a;
        dl_p = dl_p + fa * dp;
        dl_h = dl_h + fa * ds;
    }
    // lanes that are not orbiting bodies carry zero constants (make_consts): their terms vanish; reduce onto the host
    // Transposed reduction through the exchange columns: every lane leaves its six contributions, lane c adds component
    // c over the group (columns visited in rotated order: conflict-free banks), the host lane collects the six totals.
    __syncwarp();   // the totals of the previous evaluation have been read
    cold.set3(M_0, a_h); cold.set3(M_3, dl_h);
    __syncwarp();
    for (int c = b; c < 6; c += W) {
        double t;
        if (PB_FIXED_N == 8) {
            const double x0 = cold.getk(b, M_0 + c), x1 = cold.getk((b + 1) & 7, M_0 + c), x2 = cold.getk((b + 2) & 7, M_0 + c),
                         x3 = cold.getk((b + 3) & 7, M_0 + c), x4 = cold.getk((b + 4) & 7, M_0 + c), x5 = cold.getk((b + 5) & 7, M_0 + c),
                         x6 = cold.getk((b + 6) & 7, M_0 + c), x7 = cold.getk((b + 7) & 7, M_0 + c);
            t = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
        } else {
            t = 0.;
            for (int j = 0; j < W; j++) t += cold.getk((j + b) & (W - 1), M_0 + c);
        }
        cold.set(M_0 + c, t);   // only this lane reads column b of slot M_0 + c: no hazard
    }
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
