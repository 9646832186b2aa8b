// strict_effects.cuh — PB200_ARITH_STRICT: the perturbation forces in the reference's own arithmetic.
//
// Same physics as additional_effects() in whfast_kernel.cuh, but every operation is an IEEE round-to-nearest
// add/mul/div/sqrt in the association order of the reference source (transcribed from the CPU oracle, which is
// bit-exact against the reference's golden vectors), sums over planets are accumulated serially in body order, and
// `powi` follows LLVM's square-and-multiply expansion. Together with the strict core this makes the whole step
// bit-reproducible against the oracle (tests/test_gpu_parity.py::test_strict_mode_is_bit_identical...).
// Cost: ~50 true divisions per planet per evaluation instead of 1 rsqrt — about half the throughput of the fast mode.

namespace PB_NS {
using namespace pb200;

// strict-mode constants overlay the fast-mode constant slots (C_INVI..C_FA, 15 slots) plus Z_0; the host's 1/M, R^5 and
// moment of inertia are read from the host's own column
enum StrictSlot : int {
    Z_CS = C_INVI, Z_CP, Z_KCONS, Z_T1, Z_T2, Z_INVM, Z_FS0, Z_FP0, Z_R5, Z_MGS, Z_GRF, Z_MOM, Z_MFM, Z_MURED, Z_FMS,
    Z_FMP = Z_0
};

__device__ __forceinline__ sd spow5(sd x) { sd x2 = x * x; sd x4 = x2 * x2; return x * x4; }            // x * x^4
__device__ __forceinline__ sd spow10(sd x) { sd x2 = x * x; sd x4 = x2 * x2; sd x8 = x4 * x4; return x2 * x8; }  // x^2 * x^8

__device__ __forceinline__ void make_consts_strict(const KParams& P, const Roles& ro, const Cold& cold, int hl, int b, size_t sys) {
    double sigma = 0., k2t = 0., k2f = 0.;
    if (ro.valid) {
        const size_t i = (size_t)b * (size_t)P.n_sys + sys;
        sigma = P.sigma[i]; k2t = P.k2t[i]; k2f = P.k2f[i];
    }
    const sd m = sd(cold.get(K_M)), mg = sd(cold.get(K_MG)), R = sd(cold.get(K_R));
    const sd M = sd(shfl(m.v, hl)), Mg = sd(shfl(mg.v, hl)), Rh = sd(shfl(R.v, hl));
    const sd sig_h = sd(shfl(sigma, hl)), k2t_h = sd(shfl(k2t, hl)), k2f_h = sd(shfl(k2f, hl));
    const sd m2 = m * m, M2 = M * M;
    const sd Rh5 = spow5(Rh), R5 = spow5(R), Rh10 = spow10(Rh), R10 = spow10(R);
    cold.set(Z_CS, (sd(4.5) * m2 * Rh10 * sig_h).v);            // constant_time_lag.rs:232-234 numerator
    cold.set(Z_CP, (sd(4.5) * M2 * R10 * sd(sigma)).v);         // :243-245 numerator
    cold.set(Z_KCONS, (m2 * Rh5 * k2t_h + M2 * R5 * sd(k2t)).v);  // :284-285
    cold.set(Z_T1, (m2 * Rh10 * sig_h).v);                      // :291-293
    cold.set(Z_T2, (M2 * R10 * sd(sigma)).v);                   // :294-296
#if !PB_FIXED_N
    cold.set(D_0, (sd(4.5) * m2 * Rh10).v); cold.set(D_1, (sd(4.5) * M2 * R10).v);   // the same products up to sigma
    cold.set(D_2, (m2 * Rh10).v); cold.set(D_3, (M2 * R10).v);
#endif
    cold.set(Z_INVM, (sd(1.) / m).v);
    cold.set(Z_FS0, (m * k2f_h).v); cold.set(Z_FP0, (M * sd(k2f)).v);   // oblate_spheroid.rs:37, 42 leading products
    cold.set(Z_R5, R5.v);
    const sd mgs = Mg + mg;
    cold.set(Z_MGS, mgs.v);
    cold.set(Z_GRF, (Mg * mg / (mgs * mgs)).v);                 // general_relativity.rs:98
    cold.set(Z_MOM, (m / M).v);                                 // particle.mass / host.mass (:216)
    const sd msum = M + m, mdiff = M - m;
    cold.set(Z_MFM, (mdiff / msum * msum).v);                   // mass_factor * star_planet_mass (:321, 336)
    cold.set(Z_MURED, ((M * m) / msum).v);                      // :383
    cold.set(Z_FMS, (sd(2.) + sd(3.) / sd(2.) * m / M).v);      // :390
    cold.set(Z_FMP, (sd(2.) + sd(3.) / sd(2.) * M / m).v);      // :419
    __syncwarp();   // the host's column is read by the other lanes
}

// The host lane accumulates, in body order, the terms that the other lanes left in the exchange slots.
__device__ __forceinline__ S3 host_ordered_sum(const Cold& cold, int slot, int b, int n, int host) {
    S3 acc = s3(sd(0.), sd(0.), sd(0.));
    for (int k = 0; k < n; k++) {
        if (k == host) continue;
        const volatile double* p = cold.base + (k - b);   // lane gb + k of the same group
        acc.x = acc.x + sd(p[(slot + 0) * PB_BLOCK]);
        acc.y = acc.y + sd(p[(slot + 1) * PB_BLOCK]);
        acc.z = acc.z + sd(p[(slot + 2) * PB_BLOCK]);
    }
    return acc;
}
__device__ __forceinline__ void put3(const Cold& cold, int slot, S3 v) { cold.set(slot, v.x.v); cold.set(slot + 1, v.y.v); cold.set(slot + 2, v.z.v); }

template <int GR>
__device__ __forceinline__ void additional_effects_strict(const KParams& P, const Roles& ro, const Cold& cold, int hl, int b, size_t sys,
                                                          double t, bool evolve_now, Lane& q, S3 hr, sd dist, S3 hv, V3& a_out,
                                                          V3& dl_out, bool tide_save) {
    const int n = PB_N(P);
    const sd zero = sd(0.);
    // Q3: r.omega with the spins of the previous evaluation (tides/common.rs:155-160 = rotational_flattening/common.rs:105-110)
    const S3 sp_prev = strict(q.s), sh_prev = shfl3(sp_prev, hl);
    const sd rs_s = hr.x * sh_prev.x + hr.y * sh_prev.y + hr.z * sh_prev.z;
    const sd rs_p = hr.x * sp_prev.x + hr.y * sp_prev.y + hr.z * sp_prev.z;
    // calculate_spin (particles/common.rs:3-15)
    const sd I = sd(cold.get(K_I));
    const S3 L = strict(q.L);
    const S3 s = s3(L.x / I, L.y / I, L.z / I);
    const sd w2 = (s.x * s.x) + (s.y * s.y) + (s.z * s.z);
    q.s = plain(s);
    const S3 sh = shfl3(s, hl);
    const sd wh2 = shfl(w2, hl);
#if !PB_FIXED_N
    if (evolve_now && (PB_FLAGS(P) & FLAG_DYN) && (PB_FLAGS(P) & FLAG_EVO)) { update_lag_angle(P, ro, b, sys, t, w2, true); __syncwarp(); }
#endif
    // inertial_to_heliocentric (universe.rs:331-338); the host's stale heliocentric velocity is zero (validated)
    const sd radvel = (hr.x * hv.x + hr.y * hv.y + hr.z * hv.z) / dist;
    const sd normv2 = hv.x * hv.x + hv.y * hv.y + hv.z * hv.z;
    const sd d2 = dist * dist, d4 = d2 * d2;
    const sd d5 = dist * d4, d7 = (dist * d2) * d4, d8 = d4 * d4;   // powi as LLVM expands it
    const sd inv_m = sd(cold.get(Z_INVM)), inv_M = sd(cold.getk(PB_HOST(P), Z_INVM));
    S3 t_acc = s3(zero, zero, zero), t_dl = t_acc, f_acc = t_acc, f_dl = t_acc, g_acc = t_acc, g_dl = t_acc;
    // terms for the host, exchanged through shared memory: X_0.. = tides F, tides -N_s, flattening F, flattening -N_s
    S3 xF_t = t_acc, xN_t = t_acc, xF_f = t_acc, xN_f = t_acc;
    if (PB_FLAGS(P) & FLAG_TIDES) {
        sd cs = sd(cold.get(Z_CS)), cp = sd(cold.get(Z_CP)), t1 = sd(cold.get(Z_T1)), t2 = sd(cold.get(Z_T2));
#if !PB_FIXED_N
        if (PB_FLAGS(P) & FLAG_DYN) {
            // sigma is the last factor of each product in the reference, so multiplying it in here rounds identically
            sd sig_h, sig_p;
            pair_dependent_sigmas(P, ro, cold, hl, b, sys, hr, hv, w2, wh2, sig_h, sig_p);
            cs = sd(cold.get(D_0)) * sig_h; cp = sd(cold.get(D_1)) * sig_p;
            t1 = sd(cold.get(D_2)) * sig_h; t2 = sd(cold.get(D_3)) * sig_p;
        }
#endif
        const sd orth_s = P.tides_host_central ? cs / d7 : zero;
        const sd orth_p = cp / d7;
        const sd host_k = sd(cold.get(Z_KCONS));
        const sd cons = sd(-3.0 * kK2) / d7 * host_k;
        const sd factor1 = sd(-13.5) * radvel / d8;
        const sd diss_pm = factor1 * t2;
        const sd diss = diss_pm + factor1 * t1;
        const sd t_radial = cons + diss;
        const sd f3 = t_radial + (orth_s + orth_p) * radvel / dist;
        const sd osd = orth_s / dist, opd = orth_p / dist;
        S3 F;
        F.x = f3 * hr.x / dist + osd * (sh.y * hr.z - sh.z * hr.y - hv.x) + opd * (s.y * hr.z - s.z * hr.y - hv.x);
        F.y = f3 * hr.y / dist + osd * (sh.z * hr.x - sh.x * hr.z - hv.y) + opd * (s.z * hr.x - s.x * hr.z - hv.y);
        F.z = f3 * hr.z / dist + osd * (sh.x * hr.y - sh.y * hr.x - hv.z) + opd * (s.x * hr.y - s.y * hr.x - hv.z);
        const sd oned = sd(1.0) / dist;
        const sd cx = hr.y * hv.z - hr.z * hv.y, cy = hr.z * hv.x - hr.x * hv.z, cz = hr.x * hv.y - hr.y * hv.x;
        S3 Np, Ns;
        Np.x = orth_p * (dist * s.x - rs_p * hr.x / dist - oned * cx);
        Np.y = orth_p * (dist * s.y - rs_p * hr.y / dist - oned * cy);
        Np.z = orth_p * (dist * s.z - rs_p * hr.z / dist - oned * cz);
        Ns.x = orth_s * (dist * sh.x - rs_s * hr.x / dist - oned * cx);
        Ns.y = orth_s * (dist * sh.y - rs_s * hr.y / dist - oned * cy);
        Ns.z = orth_s * (dist * sh.z - rs_s * hr.z / dist - oned * cz);
        if (ro.t_on) {
            t_acc = s3(inv_m * F.x, inv_m * F.y, inv_m * F.z);
            t_dl = s3(sd(-1.0) * Np.x, sd(-1.0) * Np.y, sd(-1.0) * Np.z);
            xF_t = F;
            xN_t = s3(sd(-1.0) * Ns.x, sd(-1.0) * Ns.y, sd(-1.0) * Ns.z);
        }
        if (tide_save && ro.valid) {
            const size_t ns = (size_t)P.n_sys;
            double* ts = P.tide_scratch + (size_t)b * ns + sys;
            const size_t cs = (size_t)PB_N(P) * ns;
            ts[0 * cs] = hr.x.v; ts[1 * cs] = hr.y.v; ts[2 * cs] = hr.z.v;
            ts[3 * cs] = hv.x.v; ts[4 * cs] = hv.y.v; ts[5 * cs] = hv.z.v;
            ts[6 * cs] = dist.v; ts[7 * cs] = radvel.v; ts[8 * cs] = orth_p.v; ts[9 * cs] = diss_pm.v;
            ts[10 * cs] = t_dl.x.v; ts[11 * cs] = t_dl.y.v; ts[12 * cs] = t_dl.z.v;
        }
    }
    if (PB_FLAGS(P) & FLAG_FLAT) {
        const sd Rh5 = sd(cold.getk(PB_HOST(P), Z_R5)), R5 = sd(cold.get(Z_R5));
        const sd ffs = P.flat_host_central ? sd(cold.get(Z_FS0)) * wh2 * Rh5 / sd(6.) : zero * wh2 * Rh5 / sd(6.);
        const sd orth_s = sd(-6.) * ffs * rs_s / (wh2 * d5);
        const sd ffp = sd(cold.get(Z_FP0)) * w2 * R5 / sd(6.);
        const sd orth_p = sd(-6.) * ffp * rs_p / (w2 * d5);
        const sd radial = sd(-3.) / d5 * (ffp + ffs) + sd(15.) / d7 * (ffs * rs_s * rs_s / wh2 + ffp * rs_p * rs_p / w2);
        S3 F;
        F.x = radial * hr.x + orth_p * s.x + orth_s * sh.x;
        F.y = radial * hr.y + orth_p * s.y + orth_s * sh.y;
        F.z = radial * hr.z + orth_p * s.z + orth_s * sh.z;
        S3 Np, Ns;
        Np.x = orth_p * (hr.y * s.z - hr.z * s.y); Np.y = orth_p * (hr.z * s.x - hr.x * s.z); Np.z = orth_p * (hr.x * s.y - hr.y * s.x);
        Ns.x = orth_s * (hr.y * sh.z - hr.z * sh.y); Ns.y = orth_s * (hr.z * sh.x - hr.x * sh.z); Ns.z = orth_s * (hr.x * sh.y - hr.y * sh.x);
        if (ro.f_on) {
            f_acc = s3(inv_m * F.x, inv_m * F.y, inv_m * F.z);
            f_dl = s3(sd(-1.0) * Np.x, sd(-1.0) * Np.y, sd(-1.0) * Np.z);
            xF_f = F;
            xN_f = s3(sd(-1.0) * Ns.x, sd(-1.0) * Ns.y, sd(-1.0) * Ns.z);
        }
    }
    // ---- exchange round A: host sums of tides / flattening
    S3 h_t_acc = t_acc, h_t_dl = t_acc, h_f_acc = t_acc, h_f_dl = t_acc;
    if (PB_FLAGS(P) & (FLAG_TIDES | FLAG_FLAT)) {
        put3(cold, X_0, xF_t); put3(cold, X_0 + 3, xN_t); put3(cold, X_0 + 6, xF_f); put3(cold, X_0 + 9, xN_f);
        __syncwarp();
        if (ro.host) {
            const sd neg_inv_M = sd(-1.0) * inv_M;   // -1.0 * factor2
            S3 sF = host_ordered_sum(cold, X_0, b, n, PB_HOST(P));
            h_t_acc = s3(neg_inv_M * sF.x, neg_inv_M * sF.y, neg_inv_M * sF.z);
            h_t_dl = host_ordered_sum(cold, X_0 + 3, b, n, PB_HOST(P));
            S3 sG = host_ordered_sum(cold, X_0 + 6, b, n, PB_HOST(P));
            h_f_acc = s3(neg_inv_M * sG.x, neg_inv_M * sG.y, neg_inv_M * sG.z);
            h_f_dl = host_ordered_sum(cold, X_0 + 9, b, n, PB_HOST(P));
        }
        __syncwarp();
    }
    S3 h_g_acc = g_acc, h_g_dl = g_acc;
    if (GR == PB200_GR_KIDDER1995) {
        // general_relativity.rs:177-456, transcribed from the oracle (oracle_core.hpp gr_kidder)
        const sd c2 = sd(kC2);
        const sd mgs = sd(cold.get(Z_MGS)), f = sd(cold.get(Z_GRF)), mom = sd(cold.get(Z_MOM));
        const sd normv = ssqrt(normv2);
        const sd rv2 = radvel * radvel;
        const sd pre = -mgs / (d2 * c2);
        // 1PN
        const sd radial1 = pre * ((sd(1.0) + sd(3.0) * f) * normv2 - sd(2.0) * (sd(2.0) + f) * mgs / dist - sd(1.5) * f * rv2);
        const sd orth1 = mgs / (d2 * c2) * sd(2.0) * (sd(2.0) - f) * radvel * normv;
        S3 a1;
        a1.x = radial1 * hr.x / dist + orth1 * hv.x / normv;
        a1.y = radial1 * hr.y / dist + orth1 * hv.y / normv;
        a1.z = radial1 * hr.z / dist + orth1 * hv.z / normv;
        // 2PN
        const sd v4 = normv2 * normv2, rv4 = rv2 * rv2, f2 = f * f;
        const sd mgd = mgs / dist;
        const sd radial2 = pre
            * (sd(3.0) / sd(4.0) * (sd(12.0) + sd(29.0) * f) * (mgs * mgs / d2)
               + f * (sd(3.0) - sd(4.0) * f) * v4
               + sd(15.0) / sd(8.0) * f * (sd(1.0) - sd(3.0) * f) * rv4
               - sd(3.0) / sd(2.0) * f * (sd(3.0) - sd(4.0) * f) * rv2 * normv2
               - sd(0.5) * f * (sd(13.0) - sd(4.0) * f) * mgd * normv2
               - (sd(2.0) + sd(25.0) * f + sd(2.0) * f2) * mgd * rv2);
        const sd orth2 = pre * sd(-0.5) * radvel
            * (f * (sd(15.0) + sd(4.0) * f) * normv2 - (sd(4.0) + sd(41.0) * f + sd(8.0) * f2) * mgd - sd(3.0) * f * (sd(3.0) + sd(2.0) * f) * rv2);
        S3 a2;
        a2.x = radial2 * hr.x / dist + orth2 * hv.x;
        a2.y = radial2 * hr.y / dist + orth2 * hv.y;
        a2.z = radial2 * hr.z / dist + orth2 * hv.z;
        // 1.5PN spin-orbit
        const sd Ih = sd(cold.getk(PB_HOST(P), K_I));
        const sd M = sd(cold.getk(PB_HOST(P), K_M)), m = sd(cold.get(K_M));
        const S3 Ls = s3(Ih * sh.x, Ih * sh.y, Ih * sh.z), Lp = s3(I * s.x, I * s.y, I * s.z);
        const S3 nn = s3(hr.x / dist, hr.y / dist, hr.z / dist);
        const sd mfm = sd(cold.get(Z_MFM));
        const sd msx = mfm * (Lp.x / m - Ls.x / M), msy = mfm * (Lp.y / m - Ls.y / M), msz = mfm * (Lp.z / m - Ls.z / M);
        const sd e1x = sd(6.) * nn.x * ((nn.y * hv.z - nn.z * hv.y) * (sd(2.) * (Ls.x + Lp.x) + msx));
        const sd e1y = sd(6.) * nn.y * ((nn.z * hv.x - nn.x * hv.z) * (sd(2.) * (Ls.y + Lp.y) + msy));
        const sd e1z = sd(6.) * nn.z * ((nn.x * hv.y - nn.y * hv.x) * (sd(2.) * (Ls.z + Lp.z) + msz));
        const sd e7x = sd(7.) * (Ls.x + Lp.x) + sd(3.) * msx, e7y = sd(7.) * (Ls.y + Lp.y) + sd(3.) * msy, e7z = sd(7.) * (Ls.z + Lp.z) + sd(3.) * msz;
        const sd e2x = hv.y * e7z - hv.z * e7y, e2y = hv.z * e7x - hv.x * e7z, e2z = hv.x * e7y - hv.y * e7x;
        const sd e3sx = sd(3.) * (Ls.x + Lp.x) + msx, e3sy = sd(3.) * (Ls.y + Lp.y) + msy, e3sz = sd(3.) * (Ls.z + Lp.z) + msz;
        const sd e3x = sd(3.) * radvel * (nn.y * e3sz - nn.z * e3sy);
        const sd e3y = sd(3.) * radvel * (nn.z * e3sx - nn.x * e3sz);
        const sd e3z = sd(3.) * radvel * (nn.x * e3sy - nn.y * e3sx);
        const sd fa = sd(kG) / c2;
        S3 a3 = s3(fa * (e1x - e2x + e3x), fa * (e1y - e2y + e3y), fa * (e1z - e2z + e3z));
        // Kidder 1995 eq. 2.4a / 2.4b
        const sd mu = sd(cold.get(Z_MURED));
        const S3 Lo = s3(mu * (hr.y * hv.z - hr.z * hv.y), mu * (hr.z * hv.x - hr.x * hv.z), mu * (hr.x * hv.y - hr.y * hv.x));
        const sd fms = sd(cold.get(Z_FMS)), fmp = sd(cold.get(Z_FMP));
        const sd a1x = fms * (Lo.y * Ls.z - Lo.z * Ls.y), a1y = fms * (Lo.z * Ls.x - Lo.x * Ls.z), a1z = fms * (Lo.x * Ls.y - Lo.y * Ls.x);
        const sd a2x = Lp.y * Ls.z - Lp.z * Ls.y, a2y = Lp.z * Ls.x - Lp.x * Ls.z, a2z = Lp.x * Ls.y - Lp.y * Ls.x;
        const sd spp = nn.x * Lp.x + nn.y * Lp.y + nn.z * Lp.z;
        const sd a3x = sd(3.) * spp * (nn.y * Ls.z - nn.z * Ls.y), a3y = sd(3.) * spp * (nn.z * Ls.x - nn.x * Ls.z), a3z = sd(3.) * spp * (nn.x * Ls.y - nn.y * Ls.x);
        const S3 hdl = s3(fa * (a1x - a2x + a3x), fa * (a1y - a2y + a3y), fa * (a1z - a2z + a3z));
        const sd b1x = fmp * (Lo.y * Lp.z - Lo.z * Lp.y), b1y = fmp * (Lo.z * Lp.x - Lo.x * Lp.z), b1z = fmp * (Lo.x * Lp.y - Lo.y * Lp.x);
        const sd b2x = Ls.y * Lp.z - Ls.z * Lp.y, b2y = Ls.z * Lp.x - Ls.x * Lp.z, b2z = Ls.x * Lp.y - Ls.y * Lp.x;
        const sd ssp = nn.x * Ls.x + nn.y * Ls.y + nn.z * Ls.z;
        const sd b3x = sd(3.) * ssp * (nn.y * Lp.z - nn.z * Lp.y), b3y = sd(3.) * ssp * (nn.z * Lp.x - nn.x * Lp.z), b3z = sd(3.) * ssp * (nn.x * Lp.y - nn.y * Lp.x);
        S3 x1 = s3(zero, zero, zero), x2 = x1, x3 = x1, x4 = x1;
        if (ro.g_on) {
            g_acc = s3(a1.x + a2.x + a3.x, a1.y + a2.y + a3.y, a1.z + a2.z + a3.z);
            g_dl = s3(fa * (b1x - b2x + b3x), fa * (b1y - b2y + b3y), fa * (b1z - b2z + b3z));
            x1 = s3(mom * a1.x, mom * a1.y, mom * a1.z);
            x2 = s3(mom * a2.x, mom * a2.y, mom * a2.z);
            x3 = s3(mom * a3.x, mom * a3.y, mom * a3.z);
            x4 = hdl;
        }
        // ---- exchange round B: the three host acceleration sums and the host torque
        put3(cold, X_0, x1); put3(cold, X_0 + 3, x2); put3(cold, X_0 + 6, x3); put3(cold, X_0 + 9, x4);
        __syncwarp();
        if (ro.host) {
            S3 s1 = host_ordered_sum(cold, X_0, b, n, PB_HOST(P)), s2 = host_ordered_sum(cold, X_0 + 3, b, n, PB_HOST(P));
            S3 s3_ = host_ordered_sum(cold, X_0 + 6, b, n, PB_HOST(P));
            const sd m1 = sd(-1.0);
            h_g_acc = s3(m1 * s1.x + m1 * s2.x + m1 * s3_.x, m1 * s1.y + m1 * s2.y + m1 * s3_.y, m1 * s1.z + m1 * s2.z + m1 * s3_.z);
            h_g_dl = host_ordered_sum(cold, X_0 + 9, b, n, PB_HOST(P));
        }
        __syncwarp();
    }
    // add_additional_acceleration_corrections / calculate_dangular_momentum_dt (universe.rs:540-614)
    const S3 ta = ro.host ? h_t_acc : t_acc, fa_ = ro.host ? h_f_acc : f_acc, ga = ro.host ? h_g_acc : g_acc;
    const S3 td = ro.host ? h_t_dl : t_dl, fd = ro.host ? h_f_dl : f_dl, gd = ro.host ? h_g_dl : g_dl;
    S3 a = s3(zero, zero, zero);
    if (PB_FLAGS(P) & FLAG_TIDES) a = a + ta;
    if (PB_FLAGS(P) & FLAG_FLAT) a = a + fa_;
    if (PB_FLAGS(P) & FLAG_GR) a = a + ga;
    S3 wd = s3(zero, zero, zero);
#if !PB_FIXED_N
    if (PB_FLAGS(P) & FLAG_WIND) wd = wind_dangular_momentum_dt(P, ro, cold, b, sys, s, w2);
#endif
    S3 dl = s3(td.x + fd.x + gd.x + wd.x, td.y + fd.y + gd.y + wd.y, td.z + fd.z + gd.z + wd.z);
    if (!ro.valid) { a = s3(zero, zero, zero); dl = a; }
    a_out = plain(a);
    dl_out = plain(dl);
}

}  // namespace PB_NS
