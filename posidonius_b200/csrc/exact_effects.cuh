// exact_effects.cuh — the perturbation forces in the reference's own arithmetic (PB200_ARITH_STRICT, and the committed
// evaluation of PB200_ARITH_HYBRID).
//
// Same physics as additional_effects() in forces_fast.cuh, but every operation is an IEEE round-to-nearest
// add / mul / div / sqrt in the association order of the reference source (transcribed from the CPU oracle, which is
// bit-exact against the reference's golden vectors), sums over planets are accumulated in body order, and `powi` follows
// LLVM's square-and-multiply expansion. Together with the strict core this makes the evaluation bit-reproducible against
// the oracle (tests/test_gpu_parity.py::test_strict_mode_is_bit_identical..., test_hybrid_mode_...).
//
// What is NOT done the slow way (all of it leaves every rounding where the reference has it):
//   * a / b is the branch-free correctly rounded division of strict.cuh; divisions by the same divisor (the distance and
//     its powers, |v|, |w|^2, the inertia, M + m) share one reciprocal refinement: 3 instructions per further quotient;
//   * products of step-invariant operands that LEAD an expression in the reference's association order (the tidal
//     numerators, the polynomials in the GR factor, 1 / m, mass_factor (M + m), the reduced mass) come from make_consts;
//   * the host sums: the reference's serial loops fix the order of each SCALAR sum only, so lane c of the group accumulates
//     scalar c over the bodies in index order (host_sums6: two vectors per round, one round per effect, two for GR) —
//     7 dependent additions per round instead of 84 on the host lane;
//   * the acceleration and dL/dt are accumulated effect by effect in the reference's order (tides, flattening, GR, wind)
//     instead of being kept apart until the end: nothing but the running sums stays live across the effects.
// Included inside namespace PB_NS; no include guard on purpose.

namespace PB_NS {
using namespace pb200;

// Host sums of two vectors over the non-host bodies in index order — the same additions in the same order as the
// reference's serial loops over the particles. Every lane passes its terms (zero unless it is an OrbitingBody of the
// effect); on return the HOST lane holds the sums (the other lanes get unspecified finite values).
// Three warp barriers per round (before the terms are published, before the walk, before the totals are read): the cell
// that receives the total of scalar c is the host's own term cell of that scalar (the host is not part of the sum).
__device__ __forceinline__ void host_sums6(const KParams& P, const Cold& cold, int b, S3 u, S3 w, S3& su, S3& sw) {
#if PB_DIST
    // (the previous round's totals sit in the host's term cells, which the host rewrites now: every lane has read them first)
    __syncwarp();
    dist_put3(cold, M_0, 0, plain(u)); dist_put3(cold, M_0, 3, plain(w));
    __syncwarp();
    if (b < 6) {
        const unsigned row = dist_self(cold) + (unsigned)((M_0 + b) * PB_BLOCK * 8);
        dist_st(row, 0, dist_walk<false>(row, sd(0.)).v);
    }
    __syncwarp();
    su = strict(v3(cold.getk(0, M_0), cold.getk(1, M_1), cold.getk(2, M_2)));
    sw = strict(v3(cold.getk(3, M_3), cold.getk(4, M_4), cold.getk(5, M_5)));
#else
    const int W = PB_W(P), n = PB_N(P), host = PB_HOST(P);
    cold.set3(M_0, plain(u)); cold.set3(M_3, plain(w));
    __syncwarp();
    for (int c = b; c < 6; c += W) {
        sd acc = sd(0.);
        for (int k = 0; k < n; k++) {
            if (k == host) continue;
            acc = acc + sd(cold.getk(k, M_0 + c));
        }
        cold.grp[host + (M_0 + c) * PB_BLOCK] = acc.v;
    }
    __syncwarp();
    su = strict(cold.get3(M_0)); sw = strict(cold.get3(M_3));
#endif
}

template <int GR>
__device__ __forceinline__ void additional_effects_exact(const KParams& P, const Roles& ro, const Cold& cold, int hl, int b, size_t sys,
                                                         double t, bool evolve_now, Lane& q, S3 hr, sd dist, S3 hv, V3& a_out,
                                                         V3& dl_out, bool tide_save) {
    const int host = PB_HOST(P);
    const sd zero = sd(0.);
    const S3 zero3 = s3(zero, zero, zero);
    // Q3: r.omega with the spins of the previous evaluation (tides/common.rs:155-160 = rotational_flattening/common.rs:105-110),
    // carried from it (every mode leaves them rounded like the reference when an exact evaluation can follow)
    const sd rs_s = sd(q.rs_s), rs_p = sd(q.rs_p);
    // calculate_spin (particles/common.rs:3-15)
    const sd I = sd(cold.get(K_I));
    const S3 s = strict(q.L) / make_rcp(I);
    const sd w2 = (s.x * s.x) + (s.y * s.y) + (s.z * s.z);
    q.s = plain(s);
    // (no barrier before these stores: the group's last reads of E_S / M_6 lie before the previous evaluation's exchange barriers)
    cold.set3(E_S, q.s); cold.set(M_6, w2.v);
    __syncwarp();
    const S3 sh = strict(cold.getk3(host, E_S));
    const sd wh2 = sd(cold.getk(host, M_6));
    q.rs_s = sdot(hr, sh).v; q.rs_p = sdot(hr, s).v;
#if !PB_FIXED_N
    if (evolve_now && (PB_FLAGS(P) & FLAG_DYN) && (PB_FLAGS(P) & FLAG_EVO)) { update_lag_angle(P, ro, b, sys, t, w2, true); __syncwarp(); }
#endif
    // inertial_to_heliocentric (universe.rs:331-338); the host's stale heliocentric velocity is zero (validated)
    const srcp rD = make_rcp(dist);
    const sd radvel = (hr.x * hv.x + hr.y * hv.y + hr.z * hv.z) / rD;
    const sd normv2 = hv.x * hv.x + hv.y * hv.y + hv.z * hv.z;
    const sd d2 = dist * dist, d4 = d2 * d2;
    const sd d5 = dist * d4, d7 = (dist * d2) * d4;   // powi as LLVM expands it
    const srcp rD7 = make_rcp(d7);
    const sd m = sd(cold.get(K_M)), M = sd(cold.getk(host, K_M));
    const sd inv_m = sd(cold.get(C_INVM)), inv_M = sd(cold.getk(host, C_INVM));
    const sd neg_inv_M = sd(-1.0) * inv_M;   // -1.0 * factor2
    S3 a = zero3;                            // add_additional_acceleration_corrections (universe.rs:540-567): tides, flattening, GR
    S3 td = zero3, fd = zero3, gd = zero3;   // calculate_dangular_momentum_dt (universe.rs:580-614)
    if (PB_FLAGS(P) & FLAG_TIDES) {
        const sd m2 = m * m, M2 = M * M;
        const sd Rh10 = sd(cold.getk(host, K_R10)), R10 = sd(cold.get(K_R10));
        sd sig_h = sd(cold.getk(host, K_SIG)), sig_p = sd(cold.get(K_SIG));
        const double2 num = cold.get2(C_AS, 0);
        sd cs = sd(num.x), cp = sd(num.y);       // constant_time_lag.rs:232-234, 243-245 numerators (zero when not central)
#if !PB_FIXED_N
        if (PB_FLAGS(P) & FLAG_DYN) {
            // sigma is the last factor of each product in the reference, so multiplying it in here rounds identically
            pair_dependent_sigmas(P, ro, cold, hl, b, sys, hr, hv, w2, wh2, sig_h, sig_p);
            cs = P.tides_host_central ? sd(4.5) * m2 * Rh10 * sig_h : zero;
            cp = sd(4.5) * M2 * R10 * sig_p;
        }
#endif
        const sd t1 = m2 * Rh10 * sig_h;                 // :291-293
        const sd t2 = M2 * R10 * sig_p;                  // :294-296
        const sd host_k = m2 * sd(cold.getk(host, K_R5)) * sd(cold.getk(host, K_K2T)) + M2 * sd(cold.get(K_R5)) * sd(cold.get(K_K2T));   // :284-285
        const sd d8 = d4 * d4;
        const sd orth_s = cs / rD7;
        const sd orth_p = cp / rD7;
        const sd cons = sd(-3.0 * kK2) / rD7 * host_k;
        const sd factor1 = sd(-13.5) * radvel / d8;
        const sd diss_pm = factor1 * t2;
        const sd diss = diss_pm + factor1 * t1;
        const sd t_radial = cons + diss;
        const sd f3 = t_radial + (orth_s + orth_p) * radvel / rD;
        const sd osd = orth_s / rD, opd = orth_p / rD;
        S3 F;
        F.x = f3 * hr.x / rD + osd * (sh.y * hr.z - sh.z * hr.y - hv.x) + opd * (s.y * hr.z - s.z * hr.y - hv.x);
        F.y = f3 * hr.y / rD + osd * (sh.z * hr.x - sh.x * hr.z - hv.y) + opd * (s.z * hr.x - s.x * hr.z - hv.y);
        F.z = f3 * hr.z / rD + osd * (sh.x * hr.y - sh.y * hr.x - hv.z) + opd * (s.x * hr.y - s.y * hr.x - hv.z);
        const sd oned = sd(1.0) / rD;
        const sd cx = hr.y * hv.z - hr.z * hv.y, cy = hr.z * hv.x - hr.x * hv.z, cz = hr.x * hv.y - hr.y * hv.x;
        S3 Np, Ns;
        Np.x = orth_p * (dist * s.x - rs_p * hr.x / rD - oned * cx);
        Np.y = orth_p * (dist * s.y - rs_p * hr.y / rD - oned * cy);
        Np.z = orth_p * (dist * s.z - rs_p * hr.z / rD - oned * cz);
        Ns.x = orth_s * (dist * sh.x - rs_s * hr.x / rD - oned * cx);
        Ns.y = orth_s * (dist * sh.y - rs_s * hr.y / rD - oned * cy);
        Ns.z = orth_s * (dist * sh.z - rs_s * hr.z / rD - oned * cz);
        S3 t_acc = zero3, t_dl = zero3, xF = zero3, xN = zero3;
        if (ro.t_on) {
            t_acc = s3(inv_m * F.x, inv_m * F.y, inv_m * F.z);
            t_dl = s3(sd(-1.0) * Np.x, sd(-1.0) * Np.y, sd(-1.0) * Np.z);
            xF = F;
            xN = s3(sd(-1.0) * Ns.x, sd(-1.0) * Ns.y, sd(-1.0) * Ns.z);
        }
        if (tide_save && ro.valid) {
            // internals that calculate_denergy_dt (tides/common.rs:263-279) will read at the next snapshot
            const size_t ns = (size_t)P.n_sys;
            double* ts = P.tide_scratch + (size_t)b * ns + sys;
            const size_t cs_ = (size_t)PB_N(P) * ns;
            ts[0 * cs_] = hr.x.v; ts[1 * cs_] = hr.y.v; ts[2 * cs_] = hr.z.v;
            ts[3 * cs_] = hv.x.v; ts[4 * cs_] = hv.y.v; ts[5 * cs_] = hv.z.v;
            ts[6 * cs_] = dist.v; ts[7 * cs_] = radvel.v; ts[8 * cs_] = orth_p.v; ts[9 * cs_] = diss_pm.v;
            ts[10 * cs_] = t_dl.x.v; ts[11 * cs_] = t_dl.y.v; ts[12 * cs_] = t_dl.z.v;
        }
        S3 sF, sN;
        host_sums6(P, cold, b, xF, xN, sF, sN);
        if (ro.host) { t_acc = s3(neg_inv_M * sF.x, neg_inv_M * sF.y, neg_inv_M * sF.z); t_dl = sN; }
        a = a + t_acc;
        td = t_dl;
    }
    if (PB_FLAGS(P) & FLAG_FLAT) {
        // oblate_spheroid.rs:12-97, rotational_flattening/common.rs:165-237
        const sd Rh5 = sd(cold.getk(host, K_R5)), R5 = sd(cold.get(K_R5));
        const sd fs0 = P.flat_host_central ? m * sd(cold.getk(host, K_K2F)) : zero;   // oblate_spheroid.rs:37 leading product
        const sd fp0 = M * sd(cold.get(K_K2F));                                        // :42
        const srcp r6 = make_rcp(sd(6.));
        const sd ffs = fs0 * wh2 * Rh5 / r6;
        const sd orth_s = sd(-6.) * ffs * rs_s / (wh2 * d5);
        const sd ffp = fp0 * w2 * R5 / r6;
        const sd orth_p = sd(-6.) * ffp * rs_p / (w2 * d5);
        const sd radial = sd(-3.) / d5 * (ffp + ffs) + sd(15.) / rD7 * (ffs * rs_s * rs_s / wh2 + ffp * rs_p * rs_p / w2);
        S3 F;
        F.x = radial * hr.x + orth_p * s.x + orth_s * sh.x;
        F.y = radial * hr.y + orth_p * s.y + orth_s * sh.y;
        F.z = radial * hr.z + orth_p * s.z + orth_s * sh.z;
        S3 Np, Ns;
        Np.x = orth_p * (hr.y * s.z - hr.z * s.y); Np.y = orth_p * (hr.z * s.x - hr.x * s.z); Np.z = orth_p * (hr.x * s.y - hr.y * s.x);
        Ns.x = orth_s * (hr.y * sh.z - hr.z * sh.y); Ns.y = orth_s * (hr.z * sh.x - hr.x * sh.z); Ns.z = orth_s * (hr.x * sh.y - hr.y * sh.x);
        S3 f_acc = zero3, f_dl = zero3, xF = zero3, xN = zero3;
        if (ro.f_on) {
            f_acc = s3(inv_m * F.x, inv_m * F.y, inv_m * F.z);
            f_dl = s3(sd(-1.0) * Np.x, sd(-1.0) * Np.y, sd(-1.0) * Np.z);
            xF = F;
            xN = s3(sd(-1.0) * Ns.x, sd(-1.0) * Ns.y, sd(-1.0) * Ns.z);
        }
        S3 sF, sN;
        host_sums6(P, cold, b, xF, xN, sF, sN);
        if (ro.host) { f_acc = s3(neg_inv_M * sF.x, neg_inv_M * sF.y, neg_inv_M * sF.z); f_dl = sN; }
        a = a + f_acc;
        fd = f_dl;
    }
    if (GR == PB200_GR_KIDDER1995) {
        // general_relativity.rs:177-456, transcribed from the oracle (oracle_core.hpp gr_kidder)
        const sd c2 = sd(kC2);
        const sd mgs = sd(cold.get2(C_AS, 5).y);   // mu_host + mu (zero on lanes that are not OrbitingBody: their terms are dropped below)
        const double2 g01 = cold.get2(G_0, 0), g23 = cold.get2(G_0, 1), g45 = cold.get2(G_0, 2), g67 = cold.get2(G_0, 3), g89 = cold.get2(G_0, 4), gab = cold.get2(G_0, 5);
        const sd normv = ssqrt(normv2);
        const srcp rV = make_rcp(normv);
        const sd rv2 = radvel * radvel;
        const srcp rD2c2 = make_rcp(d2 * c2);
        const sd pre = -mgs / rD2c2;
        const sd mgd = mgs / rD;
        // 1PN (:187-239)
        const sd radial1 = pre * (sd(g01.x) * normv2 - sd(g01.y) * mgs / rD - sd(g23.x) * rv2);
        const sd orth1 = mgs / rD2c2 * sd(g23.y) * radvel * normv;
        S3 a1;
        a1.x = radial1 * hr.x / rD + orth1 * hv.x / rV;
        a1.y = radial1 * hr.y / rD + orth1 * hv.y / rV;
        a1.z = radial1 * hr.z / rD + orth1 * hv.z / rV;
        // 2PN (:241-298)
        const sd v4 = normv2 * normv2, rv4 = rv2 * rv2;
        const sd radial2 = pre
            * (sd(g45.x) * (mgs * mgs / d2)
               + sd(g45.y) * v4
               + sd(g67.x) * rv4
               - sd(g67.y) * rv2 * normv2
               - sd(g89.x) * mgd * normv2
               - sd(g89.y) * mgd * rv2);
        const sd orth2 = pre * sd(-0.5) * radvel * (sd(gab.x) * normv2 - sd(gab.y) * mgd - sd(cold.get(G_0 + 12)) * rv2);
        S3 a2;
        a2.x = radial2 * hr.x / rD + orth2 * hv.x;
        a2.y = radial2 * hr.y / rD + orth2 * hv.y;
        a2.z = radial2 * hr.z / rD + orth2 * hv.z;
        // 1.5PN spin-orbit (:300-456)
        const sd Ih = sd(cold.getk(host, K_I));
        const srcp rM = srcp{M.v, cold.getk(host, K_WHDSF)};   // the host mass with its refined reciprocal (make_rcp(M), set when the CTA starts)
        const srcp rm = make_rcp(m);
        const S3 Ls = s3(Ih * sh.x, Ih * sh.y, Ih * sh.z), Lp = s3(I * s.x, I * s.y, I * s.z);
        const S3 nn = hr / rD;
        const sd mfm = sd(cold.get(Z_MFM));
        const sd msx = mfm * (Lp.x / rm - Ls.x / rM), msy = mfm * (Lp.y / rm - Ls.y / rM), msz = mfm * (Lp.z / rm - Ls.z / rM);
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
        const S3 a3 = s3(fa * (e1x - e2x + e3x), fa * (e1y - e2y + e3y), fa * (e1z - e2z + e3z));
        const sd mom = m / rM;                                  // particle.mass / host.mass (:216)
        S3 g_acc = zero3, x1 = zero3, x2 = zero3;
        if (ro.g_on) {
            g_acc = s3(a1.x + a2.x + a3.x, a1.y + a2.y + a3.y, a1.z + a2.z + a3.z);
            x1 = s3(mom * a1.x, mom * a1.y, mom * a1.z);
            x2 = s3(mom * a2.x, mom * a2.y, mom * a2.z);
        }
        // first round: the host's 1PN and 2PN sums
        S3 s1, s2;
        host_sums6(P, cold, b, x1, x2, s1, s2);
        // Kidder 1995 eq. 2.4a / 2.4b
        const sd mu = sd(cold.get(Z_MURED));
        const S3 Lo = s3(mu * (hr.y * hv.z - hr.z * hv.y), mu * (hr.z * hv.x - hr.x * hv.z), mu * (hr.x * hv.y - hr.y * hv.x));
        const sd fms = sd(2.) + sd(1.5) * m / rM;               // :390   2 + 3/2 m / M
        const sd fmp = sd(2.) + sd(1.5) * M / rm;               // :419   2 + 3/2 M / m
        const sd a1x = fms * (Lo.y * Ls.z - Lo.z * Ls.y), a1y = fms * (Lo.z * Ls.x - Lo.x * Ls.z), a1z = fms * (Lo.x * Ls.y - Lo.y * Ls.x);
        const sd a2x = Lp.y * Ls.z - Lp.z * Ls.y, a2y = Lp.z * Ls.x - Lp.x * Ls.z, a2z = Lp.x * Ls.y - Lp.y * Ls.x;
        const sd spp = nn.x * Lp.x + nn.y * Lp.y + nn.z * Lp.z;
        const sd a3x = sd(3.) * spp * (nn.y * Ls.z - nn.z * Ls.y), a3y = sd(3.) * spp * (nn.z * Ls.x - nn.x * Ls.z), a3z = sd(3.) * spp * (nn.x * Ls.y - nn.y * Ls.x);
        const S3 hdl = s3(fa * (a1x - a2x + a3x), fa * (a1y - a2y + a3y), fa * (a1z - a2z + a3z));
        const sd b1x = fmp * (Lo.y * Lp.z - Lo.z * Lp.y), b1y = fmp * (Lo.z * Lp.x - Lo.x * Lp.z), b1z = fmp * (Lo.x * Lp.y - Lo.y * Lp.x);
        const sd b2x = Ls.y * Lp.z - Ls.z * Lp.y, b2y = Ls.z * Lp.x - Ls.x * Lp.z, b2z = Ls.x * Lp.y - Ls.y * Lp.x;
        const sd ssp = nn.x * Ls.x + nn.y * Ls.y + nn.z * Ls.z;
        const sd b3x = sd(3.) * ssp * (nn.y * Lp.z - nn.z * Lp.y), b3y = sd(3.) * ssp * (nn.z * Lp.x - nn.x * Lp.z), b3z = sd(3.) * ssp * (nn.x * Lp.y - nn.y * Lp.x);
        S3 g_dl = zero3, x3 = zero3, x4 = zero3;
        if (ro.g_on) {
            g_dl = s3(fa * (b1x - b2x + b3x), fa * (b1y - b2y + b3y), fa * (b1z - b2z + b3z));
            x3 = s3(mom * a3.x, mom * a3.y, mom * a3.z);
            x4 = hdl;
        }
        // second round: the host's 1.5PN sum and its torque
        S3 s3_, s4;
        host_sums6(P, cold, b, x3, x4, s3_, s4);
        if (ro.host) {
            const sd m1 = sd(-1.0);
            g_acc = s3(m1 * s1.x + m1 * s2.x + m1 * s3_.x, m1 * s1.y + m1 * s2.y + m1 * s3_.y, m1 * s1.z + m1 * s2.z + m1 * s3_.z);
            g_dl = s4;
        }
        if (PB_FLAGS(P) & FLAG_GR) a = a + g_acc;
        gd = g_dl;
    }
    S3 wd = zero3;
#if !PB_FIXED_N
    if (PB_FLAGS(P) & FLAG_WIND) wd = wind_dangular_momentum_dt(P, ro, cold, b, sys, s, w2);
#endif
    S3 dl = s3(td.x + fd.x + gd.x + wd.x, td.y + fd.y + gd.y + wd.y, td.z + fd.z + gd.z + wd.z);
    if (!ro.valid) { a = zero3; dl = zero3; }
    a_out = plain(a);
    dl_out = plain(dl);
}

}  // namespace PB_NS
