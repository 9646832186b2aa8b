// strict_gr_variants.cuh — PB200_ARITH_STRICT versions of the Anderson1975 and Newhall1983 general relativity variants
// (effects/general_relativity.rs:461-895): every operation an IEEE round-to-nearest add/mul/div/sqrt in the association
// order of the reference source, transcribed from the CPU oracle (oracle_core.hpp: gr_newtonian, gr_anderson, gr_newhall),
// sums over bodies in the reference's loop order. The lanes of a group exchange through the columns X_0 .. X_8 (the slots of
// the Kidder1995 polynomials, which these variants never read; cold_slots.cuh); inertial positions are in S_RX while the
// midpoint runs. Included inside namespace PB_NS after exact_effects.cuh; no include guard on purpose.

namespace PB_NS {
using namespace pb200;

__device__ __forceinline__ S3 getk3s(const Cold& cold, int k, int slot) { return strict(cold.getk3(k, slot)); }

// Newtonian inertial accelerations with the terms WHFast ignored re-added (general_relativity.rs:641-678).
// jacobi_coords: IgnoreGravityTerms::WHFastOne (only the first non-host particle is re-added), else WHFastTwo.
__device__ __forceinline__ S3 gr_newtonian_strict(const KParams& P, const Roles& ro, const Cold& cold, int b, S3 hr, S3 acc_newton,
                                                  bool jacobi_coords) {
    const int n = PB_N(P), host = PB_HOST(P);
    const int first_other = host == 0 ? 1 : 0;
    const bool included = ro.planet && (!jacobi_coords || b == first_other);
    const sd m = sd(cold.get(K_M)), M = sd(cold.getk(host, K_M));
    const S3 rh = getk3s(cold, host, S_RX);
    // Q9: host INERTIAL position minus the particle's HELIOCENTRIC position
    const S3 dx = ro.planet ? rh - hr : s3(sd(1.), sd(0.), sd(0.));
    const sd r2 = dx.x * dx.x + dx.y * dx.y + dx.z * dx.z;
    const sd r = ssqrt(r2);
    const sd prefac = sd(kG) / (r2 * r);
    const sd pms = prefac * M, pmp = prefac * m;
    __syncwarp();
    cold.set3(X_0, plain(s3(pmp * dx.x, pmp * dx.y, pmp * dx.z)));
    __syncwarp();
    S3 a = acc_newton;
    if (included) a = s3(a.x + pms * dx.x, a.y + pms * dx.y, a.z + pms * dx.z);
    if (ro.host) {
        for (int k = 0; k < n; k++) {
            if (k == host) continue;
            if (jacobi_coords && k != first_other) continue;
            const S3 t = getk3s(cold, k, X_0);
            a = s3(a.x - t.x, a.y - t.y, a.z - t.z);
        }
    }
    return a;
}

// general_relativity.rs:461-636
__device__ __forceinline__ void gr_anderson1975_strict(const KParams& P, const Roles& ro, const Cold& cold, int b, size_t sys, const Lane& q, S3 hr,
                                                       S3 acc_newton, bool jacobi_coords, V3& a_out) {
    const int n = PB_N(P), host = PB_HOST(P);
    const S3 an = gr_newtonian_strict(P, ro, cold, b, hr, acc_newton, jacobi_coords);
    const sd c2 = sd(kC2), one = sd(1.);
    __syncwarp();
    cold.set3(X_0, plain(an)); cold.set3(X_3, plain(q.v));
    __syncwarp();
    // inertial -> Jacobi over the OrbitingBody particles (:539-602); every lane carries the running sums
    sd eta = sd(cold.getk(host, K_M));
    const S3 r_h = getk3s(cold, host, S_RX), v_h = getk3s(cold, host, X_3), a_h = getk3s(cold, host, X_0);
    S3 s = s3(eta * r_h.x, eta * r_h.y, eta * r_h.z), sv = s3(eta * v_h.x, eta * v_h.y, eta * v_h.z), sa = s3(eta * a_h.x, eta * a_h.y, eta * a_h.z);
    S3 jp = s3(one, sd(0.), sd(0.)), jv = s3(sd(0.), one, sd(0.)), ja = s3(sd(0.), sd(0.), sd(0.));   // benign for idle lanes
    for (int k = 0; k < n; k++) {
        if (k == host || !((P.gr_orbiting >> k) & 1u)) continue;
        const sd mk = sd(cold.getk(k, K_M));
        const S3 rk = getk3s(cold, k, S_RX), vk = getk3s(cold, k, X_3), ak = getk3s(cold, k, X_0);
        const sd ei = one / eta;
        eta = eta + mk;
        const sd pme = eta * ei;
        const S3 pk = s3(rk.x - s.x * ei, rk.y - s.y * ei, rk.z - s.z * ei);
        const S3 wk = s3(vk.x - sv.x * ei, vk.y - sv.y * ei, vk.z - sv.z * ei);
        const S3 ck = s3(ak.x - sa.x * ei, ak.y - sa.y * ei, ak.z - sa.z * ei);
        if (b == k && ro.g_on) { jp = pk; jv = wk; ja = ck; }
        s = s3(s.x * pme + mk * pk.x, s.y * pme + mk * pk.y, s.z * pme + mk * pk.z);
        sv = s3(sv.x * pme + mk * wk.x, sv.y * pme + mk * wk.y, sv.z * pme + mk * wk.z);
        sa = s3(sa.x * pme + mk * ck.x, sa.y * pme + mk * ck.y, sa.z * pme + mk * ck.z);
    }
    const sd jacobi_star_mass = eta;
    const sd mu = sd(P.mass_g[(size_t)host * (size_t)P.n_sys + (ro.valid ? sys : 0)]);
    // fixed point on the velocity (:478-516), this lane's body
    {
        S3 vi = jv;
        sd vi2 = jv.x * jv.x + jv.y * jv.y + jv.z * jv.z;
        const sd ri = ssqrt(jp.x * jp.x + jp.y * jp.y + jp.z * jp.z);
        sd fa = (sd(0.5) * vi2 + sd(3.) * mu / ri) / c2;
        bool lane_done = false;
        for (int it = 0; it < 10; it++) {
            if (!__any_sync(FULL, !lane_done)) break;
            const S3 old = vi;
            const srcp d = make_rcp(one - fa);
            const S3 vn = s3(jv.x / d, jv.y / d, jv.z / d);
            const sd vn2 = vn.x * vn.x + vn.y * vn.y + vn.z * vn.z;
            const sd fan = (sd(0.5) * vn2 + sd(3.) * mu / ri) / c2;
            const sd dvx = vn.x - old.x, dvy = vn.y - old.y, dvz = vn.z - old.z;
            if (!lane_done) {
                vi = vn; vi2 = vn2; fa = fan;
                if (((dvx * dvx + dvy * dvy + dvz * dvz) / vi2).v < kEps2) lane_done = true;
            }
        }
        const sd ri3 = ri * ri * ri;
        const sd fb = (mu / ri - sd(1.5) * vi2) * mu / ri3 / c2;
        const sd rdotrdot = jp.x * jv.x + jp.y * jv.y + jp.z * jv.z;
        const S3 vidot = s3(ja.x + fb * jp.x, ja.y + fb * jp.y, ja.z + fb * jp.z);
        const sd vdotvdot = vi.x * vidot.x + vi.y * vidot.y + vi.z * vidot.z;
        const sd fd = (vdotvdot - sd(3.) * mu / ri3 * rdotrdot) / c2;
        const sd omfa = one - fa;
        ja = s3(fb * omfa * jp.x - fa * ja.x - fd * vi.x, fb * omfa * jp.y - fa * ja.y - fd * vi.y, fb * omfa * jp.z - fa * ja.z - fd * vi.z);
    }
    // Jacobi -> inertial accelerations (:604-636); the star's Jacobi acceleration is zero
    __syncwarp();
    cold.set3(X_6, plain(ja));
    __syncwarp();
    eta = jacobi_star_mass;
    S3 sacc = s3(eta * sd(0.), eta * sd(0.), eta * sd(0.));
    S3 mine = s3(sd(0.), sd(0.), sd(0.));
    for (int k = n - 1; k >= 0; k--) {
        if (k == host || !((P.gr_orbiting >> k) & 1u)) continue;
        const sd mk = sd(cold.getk(k, K_M));
        const S3 jk = getk3s(cold, k, X_6);
        const sd ei = one / eta;
        sacc = s3((sacc.x - mk * jk.x) * ei, (sacc.y - mk * jk.y) * ei, (sacc.z - mk * jk.z) * ei);
        if (b == k) mine = s3(jk.x + sacc.x, jk.y + sacc.y, jk.z + sacc.z);
        eta = eta - mk;
        sacc = s3(sacc.x * eta, sacc.y * eta, sacc.z * eta);
    }
    const sd mtot_i = one / eta;
    const S3 star = s3(sacc.x * mtot_i, sacc.y * mtot_i, sacc.z * mtot_i);
    a_out = ro.host ? plain(star) : (ro.g_on ? plain(mine) : v3(0., 0., 0.));
}

// general_relativity.rs:683-895. Body order of every loop: the host first, then the others in list order.
__device__ __forceinline__ void gr_newhall1983_strict(const KParams& P, const Roles& ro, const Cold& cold, int b, const Lane& q, S3 hr,
                                                      S3 acc_newton, bool jacobi_coords, V3& a_out) {
    const int n = PB_N(P), host = PB_HOST(P);
    const S3 an = gr_newtonian_strict(P, ro, cold, b, hr, acc_newton, jacobi_coords);
    const sd c2 = sd(kC2), G = sd(kG);
    const bool en_i = (P.gr_enabled >> b) & 1u;
    const S3 qr = ro.valid ? strict(cold.get3(S_RX)) : s3(sd(1. + b), sd(0.), sd(0.));
    const S3 qv = q.v;
    __syncwarp();
    cold.set3(X_0, plain(an)); cold.set3(X_3, plain(qv));
    __syncwarp();
    // a1 of this body and the a2 sum that every pair (., this body) uses: the same terms with 4/c^2 and 1/c^2 (:745-760)
    sd a1 = sd(0.), a2_self = sd(0.);
    for (int kk = -1; kk < n; kk++) {
        const int k = kk < 0 ? host : kk;
        if (kk == host || k == b) continue;
        const S3 rk = getk3s(cold, k, S_RX);
        const S3 dr = s3(qr.x - rk.x, qr.y - rk.y, qr.z - rk.z);
        const sd rs = ssqrt(dr.x * dr.x + dr.y * dr.y + dr.z * dr.z);
        const sd mk = sd(cold.getk(k, K_M));
        a1 = a1 + sd(4. / kC2) * G * mk / rs;
        a2_self = a2_self + sd(1. / kC2) * G * mk / rs;
    }
    cold.set(X_6, a2_self.v);
    __syncwarp();
    const sd vi2 = qv.x * qv.x + qv.y * qv.y + qv.z * qv.z;
    S3 ac = s3(sd(0.), sd(0.), sd(0.)), nc = ac;
    for (int kk = -1; kk < n; kk++) {
        const int j = kk < 0 ? host : kk;
        if (kk == host || j == b) continue;
        const bool en_j = (P.gr_enabled >> j) & 1u;
        if (!(en_i || en_j) || !ro.valid) continue;
        const S3 rj = getk3s(cold, j, S_RX), vj = getk3s(cold, j, X_3), tj = getk3s(cold, j, X_0);
        const sd mj = sd(cold.getk(j, K_M));
        const sd a2 = sd(cold.getk(j, X_6));
        const S3 dr = s3(qr.x - rj.x, qr.y - rj.y, qr.z - rj.z);
        const sd rij = ssqrt(dr.x * dr.x + dr.y * dr.y + dr.z * dr.z);
        const sd rij2 = rij * rij;
        const sd rij3 = rij2 * rij;
        const sd a3 = -vi2 / c2;
        const sd vj2 = vj.x * vj.x + vj.y * vj.y + vj.z * vj.z;
        const sd a4 = sd(-2.) * vj2 / c2;
        const sd a5 = sd(4. / kC2) * (qv.x * vj.x + qv.y * vj.y + qv.z * vj.z);
        const sd a60 = dr.x * vj.x + dr.y * vj.y + dr.z * vj.z;
        const sd a6 = sd(3. / (2. * kC2)) * (a60 * a60) / rij2;
        const sd factor1 = a1 + a2 + a3 + a4 + a5 + a6;
        const srcp r3 = make_rcp(rij3);
        ac = s3(ac.x + G * mj * dr.x * factor1 / r3, ac.y + G * mj * dr.y * factor1 / r3, ac.z + G * mj * dr.z * factor1 / r3);
        const sd dvx = qv.x - vj.x, dvy = qv.y - vj.y, dvz = qv.z - vj.z;
        const sd factor2 = dr.x * (sd(4.) * qv.x - sd(3.) * vj.x) + dr.y * (sd(4.) * qv.y - sd(3.) * vj.y) + dr.z * (sd(4.) * qv.z - sd(3.) * vj.z);
        ac = s3(ac.x + G * mj * factor2 * dvx / r3 / c2, ac.y + G * mj * factor2 * dvy / r3 / c2, ac.z + G * mj * factor2 * dvz / r3 / c2);
        // the substitution pass (:806-851) with a_old = 0 (Q8: the inverted deviation test ends the loop after one pass)
        const sd tx = tj.x + sd(0.), ty = tj.y + sd(0.), tz = tj.z + sd(0.);
        const sd proj = dr.x * tx + dr.y * ty + dr.z * tz;
        const sd twoc2 = sd(2. * kC2), k7 = sd(7. / (2. * kC2));
        nc = s3(nc.x + (G * mj * dr.x / r3) * proj / twoc2 + k7 * G * mj * tx / rij,
                nc.y + (G * mj * dr.y / r3) * proj / twoc2 + k7 * G * mj * ty / rij,
                nc.z + (G * mj * dr.z / r3) * proj / twoc2 + k7 * G * mj * tz / rij);
    }
    const S3 a_new = s3(ac.x + nc.x, ac.y + nc.y, ac.z + nc.z);
    a_out = (ro.host || ro.g_on) ? plain(a_new) : v3(0., 0., 0.);
}

}  // namespace PB_NS
