// gr_variants.cuh — Anderson1975 and Newhall1983 general relativity (REBOUNDx-derived variants of the
// reference, effects/general_relativity.rs:461-895). Lane = body; the serial Jacobi recurrences are
// carried redundantly by every lane of the group, the all-pairs sums of Newhall walk the group with
// shuffles. These variants are selected by one fixture each upstream; they are correct, not tuned.
#include "forces_fast.cuh"

namespace PB_NS {
using namespace pb200;

// Newtonian inertial accelerations with the terms WHFast ignored re-added (general_relativity.rs:641-678).
// jacobi_coords: IgnoreGravityTerms::WHFastOne (only the first non-host particle is re-added), else WHFastTwo.
__device__ __forceinline__ V3 gr_newtonian(const KParams& P, const Roles& ro, const Cold& cold, int gb, int hl, int b, const Lane& q, V3 hr,
                                           V3 acc_newton, bool jacobi_coords) {
    const double q_m = cold.get(K_M);
    const int first_other = PB_HOST(P) == 0 ? 1 : 0;
    bool included = ro.planet && (!jacobi_coords || b == first_other);
    V3 rh = shfl3(plain(q.r), hl);
    double M = shfl(q_m, hl);
    // Q9: host INERTIAL position minus the particle's HELIOCENTRIC position
    V3 dx = rh - hr;
    double r2 = dot(dx, dx);
    double r = sqrt(r2);
    double prefac = kG / (r2 * r);
    V3 to_host = included ? (-(prefac * q_m)) * dx : v3(0., 0., 0.);
    V3 own = included ? (prefac * M) * dx : v3(0., 0., 0.);
    // ordered sum over the non-host bodies, as the reference accumulates
    V3 hsum = v3(0., 0., 0.);
    for (int k = 0; k < PB_N(P); k++) {
        if (k == PB_HOST(P)) continue;
        V3 t = shfl3(to_host, gb + k);
        hsum = hsum + t;
    }
    V3 a = acc_newton + own;
    if (ro.host) a = acc_newton + hsum;
    return a;
}

// general_relativity.rs:461-636
__device__ __forceinline__ void gr_anderson1975(const KParams& P, const Roles& ro, const Cold& cold, int gb, int hl, int b, size_t sys, const Lane& q, V3 hr,
                                                V3 acc_newton, bool jacobi_coords, V3& a_out) {
    const double q_m = cold.get(K_M), q_mg = ro.valid ? P.mass_g[(size_t)b * (size_t)P.n_sys + sys] : 1.;
    V3 an = gr_newtonian(P, ro, cold, gb, hl, b, q, hr, acc_newton, jacobi_coords);
    // inertial -> Jacobi over the OrbitingBody particles (:539-602); every lane carries the running sums
    double eta = shfl(q_m, hl);
    V3 s = eta * shfl3(plain(q.r), hl), sv = eta * shfl3(plain(q.v), hl), sa = eta * shfl3(an, hl);
    V3 jp = v3(0., 0., 0.), jv = v3(0., 0., 0.), ja = v3(0., 0., 0.);
    for (int k = 0; k < PB_N(P); k++) {
        if (k == PB_HOST(P) || !((P.gr_orbiting >> k) & 1u)) continue;
        double mk = shfl(q_m, gb + k);
        V3 rk = shfl3(plain(q.r), gb + k), vk = shfl3(plain(q.v), gb + k), ak = shfl3(an, gb + k);
        double ei = 1. / eta;
        eta += mk;
        double pme = eta * ei;
        V3 pk = rk - ei * s, wk = vk - ei * sv, ck = ak - ei * sa;
        if (b == k) { jp = pk; jv = wk; ja = ck; }
        s = pme * s + mk * pk; sv = pme * sv + mk * wk; sa = pme * sa + mk * ck;
    }
    const double jacobi_star_mass = eta;
    const double mu = shfl(q_mg, hl);
    // fixed point on the velocity (:478-516)
    {
        V3 vi = jv;
        double vi2 = dot(jv, jv);
        double ri = sqrt(dot(jp, jp));
        double fa = (0.5 * vi2 + 3. * mu / ri) * kInvC2;
        bool lane_done = !ro.g_on;
        for (int it = 0; it < 10; it++) {
            if (!__any_sync(FULL, !lane_done)) break;
            double inv = 1. / (1. - fa);
            V3 vn = inv * jv;
            double vn2 = dot(vn, vn);
            double fan = (0.5 * vn2 + 3. * mu / ri) * kInvC2;
            V3 dv = vn - vi;
            if (!lane_done) {
                vi = vn; vi2 = vn2; fa = fan;
                if (dot(dv, dv) / vi2 < kEps2) lane_done = true;
            }
        }
        double ri3 = ri * ri * ri;
        double fb = (mu / ri - 1.5 * vi2) * mu / ri3 * kInvC2;
        double rdotrdot = dot(jp, jv);
        V3 vidot = ja + fb * jp;
        double vdotvdot = dot(vi, vidot);
        double fd = (vdotvdot - 3. * mu / ri3 * rdotrdot) * kInvC2;
        ja = (fb * (1. - fa)) * jp - fa * ja - fd * vi;
    }
    // Jacobi -> inertial accelerations (:604-636); the star's Jacobi acceleration is zero
    eta = jacobi_star_mass;
    V3 sacc = v3(0., 0., 0.);
    V3 mine = v3(0., 0., 0.);
    for (int k = PB_N(P) - 1; k >= 0; k--) {
        if (k == PB_HOST(P) || !((P.gr_orbiting >> k) & 1u)) continue;
        double mk = shfl(q_m, gb + k);
        V3 jk = shfl3(ja, gb + k);
        double ei = 1. / eta;
        sacc = ei * (sacc - mk * jk);
        if (b == k) mine = jk + sacc;
        eta -= mk;
        sacc = eta * sacc;
    }
    V3 star = (1. / eta) * sacc;
    a_out = ro.host ? star : (ro.g_on ? mine : v3(0., 0., 0.));
}

// general_relativity.rs:683-895
__device__ __forceinline__ void gr_newhall1983(const KParams& P, const Roles& ro, const Cold& cold, int gb, int hl, int b, const Lane& q, V3 hr,
                                               V3 acc_newton, bool jacobi_coords, V3& a_out) {
    const double q_m = cold.get(K_M);
    V3 an = gr_newtonian(P, ro, cold, gb, hl, b, q, hr, acc_newton, jacobi_coords);
    const int n = PB_N(P);
    const bool en_i = (P.gr_enabled >> b) & 1u;
    const V3 qr = plain(q.r), qv = plain(q.v);
    // potential-like sums: pot_i = sum_{k != i} G m_k / r_ik
    double pot = 0.;
    for (int kk = -1; kk < n; kk++) {
        int k = kk < 0 ? PB_HOST(P) : kk;
        if (kk == PB_HOST(P)) continue;
        double mk = shfl(q_m, gb + k);
        V3 rk = shfl3(qr, gb + k);
        if (k == b) continue;
        V3 dr = qr - rk;
        pot += kG * mk / sqrt(dot(dr, dr));
    }
    double vi2 = dot(qv, qv);
    V3 ac = v3(0., 0., 0.);
    for (int kk = -1; kk < n; kk++) {
        int j = kk < 0 ? PB_HOST(P) : kk;
        if (kk == PB_HOST(P)) continue;
        double mj = shfl(q_m, gb + j);
        V3 rj = shfl3(qr, gb + j), vj = shfl3(qv, gb + j);
        double potj = shfl(pot, gb + j);
        bool en_j = (P.gr_enabled >> j) & 1u;
        if (j == b || !(en_i || en_j)) continue;
        V3 dr = qr - rj;
        double rij2 = dot(dr, dr);
        double rij = sqrt(rij2);
        double rij3 = rij2 * rij;
        double a1 = 4. * kInvC2 * pot;
        double a2 = kInvC2 * potj;
        double a3 = -vi2 * kInvC2;
        double a4 = -2. * dot(vj, vj) * kInvC2;
        double a5 = 4. * kInvC2 * dot(qv, vj);
        double a60 = dot(dr, vj);
        double a6 = 1.5 * kInvC2 * a60 * a60 / rij2;
        double factor1 = a1 + a2 + a3 + a4 + a5 + a6;
        double gm = kG * mj / rij3;
        ac = ac + (gm * factor1) * dr;
        V3 dv = qv - vj;
        double factor2 = dr.x * (4. * qv.x - 3. * vj.x) + dr.y * (4. * qv.y - 3. * vj.y) + dr.z * (4. * qv.z - 3. * vj.z);
        ac = ac + (gm * factor2 * kInvC2) * dv;
    }
    // substitution loop with the reference's (inverted) deviation test (Q8): one pass unless |a| < 1e-30
    V3 a_new = v3(0., 0., 0.);
    bool group_done = false;
    for (int it = 0; it < 10; it++) {
        if (!__any_sync(FULL, ro.valid && !group_done)) break;
        V3 a_old = a_new;
        V3 nc = v3(0., 0., 0.);
        V3 tot = an + a_old;
        for (int kk = -1; kk < n; kk++) {
            int j = kk < 0 ? PB_HOST(P) : kk;
            if (kk == PB_HOST(P)) continue;
            double mj = shfl(q_m, gb + j);
            V3 rj = shfl3(qr, gb + j);
            V3 tj = shfl3(tot, gb + j);
            bool en_j = (P.gr_enabled >> j) & 1u;
            if (j == b || !(en_i || en_j)) continue;
            V3 dr = qr - rj;
            double rij2 = dot(dr, dr);
            double rij = sqrt(rij2);
            double rij3 = rij2 * rij;
            double gm = kG * mj;
            double proj = dot(dr, tj) * (0.5 * kInvC2);
            nc = nc + (gm / rij3 * proj) * dr + (3.5 * kInvC2 * gm / rij) * tj;
        }
        V3 cand = ac + nc;
        const double dev_limit = 1.0e-30;
        double dev = 0.;
        if (ro.valid && en_i) {
            if (fabs(cand.x) < dev_limit) dev = fmax(dev, fabs(cand.x - a_old.x) / cand.x);
            if (fabs(cand.y) < dev_limit) dev = fmax(dev, fabs(cand.y - a_old.y) / cand.y);
            if (fabs(cand.z) < dev_limit) dev = fmax(dev, fabs(cand.z - a_old.z) / cand.z);
        }
        // group max (NaN compares false, as in the reference)
        double mx = dev;
        for (int off = PB_W(P) >> 1; off > 0; off >>= 1) { double o = shfl_xor(mx, off); mx = o > mx ? o : mx; }
        if (!group_done) {
            a_new = cand;
            if (mx < dev_limit) group_done = true;
        }
    }
    a_out = (ro.host || ro.g_on) ? a_new : v3(0., 0., 0.);
}

}  // namespace PB_NS
