// small_step.cuh — the step kernel for 2- and 3-body systems (BASELINE configs 1, 2, 3, 3-evolving, 5 and every other effect
// subset on such systems): lane = PLANET.
//
// The lane = body mapping of whfast_step.cuh leaves the host lane (and the padding lane of a 3-body system) without work in
// the force evaluations and in the Kepler drift: 1 of 2 lanes (N = 2) or 2 of 4 lanes (N = 3) do the arithmetic, and every
// exchange still pays a warp barrier and a shared-memory round trip. Here a system takes N - 1 lanes, one per non-host
// body; the host body (index 0) is carried REDUNDANTLY by every lane of the group (bit-identical copies: same operations
// on the same operands), so
//   * N = 2: one thread = one system, nothing is exchanged at all;
//   * N = 3: two lanes per system; the partner's terms of the ordered sums, its Kepler-drifted coordinates and its
//     contributions to the host sums travel by one shfl_xor each.
// Same arithmetic as whfast_step.cuh, operation by operation (the strict core in the reference's association order, the
// fast / exact / hybrid perturbation forces of forces_fast.cuh and exact_effects.cuh): the two kernels agree bit for bit
// in strict mode and to rounding in the others (tests/test_gpu_parity.py::test_small_system_specialisations_...).
//
// Reference path (file:line under /root/reference/src): WHFast::iterate integrator/whfast.rs:235-305; midpoint :322-466;
// transforms / jump / kick :495-672, 881-1155; Kepler drift :676-876; gravity particles/universe.rs:198-303; effects
// particles/universe.rs:428-614, effects/tides/{common,constant_time_lag}.rs, effects/rotational_flattening/{common,
// oblate_spheroid}.rs, effects/general_relativity.rs:177-456, effects/evolution.rs:449-546.
//
// Compile-time: N (2 | 3), coordinates (democratic heliocentric | Jacobi), effect set (FLAG_* mask; GR = Kidder1995; with
// SMALL_RT the ensemble's flag word selects among the compiled-in effects at run time), arithmetic mode, CTA size, and the
// passive-planet mapping (3 bodies, Jacobi, body 2 outside every effect: one thread per system). Host at index 0, spin
// integrated (at least one effect enabled). The force formulas are the ones of forces_fast.cuh / exact_effects.cuh with the
// host's quantities and the per-planet products held by the lane itself instead of being read from the host's column.
#include "whfast_kernel.cuh"

namespace PB_NS {
using namespace pb200;

extern __shared__ __align__(16) double pb_smem[];

// CTA size BLK (template parameter): 8 warps are resident per SM either way (<= 255 registers, ~100 slots per thread). Small
// ensembles take 32-thread CTAs (they spread over all SMs); large ones take one CTA of 8 warps per SM (N = 3) or two of 4
// (N = 2): warps that start together stay loosely in phase and share the instruction-cache lines of the ~100 KB loop
// body (N = 3: + 8-10 %, profiles/r2_small_variants.md).
#define PB_HIST_FIELDS 17
#define PB_TIDE_SCRATCH 13

// Per-thread shared-memory slots, [slot][thread]: what is read once per evaluation or once per step.
enum SmallSlot : int {
    // midpoint working set of the planet and of the (replicated) host: Kahan residuals (whfast.rs:117-119), originals,
    // increments (whfast.rs:333-337); (v, L) each
    P_ERR = 0, P_ORIG = 6, P_INCR = 12, H_ERR = 18, H_ORIG = 24, H_INCR = 30,
    K_M = 36, K_R, K_I, K_RG2, H_M, H_R, H_I, H_RG2,
    // exact forces: leading products of the reference's expressions (step-invariant operands only)
    X_T1, X_T2, X_HOSTK, X_FS0, X_FP0, K_R5, H_R5,
    C_AS, C_AP, C_INVM, C_INVMH, C_MGS,
    G_0, G_1, G_2, G_3, G_4, G_5, G_6, G_7, G_8, G_9, G_10, G_11, G_12,
    Z_MFM, Z_MURED, Z_MOM, Z_FMS, Z_FMP,
    Y_M, Y_MH, Y_I, Y_IH,            // refined reciprocals (strict.cuh, srcp) of m, M, I, I_host
    // fast forces: folded constants
    C_INVI, C_INVIH, C_KS, C_KP, C_ZP, C_ZH, C_DP1, C_DS1, C_MFA, C_SXS, C_BK,
    // coordinate transforms and gravity
    K_KMU, K_BACKW, K_MTOT, Y_MTOT, K_MP, K_BACKWP,
    K_RS2H, K_RR2H, K_RS2P, K_RR2P,  // squared collision / Roche distances of the pairs (host, planet) and (planet 1, planet 2)
    N_SMALL_SLOTS_DH,
    // Jacobi only: the step-invariant scalars of the recurrences (whfast.rs:881-963, 1026-1088)
    J_EI1 = N_SMALL_SLOTS_DH, J_PME1, J_EI2, J_PME2, J_MI, J_ET1, J_BEI1, J_ET0, J_BMI, J_ETAK,
    N_SMALL_SLOTS_JACOBI,
    // passive-planet build only (body 2 carried by the thread of body 1): its midpoint working set, base quantities, Kepler
    // mu, squared collision / Roche distances of the pair (host, body 2)
    P2_ERR = N_SMALL_SLOTS_JACOBI, P2_ORIG = P2_ERR + 6, P2_INCR = P2_ORIG + 6,
    K2_M = P2_INCR + 6, K2_R, K2_I, K2_RG2, K2_KMU, K2_RS2H, K2_RR2H,
    K2_L, K2_LY, K2_LZ, K2_S, K2_SY, K2_SZ,   // its angular momentum and spin (touched once per midpoint)
    N_SMALL_SLOTS_PASSIVE
};
template <int COORD, int BLK, bool PASSIVE> constexpr size_t small_smem_bytes() {
    return (size_t)(PASSIVE ? N_SMALL_SLOTS_PASSIVE : COORD == PB200_COORD_JACOBI ? N_SMALL_SLOTS_JACOBI : N_SMALL_SLOTS_DH) * BLK * sizeof(double);
}

template <int BLK>
struct SlT {
    volatile double* base;
    __device__ __forceinline__ double get(int i) const { return base[i * BLK]; }
    __device__ __forceinline__ void set(int i, double v) const { base[i * BLK] = v; }
    __device__ __forceinline__ V3 get3(int i) const { return v3(get(i), get(i + 1), get(i + 2)); }
    __device__ __forceinline__ void set3(int i, V3 v) const { set(i, v.x); set(i + 1, v.y); set(i + 2, v.z); }
    __device__ __forceinline__ srcp rcp(int val, int y) const { srcp r; r.b = get(val); r.y = get(y); return r; }
};

// partner lane of the group (N = 3 only; every lane of the warp takes part)
template <bool PAIR> __device__ __forceinline__ double xd(double v) { return PAIR ? __shfl_xor_sync(FULL, v, 1) : v; }
template <bool PAIR> __device__ __forceinline__ V3 x3(V3 v) { return v3(xd<PAIR>(v.x), xd<PAIR>(v.y), xd<PAIR>(v.z)); }
template <bool PAIR> __device__ __forceinline__ S3 x3(S3 v) { return strict(x3<PAIR>(plain(v))); }
// (planet 1's, planet 2's) from (mine, the partner's)
__device__ __forceinline__ double sel(bool c, double a, double b) { return c ? a : b; }
__device__ __forceinline__ S3 sel(bool c, S3 a, S3 b) { return s3(sd(sel(c, a.x.v, b.x.v)), sd(sel(c, a.y.v, b.y.v)), sd(sel(c, a.z.v, b.z.v))); }
__device__ __forceinline__ V3 sel(bool c, V3 a, V3 b) { return v3(sel(c, a.x, b.x), sel(c, a.y, b.y), sel(c, a.z, b.z)); }

// Effect set: FLAGS names what is compiled in; with SMALL_RT set the ensemble's own flag word selects among it at run time
// (uniform branches) — the catch-all builds for effect sets that have no compile-time build of their own.
enum : int { SMALL_RT = 1 << 10 };
template <int FLAGS> __device__ __forceinline__ bool has(const KParams& P, int f) { return (FLAGS & f) && (!(FLAGS & SMALL_RT) || (P.flags & f)); }

struct SmallState {
    S3 r, v, r0, v0;        // inertial position / velocity of the planet and of the host
    V3 L, s, L0, s0;        // angular momentum and spin (of the previous evaluation)
    double rs_s, rs_p;      // r . w_host, r . w_planet with the spins of the previous evaluation (Q3)
    S3 r2, v2;              // passive-planet build: body 2 (no effect acts on it; it drifts, is kicked and is checked)
};
struct SmallSys {
    double t, last_hist;
    unsigned int steps_done, n_hist_new, event_step;
    int status;
    unsigned int warnings;
    int hist_count;
    bool tswarn;
};
struct SmallRoles { bool t_on, f_on, g_on; };

// ---- constants (launch start and whenever a radius evolves). What the exact forces read is computed with `sd` in the
// reference's association order (see exact_effects.cuh / forces_fast.cuh::make_consts: the same expressions).
template <int N, int COORD, int FLAGS, bool PASSIVE, class Sl>
__device__ __forceinline__ void small_consts(const KParams& P, const Sl& sl, const SmallRoles& ro, bool valid, int b, size_t sys) {
    const size_t ns = (size_t)P.n_sys;
    double sigma = 0., k2t = 0., k2f = 0., mg = 1., sig_h = 0., k2t_h = 0., k2f_h = 0., Mg = 1.;
    if (valid) {
        const size_t i = (size_t)b * ns + sys;
        sigma = P.sigma[i]; k2t = P.k2t[i]; k2f = P.k2f[i]; mg = P.mass_g[i];
        sig_h = P.sigma[sys]; k2t_h = P.k2t[sys]; k2f_h = P.k2f[sys]; Mg = P.mass_g[sys];
    }
    const double m = sl.get(K_M), R = sl.get(K_R), I = sl.get(K_I), M = sl.get(H_M), Rh = sl.get(H_R), Ih = sl.get(H_I);
    const sd R_s = sd(R), R2_s = R_s * R_s, R4_s = R2_s * R2_s, R8_s = R4_s * R4_s;
    const sd Rh_s = sd(Rh), Rh2_s = Rh_s * Rh_s, Rh4_s = Rh2_s * Rh2_s, Rh8_s = Rh4_s * Rh4_s;
    const sd R5_s = R_s * R4_s, R10_s = R2_s * R8_s, Rh5_s = Rh_s * Rh4_s, Rh10_s = Rh2_s * Rh8_s;   // powi as LLVM expands it (Q10)
    const double R5 = R5_s.v, Rh5 = Rh5_s.v;
    sl.set(K_R5, R5); sl.set(H_R5, Rh5);
    const sd m_s = sd(m), M_s = sd(M);
    const sd m2_s = m_s * m_s, M2_s = M_s * M_s;
    const double m2 = m2_s.v, M2 = M2_s.v;
    const double gt = ro.t_on ? 1. : 0., gf = ro.f_on ? 1. : 0., gg = ro.g_on ? 1. : 0.;
    const double gts = P.tides_host_central ? gt : 0., gfs = P.flat_host_central ? gf : 0.;
    // constant_time_lag.rs:232-234, 243-245, 284-285, 291-296
    sl.set(C_AS, gts * (sd(4.5) * m2_s * Rh10_s * sd(sig_h)).v);
    sl.set(C_AP, gt * (sd(4.5) * M2_s * R10_s * sd(sigma)).v);
    sl.set(X_T1, (m2_s * Rh10_s * sd(sig_h)).v);
    sl.set(X_T2, (M2_s * R10_s * sd(sigma)).v);
    sl.set(X_HOSTK, (m2_s * Rh5_s * sd(k2t_h) + M2_s * R5_s * sd(k2t)).v);
    // oblate_spheroid.rs:37, 42 leading products
    sl.set(X_FS0, P.flat_host_central ? (m_s * sd(k2f_h)).v : 0.);
    sl.set(X_FP0, (M_s * sd(k2f)).v);
    sl.set(C_INVM, 1. / m); sl.set(C_INVMH, 1. / M);
    sl.set(C_INVI, 1. / I); sl.set(C_INVIH, 1. / Ih);
    sl.set(Y_M, make_rcp(m_s).y); sl.set(Y_MH, make_rcp(M_s).y); sl.set(Y_I, make_rcp(sd(I)).y); sl.set(Y_IH, make_rcp(sd(Ih)).y);
    const sd mgs_s = sd(Mg) + sd(mg);
    sl.set(C_MGS, gg * mgs_s.v);
    const sd msum_s = M_s + m_s, mdiff_s = M_s - m_s;
    const sd mured_s = (M_s * m_s) / msum_s;                      // general_relativity.rs:383
    sl.set(Z_MURED, mured_s.v);
    sl.set(Z_MFM, (mdiff_s / msum_s * msum_s).v);                 // mass_factor * star_planet_mass (:321, 336)
    sl.set(Z_MOM, (m_s / M_s).v);                                 // particle.mass / host.mass (:216)
    sl.set(Z_FMS, (sd(2.) + sd(1.5) * m_s / M_s).v);              // :390
    sl.set(Z_FMP, (sd(2.) + sd(1.5) * M_s / m_s).v);              // :419
    const sd f = sd(Mg) * sd(mg) / (mgs_s * mgs_s), f2 = f * f;   // general_relativity.rs:98
    sl.set(G_0, (sd(1.0) + sd(3.0) * f).v); sl.set(G_1, (sd(2.0) * (sd(2.0) + f)).v);
    sl.set(G_2, (sd(1.5) * f).v); sl.set(G_3, (sd(2.0) * (sd(2.0) - f)).v);
    sl.set(G_4, (sd(0.75) * (sd(12.0) + sd(29.0) * f)).v); sl.set(G_5, (f * (sd(3.0) - sd(4.0) * f)).v);
    sl.set(G_6, (sd(1.875) * f * (sd(1.0) - sd(3.0) * f)).v); sl.set(G_7, (sd(1.5) * f * (sd(3.0) - sd(4.0) * f)).v);
    sl.set(G_8, (sd(0.5) * f * (sd(13.0) - sd(4.0) * f)).v); sl.set(G_9, (sd(2.0) + sd(25.0) * f + sd(2.0) * f2).v);
    sl.set(G_10, (f * (sd(15.0) + sd(4.0) * f)).v); sl.set(G_11, (sd(4.0) + sd(41.0) * f + sd(8.0) * f2).v);
    sl.set(G_12, (sd(3.0) * f * (sd(3.0) + sd(2.0) * f)).v);
    // fast forces (forces_fast.cuh::make_consts)
    sl.set(C_BK, gt * (3.0 * kK2 * (m2 * Rh5 * k2t_h + M2 * R5 * k2t)));
    sl.set(C_KS, gfs * (m * k2f_h * Rh5)); sl.set(C_KP, gf * (M * k2f * R5));
    const double fa = gg * (kG * kInvC2);
    sl.set(C_ZP, I * (M / m)); sl.set(C_ZH, Ih * (m / M));
    sl.set(C_DP1, fa * I * ((2. + 1.5 * M / m) * mured_s.v)); sl.set(C_DS1, fa * Ih * ((2. + 1.5 * m / M) * mured_s.v));
    sl.set(C_MFA, fa * m); sl.set(C_SXS, fa * I * Ih);
    // collision distance of the pair (host, planet): universe.rs:229-233
    { const double rs = __dadd_rn(Rh, R); sl.set(K_RS2H, __dmul_rn(rs, rs)); }
    if (N == 3) {
        constexpr bool PAIR = !PASSIVE;
        const double Rp = PASSIVE ? sl.get(K2_R) : xd<PAIR>(R);
        const double rs = b == 1 ? __dadd_rn(R, Rp) : __dadd_rn(Rp, R);
        sl.set(K_RS2P, __dmul_rn(rs, rs));
        if (PASSIVE) { const double rs2 = __dadd_rn(Rh, Rp); sl.set(K2_RS2H, __dmul_rn(rs2, rs2)); }
    }
}

// ---- Universe::calculate_additional_effects for one planet and its share of the host sums, fast arithmetic
// (forces_fast.cuh::additional_effects with the host's quantities held by the lane itself).
// Out: acceleration and dL/dt of the planet; a_h / dl_h = this planet's contributions to the host's.
template <int FLAGS, bool HYB, class Sl>
__device__ __forceinline__ void small_effects_fast(const KParams& P, const Sl& sl, bool valid, int b, size_t sys, SmallState& q, V3 hr, double inv_d,
                                                   V3 hv, V3& a_p, V3& dl_p, V3& a_h, V3& dl_h, bool tide_save) {
    const double rs_s = q.rs_s, rs_p = q.rs_p;
    if (HYB) { q.s = plain(strict(q.L) / sl.rcp(K_I, Y_I)); q.s0 = plain(strict(q.L0) / sl.rcp(H_I, Y_IH)); }
    else { q.s = sl.get(C_INVI) * q.L; q.s0 = sl.get(C_INVIH) * q.L0; }
    const V3 sh = q.s0;
    const double w2 = HYB ? sdot(strict(q.s), strict(q.s)).v : dot(q.s, q.s);
    const double wh2 = HYB ? sdot(strict(sh), strict(sh)).v : dot(sh, sh);
    if (HYB) { q.rs_s = sdot(strict(hr), strict(sh)).v; q.rs_p = sdot(strict(hr), strict(q.s)).v; }
    else { q.rs_s = dot(hr, sh); q.rs_p = dot(hr, q.s); }
    const double inv_d2 = inv_d * inv_d, inv_d4 = inv_d2 * inv_d2;
    const double radvel = dot(hr, hv) * inv_d;
    const V3 rxv = cross(hr, hv);
    const V3 cs = cross(hr, sh), cp = cross(hr, q.s);
    const double inv_m = sl.get(C_INVM), inv_M = sl.get(C_INVMH);
    double Kr = 0., Kv = 0., Pcp = 0., Hcs = 0.;
    V3 F = v3(0., 0., 0.);
    dl_p = v3(0., 0., 0.); dl_h = v3(0., 0., 0.);
    if (has<FLAGS>(P, FLAG_TIDES)) {
        const double inv_d6 = inv_d4 * inv_d2, inv_d7 = inv_d6 * inv_d;
        const double FodS = sl.get(C_AS) * inv_d6, FodP = sl.get(C_AP) * inv_d6;
        const double Fos = FodS * inv_d, Fop = FodP * inv_d;
        const double ks = Fos * inv_d, kp = Fop * inv_d;
        Kr = -inv_d * (sl.get(C_BK) * inv_d7 + (2.0 * inv_d) * ((Fos + Fop) * radvel));
        Kv = -(ks + kp);
        F = v3(-(ks * cs.x + kp * cp.x), -(ks * cs.y + kp * cp.y), -(ks * cs.z + kp * cp.z));
        const double krp = kp * rs_p, krs = ks * rs_s;
        dl_p = v3(krp * hr.x + kp * rxv.x - FodP * q.s.x, krp * hr.y + kp * rxv.y - FodP * q.s.y, krp * hr.z + kp * rxv.z - FodP * q.s.z);
        dl_h = v3(krs * hr.x + ks * rxv.x - FodS * sh.x, krs * hr.y + ks * rxv.y - FodS * sh.y, krs * hr.z + ks * rxv.z - FodS * sh.z);
        if (tide_save && valid) {
            const size_t ns = (size_t)P.n_sys;
            double* ts = P.tide_scratch + (size_t)b * ns + sys;
            const size_t cs_ = (size_t)P.n_bodies * ns;
            const double d = dot(hr, hr) * inv_d;
            ts[0 * cs_] = hr.x; ts[1 * cs_] = hr.y; ts[2 * cs_] = hr.z;
            ts[3 * cs_] = hv.x; ts[4 * cs_] = hv.y; ts[5 * cs_] = hv.z;
            ts[6 * cs_] = d; ts[7 * cs_] = radvel; ts[8 * cs_] = Fop;
            ts[9 * cs_] = -3.0 * Fop * radvel * inv_d;
            ts[10 * cs_] = dl_p.x; ts[11 * cs_] = dl_p.y; ts[12 * cs_] = dl_p.z;
        }
    }
    if (has<FLAGS>(P, FLAG_FLAT)) {
        const double inv_d5 = inv_d4 * inv_d;
        const double k_s = sl.get(C_KS), k_p = sl.get(C_KP);
        const double KsRs = k_s * rs_s, KpRp = k_p * rs_p;
        const double Fos = -KsRs * inv_d5, Fop = -KpRp * inv_d5;
        const double q1 = k_s * wh2 + k_p * w2;
        const double q2 = KsRs * rs_s + KpRp * rs_p;
        Kr += inv_d5 * ((2.5 * inv_d2) * q2 - 0.5 * q1);
        F = v3(F.x + Fop * q.s.x + Fos * sh.x, F.y + Fop * q.s.y + Fos * sh.y, F.z + Fop * q.s.z + Fos * sh.z);
        Pcp = -Fop; Hcs = -Fos;
    }
    if (has<FLAGS>(P, FLAG_GR)) {
        const double v2 = dot(hv, hv);
        const double mgs = sl.get(C_MGS);
        const double A = mgs * inv_d2 * kInvC2;
        const double u = mgs * inv_d;
        const double rv2 = radvel * radvel;
        double rad = -A * (sl.get(G_0) * v2 - sl.get(G_1) * u - sl.get(G_2) * rv2);
        double orth = A * sl.get(G_3) * radvel;
        rad += -A * (sl.get(G_4) * (u * u) + sl.get(G_5) * (v2 * v2) + sl.get(G_6) * (rv2 * rv2) - sl.get(G_7) * rv2 * v2 - sl.get(G_8) * u * v2 - sl.get(G_9) * u * rv2);
        orth += 0.5 * A * radvel * (sl.get(G_10) * v2 - sl.get(G_11) * u - sl.get(G_12) * rv2);
        const double m = sl.get(K_M);
        Kr += m * (rad * inv_d);
        Kv += m * orth;
        const double Ip = sl.get(K_I), Ih = sl.get(H_I), zp = sl.get(C_ZP), zh = sl.get(C_ZH);
        const V3 S = v3(Ip * q.s.x + Ih * sh.x, Ip * q.s.y + Ih * sh.y, Ip * q.s.z + Ih * sh.z);
        const V3 Z = v3(zp * q.s.x + zh * sh.x, zp * q.s.y + zh * sh.y, zp * q.s.z + zh * sh.z);
        const V3 A1 = S + Z;
        const V3 A3 = A1 + S;
        const V3 A7 = v3(3. * A1.x + S.x, 3. * A1.y + S.y, 3. * A1.z + S.z);
        const double mfa = sl.get(C_MFA);
        const double s1 = 6. * mfa * inv_d2, s3_ = 3. * mfa * (radvel * inv_d);
        const V3 e2 = cross(hv, A7), e3 = cross(hr, A3);
        F = v3(F.x + s1 * (hr.x * rxv.x * A1.x) - mfa * e2.x + s3_ * e3.x,
               F.y + s1 * (hr.y * rxv.y * A1.y) - mfa * e2.y + s3_ * e3.y,
               F.z + s1 * (hr.z * rxv.z * A1.z) - mfa * e2.z + s3_ * e3.z);
        const double sxs_k = sl.get(C_SXS);
        const double c3 = 3. * sxs_k * inv_d2;
        Pcp += c3 * q.rs_s;
        Hcs += c3 * q.rs_p;
        const V3 wxw = cross(q.s, sh), jp = cross(rxv, q.s), js = cross(rxv, sh);
        const double dp1 = sl.get(C_DP1), ds1 = sl.get(C_DS1);
        dl_p = v3(dl_p.x + dp1 * jp.x + sxs_k * wxw.x, dl_p.y + dp1 * jp.y + sxs_k * wxw.y, dl_p.z + dp1 * jp.z + sxs_k * wxw.z);
        dl_h = v3(dl_h.x + ds1 * js.x - sxs_k * wxw.x, dl_h.y + ds1 * js.y - sxs_k * wxw.y, dl_h.z + ds1 * js.z - sxs_k * wxw.z);
    }
    F = v3(F.x + Kr * hr.x + Kv * hv.x, F.y + Kr * hr.y + Kv * hv.y, F.z + Kr * hr.z + Kv * hv.z);
    dl_p = v3(dl_p.x + Pcp * cp.x, dl_p.y + Pcp * cp.y, dl_p.z + Pcp * cp.z);
    dl_h = v3(dl_h.x + Hcs * cs.x, dl_h.y + Hcs * cs.y, dl_h.z + Hcs * cs.z);
    a_p = inv_m * F;
    a_h = (-inv_M) * F;
}

// ---- The same in the reference's own arithmetic (exact_effects.cuh::additional_effects_exact, operation by operation).
// Out: the planet's acceleration and dL/dt, and the HOST's (sums over the planets in index order, partner's terms by
// shuffle): bit-identical in every lane of the group.
template <int N, int FLAGS, bool PASSIVE, class Sl>
__device__ __forceinline__ void small_effects_exact(const KParams& P, const Sl& sl, const SmallRoles& ro, bool valid, int b, size_t sys, SmallState& q,
                                                    S3 hr, sd dist, S3 hv, S3& a_p, S3& dl_p, S3& a_h, S3& dl_h, bool tide_save) {
    const bool first = b == 1;
    const sd zero = sd(0.);
    const S3 zero3 = s3(zero, zero, zero);
    // the host sums of two vectors: 0 + planet 1's + planet 2's (the reference's serial loops)
    constexpr bool PAIR = N == 3 && !PASSIVE;
    auto host_sums = [&](S3 u, S3 w, S3& su, S3& sw) {
        if (PASSIVE) {
            // body 2 is no OrbitingBody of any effect: its terms are zeros (added like the general kernel adds them)
            su = (zero3 + u) + zero3; sw = (zero3 + w) + zero3;
        } else if (N == 3) {
            const S3 uo = x3<PAIR>(u), wo = x3<PAIR>(w);
            su = (zero3 + sel(first, u, uo)) + sel(first, uo, u);
            sw = (zero3 + sel(first, w, wo)) + sel(first, wo, w);
        } else { su = zero3 + u; sw = zero3 + w; }
    };
    const sd rs_s = sd(q.rs_s), rs_p = sd(q.rs_p);
    const sd I = sd(sl.get(K_I)), Ih = sd(sl.get(H_I));
    const S3 s = strict(q.L) / sl.rcp(K_I, Y_I);
    const S3 sh = strict(q.L0) / sl.rcp(H_I, Y_IH);
    const sd w2 = (s.x * s.x) + (s.y * s.y) + (s.z * s.z);
    const sd wh2 = (sh.x * sh.x) + (sh.y * sh.y) + (sh.z * sh.z);
    q.s = plain(s); q.s0 = plain(sh);
    q.rs_s = sdot(hr, sh).v; q.rs_p = sdot(hr, s).v;
    const srcp rD = make_rcp(dist);
    const sd radvel = (hr.x * hv.x + hr.y * hv.y + hr.z * hv.z) / rD;
    const sd normv2 = hv.x * hv.x + hv.y * hv.y + hv.z * hv.z;
    const sd d2 = dist * dist, d4 = d2 * d2;
    const sd d5 = dist * d4, d7 = (dist * d2) * d4;
    const srcp rD7 = make_rcp(d7);
    const sd m = sd(sl.get(K_M)), M = sd(sl.get(H_M));
    const sd inv_m = sd(sl.get(C_INVM)), inv_M = sd(sl.get(C_INVMH));
    const sd neg_inv_M = sd(-1.0) * inv_M;
    S3 a = zero3, ah = zero3;
    S3 td = zero3, fd = zero3, gd = zero3, tdh = zero3, fdh = zero3, gdh = zero3;
    if (has<FLAGS>(P, FLAG_TIDES)) {
        const sd cs = sd(sl.get(C_AS)), cp = sd(sl.get(C_AP));
        const sd t1 = sd(sl.get(X_T1)), t2 = sd(sl.get(X_T2)), host_k = sd(sl.get(X_HOSTK));
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
        if (tide_save && valid) {
            const size_t ns = (size_t)P.n_sys;
            double* ts = P.tide_scratch + (size_t)b * ns + sys;
            const size_t cs_ = (size_t)P.n_bodies * ns;
            ts[0 * cs_] = hr.x.v; ts[1 * cs_] = hr.y.v; ts[2 * cs_] = hr.z.v;
            ts[3 * cs_] = hv.x.v; ts[4 * cs_] = hv.y.v; ts[5 * cs_] = hv.z.v;
            ts[6 * cs_] = dist.v; ts[7 * cs_] = radvel.v; ts[8 * cs_] = orth_p.v; ts[9 * cs_] = diss_pm.v;
            ts[10 * cs_] = t_dl.x.v; ts[11 * cs_] = t_dl.y.v; ts[12 * cs_] = t_dl.z.v;
        }
        S3 sF, sN;
        host_sums(xF, xN, sF, sN);
        a = a + t_acc; td = t_dl;
        ah = ah + s3(neg_inv_M * sF.x, neg_inv_M * sF.y, neg_inv_M * sF.z); tdh = sN;
    }
    if (has<FLAGS>(P, FLAG_FLAT)) {
        const sd Rh5 = sd(sl.get(H_R5)), R5 = sd(sl.get(K_R5));
        const sd fs0 = sd(sl.get(X_FS0)), fp0 = sd(sl.get(X_FP0));
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
        host_sums(xF, xN, sF, sN);
        a = a + f_acc; fd = f_dl;
        ah = ah + s3(neg_inv_M * sF.x, neg_inv_M * sF.y, neg_inv_M * sF.z); fdh = sN;
    }
    if (has<FLAGS>(P, FLAG_GR)) {
        // general_relativity.rs:177-456 (Kidder1995)
        const sd c2 = sd(kC2);
        const sd mgs = sd(sl.get(C_MGS));
        const sd g0 = sd(sl.get(G_0)), g1 = sd(sl.get(G_1)), g2 = sd(sl.get(G_2)), g3 = sd(sl.get(G_3)), g4 = sd(sl.get(G_4)), g5 = sd(sl.get(G_5)),
                 g6 = sd(sl.get(G_6)), g7 = sd(sl.get(G_7)), g8 = sd(sl.get(G_8)), g9 = sd(sl.get(G_9)), g10 = sd(sl.get(G_10)), g11 = sd(sl.get(G_11));
        const sd normv = ssqrt(normv2);
        const srcp rV = make_rcp(normv);
        const sd rv2 = radvel * radvel;
        const srcp rD2c2 = make_rcp(d2 * c2);
        const sd pre = -mgs / rD2c2;
        const sd mgd = mgs / rD;
        const sd radial1 = pre * (g0 * normv2 - g1 * mgs / rD - g2 * rv2);
        const sd orth1 = mgs / rD2c2 * g3 * radvel * normv;
        S3 a1;
        a1.x = radial1 * hr.x / rD + orth1 * hv.x / rV;
        a1.y = radial1 * hr.y / rD + orth1 * hv.y / rV;
        a1.z = radial1 * hr.z / rD + orth1 * hv.z / rV;
        const sd v4 = normv2 * normv2, rv4 = rv2 * rv2;
        const sd radial2 = pre
            * (g4 * (mgs * mgs / d2)
               + g5 * v4
               + g6 * rv4
               - g7 * rv2 * normv2
               - g8 * mgd * normv2
               - g9 * mgd * rv2);
        const sd orth2 = pre * sd(-0.5) * radvel * (g10 * normv2 - g11 * mgd - sd(sl.get(G_12)) * rv2);
        S3 a2;
        a2.x = radial2 * hr.x / rD + orth2 * hv.x;
        a2.y = radial2 * hr.y / rD + orth2 * hv.y;
        a2.z = radial2 * hr.z / rD + orth2 * hv.z;
        const srcp rM = sl.rcp(H_M, Y_MH);
        const srcp rm = sl.rcp(K_M, Y_M);
        const S3 Ls = s3(Ih * sh.x, Ih * sh.y, Ih * sh.z), Lp = s3(I * s.x, I * s.y, I * s.z);
        const S3 nn = hr / rD;
        const sd mfm = sd(sl.get(Z_MFM));
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
        const sd mom = sd(sl.get(Z_MOM));
        S3 g_acc = zero3, x1 = zero3, x2 = zero3;
        if (ro.g_on) {
            g_acc = s3(a1.x + a2.x + a3.x, a1.y + a2.y + a3.y, a1.z + a2.z + a3.z);
            x1 = s3(mom * a1.x, mom * a1.y, mom * a1.z);
            x2 = s3(mom * a2.x, mom * a2.y, mom * a2.z);
        }
        S3 s1, s2;
        host_sums(x1, x2, s1, s2);
        const sd mu = sd(sl.get(Z_MURED));
        const S3 Lo = s3(mu * (hr.y * hv.z - hr.z * hv.y), mu * (hr.z * hv.x - hr.x * hv.z), mu * (hr.x * hv.y - hr.y * hv.x));
        const sd fms = sd(sl.get(Z_FMS)), fmp = sd(sl.get(Z_FMP));
        const sd a1x = fms * (Lo.y * Ls.z - Lo.z * Ls.y), a1y = fms * (Lo.z * Ls.x - Lo.x * Ls.z), a1z = fms * (Lo.x * Ls.y - Lo.y * Ls.x);
        const sd a2x = Lp.y * Ls.z - Lp.z * Ls.y, a2y = Lp.z * Ls.x - Lp.x * Ls.z, a2z = Lp.x * Ls.y - Lp.y * Ls.x;
        const sd spp = nn.x * Lp.x + nn.y * Lp.y + nn.z * Lp.z;
        const sd a3x = sd(3.) * spp * (nn.y * Ls.z - nn.z * Ls.y), a3y = sd(3.) * spp * (nn.z * Ls.x - nn.x * Ls.z), a3z = sd(3.) * spp * (nn.x * Ls.y - nn.y * Ls.x);
        const S3 hdl = s3(fa * (a1x - a2x + a3x), fa * (a1y - a2y + a3y), fa * (a1z - a2z + a3z));
        const sd b1x = fmp * (Lo.y * Lp.z - Lo.z * Lp.y), b1y = fmp * (Lo.z * Lp.x - Lo.x * Lp.z), b1z = fmp * (Lo.x * Lp.y - Lo.y * Lp.x);
        const sd b2x = Ls.y * Lp.z - Ls.z * Lp.y, b2y = Ls.z * Lp.x - Ls.x * Lp.z, b2z = Ls.x * Lp.y - Ls.y * Lp.x;
        const sd ssp = nn.x * Ls.x + nn.y * Ls.y + nn.z * Ls.z;
        const sd b3x = sd(3.) * ssp * (nn.y * Lp.z - nn.z * Lp.y), b3y = sd(3.) * ssp * (nn.z * Lp.x - nn.x * Lp.z), b3z = sd(3.) * ssp * (nn.x * Lp.y - nn.y * Lp.x);
        S3 g_dl = zero3, x3_ = zero3, x4 = zero3;
        if (ro.g_on) {
            g_dl = s3(fa * (b1x - b2x + b3x), fa * (b1y - b2y + b3y), fa * (b1z - b2z + b3z));
            x3_ = s3(mom * a3.x, mom * a3.y, mom * a3.z);
            x4 = hdl;
        }
        S3 s3_, s4;
        host_sums(x3_, x4, s3_, s4);
        const sd m1 = sd(-1.0);
        a = a + g_acc; gd = g_dl;
        ah = ah + s3(m1 * s1.x + m1 * s2.x + m1 * s3_.x, m1 * s1.y + m1 * s2.y + m1 * s3_.y, m1 * s1.z + m1 * s2.z + m1 * s3_.z);
        gdh = s4;
    }
    const S3 wd = zero3;   // no wind in these builds
    a_p = a; a_h = ah;
    dl_p = s3(td.x + fd.x + gd.x + wd.x, td.y + fd.y + gd.y + wd.y, td.z + fd.z + gd.z + wd.z);
    dl_h = s3(tdh.x + fdh.x + gdh.x + wd.x, tdh.y + fdh.y + gdh.y + wd.y, tdh.z + fdh.z + gdh.z + wd.z);
}

// effects/evolution.rs:516-546 for one body. R / I live in slots (slot_r, slot_i), rg2 in slot_g. Returns true when something changed.
template <class Sl>
__device__ __forceinline__ bool small_evolve(const KParams& P, const Sl& sl, int body, size_t sys, double t, bool commit, bool writer,
                                             int slot_m, int slot_r, int slot_i, int slot_g) {
    if (!commit) return false;
    const int ti = P.evo_table[body];
    if (ti < 0) return false;
    const DevTable& T = P.tables[ti];
    const double R = sl.get(slot_r), rg2 = sl.get(slot_g);
    const int i = table_upper(T.time, T.n_rows, t);
    const double nr = T.interp_radius ? table_interp(T.time, T.radius, T.n_rows, i, t) : R;
    const double ng = ((P.evo_rg2 >> body) & 1u) ? table_interp(T.time, T.rg2, T.n_rows, i, t) : rg2;
    if (nr != R || ng != rg2) {
        const double I = (sd(sl.get(slot_m)) * sd(ng) * (sd(nr) * sd(nr))).v;
        sl.set(slot_r, nr); sl.set(slot_i, I); sl.set(slot_g, ng);
        if (writer) {
            const size_t idx = (size_t)body * (size_t)P.n_sys + sys;
            P.radius[idx] = nr; P.rg2[idx] = ng; P.moi[idx] = I;
        }
        return true;
    }
    return false;
}

// PASSIVE (3 bodies, Jacobi): body 2 is no OrbitingBody of any effect (the circumbinary planet of config 5). In the two-lane
// mapping its lane would execute every force evaluation next to body 1's lane for nothing, so here ONE thread carries the
// system: body 1 (with the forces), the host, and body 2, which only drifts (a second pass through the Kepler solver), is
// kicked and is checked. Half the instructions per system-step, no exchange at all.
template <int N, int COORD, int FLAGS, int ARITH, int BLK, bool PASSIVE>
__global__ void __launch_bounds__(BLK, 256 / BLK) small_steps_kernel(const __grid_constant__ KParams P, unsigned long long n_steps) {
    constexpr int W = PASSIVE ? 1 : N - 1;
    constexpr bool PAIR = W == 2;        // two lanes per system: the partner's terms travel by shuffle
    constexpr bool JAC = COORD == PB200_COORD_JACOBI;
    static_assert(N == 2 || N == 3, "small_steps_kernel: 2 or 3 bodies");
    static_assert(!JAC || N == 3, "Jacobi build: 3 bodies");
    static_assert(!PASSIVE || JAC, "passive-planet build: 3 bodies, Jacobi coordinates");
    // ---- time slicing (see whfast_step.cuh): ticket -> (piece, group)
    __shared__ unsigned int s_ticket;
    unsigned int piece = 0, group = blockIdx.x;
    if (P.n_pieces > 1) {
        if (threadIdx.x == 0) s_ticket = atomicAdd(P.sched, 1u);
        __syncthreads();
        piece = s_ticket / P.n_groups; group = s_ticket % P.n_groups;
        if (piece > 0) {
            if (threadIdx.x == 0) {
                const volatile unsigned int* flag = P.sched + 1 + group;
                while (*flag < piece) __nanosleep(500);
                __threadfence();
            }
            __syncthreads();
        }
    }
    const unsigned long long step_begin = n_steps * piece / P.n_pieces, step_end = n_steps * (piece + 1ull) / P.n_pieces;
    const size_t gtid = (size_t)group * blockDim.x + threadIdx.x;
    const int p = (int)(gtid & (size_t)(W - 1));
    const int b = p + 1;                 // this lane's body
    const bool first = b == 1;
    const size_t sys_raw = W == 2 ? (gtid >> 1) : gtid;
    const bool sys_ok = sys_raw < (size_t)P.n_sys;
    // padding lanes shadow the LAST system, which this same CTA owns: its loads (kernel start) precede its stores (kernel end /
    // piece hand-over), so no lane ever reads state that another CTA is writing; padding lanes never store
    const size_t sys = sys_ok ? sys_raw : (size_t)P.n_sys - 1;
    const bool valid = sys_ok;
    const bool writer = valid && first;  // the lane that stores the host body and the per-system words
    SmallRoles ro;
    ro.t_on = (P.tides_orbiting >> b) & 1u; ro.f_on = (P.flat_orbiting >> b) & 1u; ro.g_on = (P.gr_orbiting >> b) & 1u;
    typedef SlT<BLK> Sl;
    Sl sl;
    sl.base = pb_smem + threadIdx.x;
    const size_t ns = (size_t)P.n_sys;
    const size_t cs = (size_t)N * ns;

    SmallState q;
    SmallSys st;
    {
        auto ld3 = [&](const double* a, int body) { const size_t i = (size_t)body * ns + sys; return v3(ldm(a + i), ldm(a + i + cs), ldm(a + i + 2 * cs)); };
        q.r = strict(ld3(P.pos, b)); q.v = strict(ld3(P.vel, b)); q.L = ld3(P.L, b); q.s = ld3(P.spin, b);
        q.r0 = strict(ld3(P.pos, 0)); q.v0 = strict(ld3(P.vel, 0)); q.L0 = ld3(P.L, 0); q.s0 = ld3(P.spin, 0);
        q.rs_s = 0.; q.rs_p = 0.;
        sl.set3(P_ERR, ld3(P.verr, b)); sl.set3(P_ERR + 3, ld3(P.lerr, b));
        sl.set3(H_ERR, ld3(P.verr, 0)); sl.set3(H_ERR + 3, ld3(P.lerr, 0));
        if (PASSIVE) {
            q.r2 = strict(ld3(P.pos, 2)); q.v2 = strict(ld3(P.vel, 2)); sl.set3(K2_L, ld3(P.L, 2)); sl.set3(K2_S, ld3(P.spin, 2));
            sl.set3(P2_ERR, ld3(P.verr, 2)); sl.set3(P2_ERR + 3, ld3(P.lerr, 2));
            const size_t i2 = (size_t)2 * ns + sys;
            sl.set(K2_M, P.mass[i2]); sl.set(K2_R, ldm(P.radius + i2)); sl.set(K2_I, ldm(P.moi + i2)); sl.set(K2_RG2, ldm(P.rg2 + i2));
        }
        const size_t i = (size_t)b * ns + sys;
        sl.set(K_M, P.mass[i]); sl.set(K_R, ldm(P.radius + i)); sl.set(K_I, ldm(P.moi + i)); sl.set(K_RG2, ldm(P.rg2 + i));
        sl.set(H_M, P.mass[sys]); sl.set(H_R, ldm(P.radius + sys)); sl.set(H_I, ldm(P.moi + sys)); sl.set(H_RG2, ldm(P.rg2 + sys));
        st.t = ldm(P.t + sys); st.last_hist = ldm(P.last_hist + sys);
        st.tswarn = ldm(P.tswarn + sys) != 0; st.status = ldm(P.status + sys); st.warnings = ldm(P.warnings + sys); st.hist_count = ldm(P.hist_count + sys);
        if (!sys_ok) { st.status = PB200_STATUS_COMPLETED; st.tswarn = true; }
        st.steps_done = 0; st.n_hist_new = 0; st.event_step = 0;
    }
    bool alive = sys_ok && st.status == PB200_STATUS_OK;
    small_consts<N, COORD, FLAGS, PASSIVE>(P, sl, ro, valid, b, sys);
    {
        // Roche distances (universe.rs:177-196: filled for lower index < higher index)
        const double rr = __ldg(P.roche + ((size_t)(0 * N + b)) * ns + sys);
        sl.set(K_RR2H, __dmul_rn(rr, rr));
        if (N == 3) { const double rp = __ldg(P.roche + ((size_t)(1 * N + 2)) * ns + sys); sl.set(K_RR2P, __dmul_rn(rp, rp)); }
        if (PASSIVE) { const double r2h = __ldg(P.roche + ((size_t)(0 * N + 2)) * ns + sys); sl.set(K2_RR2H, __dmul_rn(r2h, r2h)); }
        // constants of the transforms, strict and in the reference's order (whfast_step.cuh, kernel prologue)
        const sd m_s = sd(sl.get(K_M)), M_s = sd(sl.get(H_M));
        const sd mg_s = sd(P.mass_g[(size_t)b * ns + sys]), Mg_s = sd(P.mass_g[sys]);
        const sd mp_s = PASSIVE ? sd(sl.get(K2_M)) : sd(xd<PAIR>(m_s.v));
        const sd mgp_s = PASSIVE ? sd(P.mass_g[(size_t)2 * ns + sys]) : sd(xd<PAIR>(mg_s.v));
        const sd m1 = first ? m_s : mp_s, m2 = first ? mp_s : m_s;        // masses of planets 1 and 2 (N = 3)
        const sd mg1 = first ? mg_s : mgp_s, mg2 = first ? mgp_s : mg_s;
        sd mtot = JAC ? M_s : sd(0.) + M_s;
        sd mu = Mg_s;
        sd eta_k, mu_k;
        mtot = mtot + m1; mu = mu + mg1;
        eta_k = mtot; mu_k = mu;
        if (N == 3) { mtot = mtot + m2; mu = mu + mg2; if (!first) { eta_k = mtot; mu_k = mu; } }
        if (PASSIVE) sl.set(K2_KMU, mu.v);   // body 2's Kepler mu: the cumulative gravitational mass up to and including it
        sl.set(K_KMU, JAC ? mu_k.v : Mg_s.v);
        sl.set(K_BACKW, (m_s / M_s).v);
        sl.set(K_BACKWP, (mp_s / M_s).v);
        sl.set(K_MP, mp_s.v);
        sl.set(K_MTOT, mtot.v); sl.set(Y_MTOT, make_rcp(mtot).y);
        if (JAC) {
            const sd one = sd(1.);
            sd eta = M_s;
            const sd ei1 = one / eta; eta = eta + m1; const sd pme1 = eta * ei1;
            const sd ei2 = one / eta; eta = eta + m2; const sd pme2 = eta * ei2;
            const sd mi = one / eta;
            sd et = mtot;
            const sd bei2 = one / et; (void)bei2;   // = mi (the same operands)
            et = et - m2; const sd et1 = et; const sd bei1 = one / et;
            et = et - m1; const sd et0 = et; const sd bmi = one / et;
            sl.set(J_EI1, ei1.v); sl.set(J_PME1, pme1.v); sl.set(J_EI2, ei2.v); sl.set(J_PME2, pme2.v); sl.set(J_MI, mi.v);
            sl.set(J_ET1, et1.v); sl.set(J_BEI1, bei1.v); sl.set(J_ET0, et0.v); sl.set(J_BMI, bmi.v); sl.set(J_ETAK, eta_k.v);
        }
    }
    const sd zero = sd(0.);
    const S3 zero3 = s3(zero, zero, zero);
    const sd dt_s = sd(P.dt), hdt_s = sd(P.half_dt);
    S3 anew = zero3, anew0 = zero3;     // Newtonian acceleration of the planet / the host (last gravity evaluation)
    S3 anew2 = zero3;                   // ... of body 2 (passive-planet build)

    auto store_state = [&]() {
        if (!valid) return;
        auto st3 = [&](double* a, int body, V3 x) { const size_t i = (size_t)body * ns + sys; a[i] = x.x; a[i + cs] = x.y; a[i + 2 * cs] = x.z; };
        st3(P.pos, b, plain(q.r)); st3(P.vel, b, plain(q.v)); st3(P.L, b, q.L); st3(P.spin, b, q.s);
        st3(P.verr, b, sl.get3(P_ERR)); st3(P.lerr, b, sl.get3(P_ERR + 3));
        if (PASSIVE) {
            st3(P.pos, 2, plain(q.r2)); st3(P.vel, 2, plain(q.v2)); st3(P.L, 2, sl.get3(K2_L)); st3(P.spin, 2, sl.get3(K2_S));
            st3(P.verr, 2, sl.get3(P2_ERR)); st3(P.lerr, 2, sl.get3(P2_ERR + 3));
        }
        if (writer) {
            st3(P.pos, 0, plain(q.r0)); st3(P.vel, 0, plain(q.v0)); st3(P.L, 0, q.L0); st3(P.spin, 0, q.s0);
            st3(P.verr, 0, sl.get3(H_ERR)); st3(P.lerr, 0, sl.get3(H_ERR + 3));
            P.t[sys] = st.t; P.last_hist[sys] = st.last_hist;
            const unsigned long long it0 = ldm(P.iteration + sys);
            P.iteration[sys] = it0 + st.steps_done;
            P.n_hist[sys] = ldm(P.n_hist + sys) + st.n_hist_new;
            if (st.status != PB200_STATUS_OK) P.event_iteration[sys] = it0 + st.event_step;
            P.tswarn[sys] = st.tswarn ? 1ull : 0ull;
            P.status[sys] = st.status; P.warnings[sys] = st.warnings; P.hist_count[sys] = st.hist_count;
        }
    };
    auto evolve_all = [&](double t, bool commit) {
        // the planet by its lane, the host by every lane of the group (one of them stores)
        bool ch = small_evolve(P, sl, b, sys, t, commit && valid, valid, K_M, K_R, K_I, K_RG2);
        ch |= small_evolve(P, sl, 0, sys, t, commit && valid, writer, H_M, H_R, H_I, H_RG2);
        if (PASSIVE) ch |= small_evolve(P, sl, 2, sys, t, commit && valid, valid, K2_M, K2_R, K2_I, K2_RG2);
        if (__any_sync(FULL, ch)) small_consts<N, COORD, FLAGS, PASSIVE>(P, sl, ro, valid, b, sys);
    };

#pragma unroll 1
    for (unsigned long long step = step_begin; step < step_end; step++) {
        if (!__any_sync(FULL, alive)) break;
        // ---- historic snapshot (whfast.rs:237-261, output.rs:119-163)
        {
            const bool first_snap = st.last_hist < 0.;
            const bool due = __dadd_rn(st.last_hist, P.hist_period) <= st.t;
            const bool snap = alive && (first_snap || due);
            if (__any_sync(FULL, snap)) {
                if (has<FLAGS>(P, FLAG_EVO)) evolve_all(st.t, snap);
                if (snap) {
                    // spin = L / I (universe.rs:305-316, common.rs:9-11): it changes the live state too
                    q.s = plain(strict(q.L) / sd(sl.get(K_I)));
                    q.s0 = plain(strict(q.L0) / sd(sl.get(H_I)));
                    if (PASSIVE) sl.set3(K2_S, plain(strict(sl.get3(K2_L)) / sd(sl.get(K2_I))));
                }
                if (snap && valid && st.hist_count < P.hist_capacity) {
                    auto record = [&](int body, V3 r_, V3 s_, V3 v_, int sm, int sr, int sg, double denergy) {
                        const size_t i = (size_t)body * ns + sys;
                        double* h = P.hist + (size_t)st.hist_count * PB_HIST_FIELDS * cs + i;
                        h[0 * cs] = st.t;
                        h[1 * cs] = r_.x; h[2 * cs] = r_.y; h[3 * cs] = r_.z;
                        h[4 * cs] = s_.x; h[5 * cs] = s_.y; h[6 * cs] = s_.z;
                        h[7 * cs] = v_.x; h[8 * cs] = v_.y; h[9 * cs] = v_.z;
                        h[10 * cs] = sl.get(sm); h[11 * cs] = sl.get(sr); h[12 * cs] = sl.get(sg);
                        h[13 * cs] = P.k2t[i]; h[14 * cs] = P.sigma[i]; h[15 * cs] = denergy;
                        h[16 * cs] = 0.;
                    };
                    double denergy = 0.;
                    if (has<FLAGS>(P, FLAG_TIDES) && ro.t_on) {
                        // tides/common.rs:263-279 with the internals left by the last evaluation and the fresh spin
                        const size_t i = (size_t)b * ns + sys;
                        double ts[PB_TIDE_SCRATCH];
                        for (int k = 0; k < PB_TIDE_SCRATCH; k++) ts[k] = ldm(P.tide_scratch + i + k * cs);
                        const V3 tp = v3(ts[0], ts[1], ts[2]), tv = v3(ts[3], ts[4], ts[5]);
                        const double dist = ts[6], radvel = ts[7], orth_p = ts[8], diss_pm = ts[9];
                        const V3 tdl = v3(ts[10], ts[11], ts[12]);
                        const double factor2 = orth_p / dist;
                        const V3 wxr = cross(q.s, tp);
                        denergy = -((1.0 / dist * (diss_pm + factor2 * radvel)) * dot(tp, tv)
                                    + factor2 * ((wxr.x - tv.x) * tv.x + (wxr.y - tv.y) * tv.y + (wxr.z - tv.z) * tv.z))
                                  - dot(tdl, q.s);
                    }
                    record(b, plain(q.r), q.s, plain(q.v), K_M, K_R, K_RG2, denergy);
                    if (writer) record(0, plain(q.r0), q.s0, plain(q.v0), H_M, H_R, H_RG2, 0.);
                    if (PASSIVE) record(2, plain(q.r2), sl.get3(K2_S), plain(q.v2), K2_M, K2_R, K2_RG2, 0.);
                }
                if (snap) {
                    if (!first_snap) st.last_hist = __dadd_rn(st.last_hist, P.hist_period); else st.last_hist = 0.;
                    st.n_hist_new += 1;
                    if (st.hist_count < P.hist_capacity) st.hist_count += 1;
                    else st.warnings |= PB200_WARN_HISTORY_DROPPED;
                }
            }
        }
        const bool leaves_now = (step + 1 == step_end) || (__dadd_rn(__dadd_rn(st.t, P.dt), P.dt) > P.time_limit);
        const bool save_tides = (step + 1 == step_end) || (__dadd_rn(st.last_hist, P.hist_period) <= __dadd_rn(st.t, P.dt));

#pragma unroll 1
        for (int half = 0; half < 2; half++) {
            if (half == 1) {
                // =================== drift - kick - drift in the alternative coordinates ===================
                const sd m_s = sd(sl.get(K_M)), M_s = sd(sl.get(H_M)), mp_s = sd(sl.get(K_MP));
                const sd m1 = first ? m_s : mp_s, m2 = first ? mp_s : m_s;
                const srcp rM = sl.rcp(H_M, Y_MH), rT = sl.rcp(K_MTOT, Y_MTOT);
                S3 apos, avel, apos_o = zero3, avel_o = zero3;   // this planet's alternative coordinates and the partner's
                S3 spos, svel;                                     // centre of mass
                // ---- inertial -> alternative (whfast.rs:881-1023)
                if (JAC) {
                    const S3 ro_ = PASSIVE ? q.r2 : x3<PAIR>(q.r), vo_ = PASSIVE ? q.v2 : x3<PAIR>(q.v);
                    const S3 r1 = sel(first, q.r, ro_), r2 = sel(first, ro_, q.r), v1 = sel(first, q.v, vo_), v2 = sel(first, vo_, q.v);
                    const sd ei1 = sd(sl.get(J_EI1)), pme1 = sd(sl.get(J_PME1)), ei2 = sd(sl.get(J_EI2)), pme2 = sd(sl.get(J_PME2)), mi = sd(sl.get(J_MI));
                    S3 s = M_s * q.r0, sv = M_s * q.v0;
                    const S3 p1 = r1 - s * ei1, w1 = v1 - sv * ei1;
                    s = s * pme1 + m1 * p1; sv = sv * pme1 + m1 * w1;
                    const S3 p2 = r2 - s * ei2, w2 = v2 - sv * ei2;
                    s = s * pme2 + m2 * p2; sv = sv * pme2 + m2 * w2;
                    spos = s * mi; svel = sv * mi;
                    apos = sel(first, p1, p2); avel = sel(first, w1, w2);
                    apos_o = sel(first, p2, p1); avel_o = sel(first, w2, w1);
                } else {
                    // host first, then the others (whfast.rs:986-995)
                    const S3 mr = q.r * m_s, mv = q.v * m_s;
                    S3 sr = zero3 + q.r0 * M_s, sv = zero3 + q.v0 * M_s;
                    if (N == 3) {
                        const S3 mro = x3<PAIR>(mr), mvo = x3<PAIR>(mv);
                        sr = (sr + sel(first, mr, mro)) + sel(first, mro, mr);
                        sv = (sv + sel(first, mv, mvo)) + sel(first, mvo, mv);
                    } else { sr = sr + mr; sv = sv + mv; }
                    spos = sr / rT; svel = sv / rT;
                    apos = q.r - q.r0;
                    avel = q.v - svel;
                }
#pragma unroll 1
                for (int phase = 0; phase < 2; phase++) {
                    if (phase == 1) {
                        // ---- kick (whfast.rs:558-625)
                        if (JAC) {
                            // inertial_to_jacobi_acc (whfast.rs:935-963) + jacobi_interaction_step (:566-592)
                            const S3 ao = PASSIVE ? anew2 : x3<PAIR>(anew);
                            const S3 a1 = sel(first, anew, ao), a2 = sel(first, ao, anew);
                            const sd ei1 = sd(sl.get(J_EI1)), pme1 = sd(sl.get(J_PME1)), ei2 = sd(sl.get(J_EI2));
                            S3 sa = M_s * anew0;
                            const S3 c1 = a1 - sa * ei1;
                            sa = sa * pme1 + m1 * c1;
                            const S3 c2 = a2 - sa * ei2;
                            avel = avel + dt_s * sel(first, c1, c2);
                            // (the first non-host body has no extra term: whfast.rs:578)
                            const sd rj2i = sd(1.) / (apos.x * apos.x + apos.y * apos.y + apos.z * apos.z + sd(1e-12));
                            const sd rji = ssqrt(rj2i);
                            const sd rj3im = rji * rj2i * sd(kG) * sd(sl.get(J_ETAK));
                            const sd prefac = dt_s * rj3im;
                            if (!first) avel = avel + prefac * apos;
                            if (PASSIVE) {
                                // body 2 in the same thread: its Jacobi kick and the extra term (eta_2 = the total mass)
                                avel_o = avel_o + dt_s * c2;
                                const sd q2i = sd(1.) / (apos_o.x * apos_o.x + apos_o.y * apos_o.y + apos_o.z * apos_o.z + sd(1e-12));
                                const sd qi = ssqrt(q2i);
                                const sd q3im = qi * q2i * sd(kG) * sd(sl.get(K_MTOT));
                                avel_o = avel_o + (dt_s * q3im) * apos_o;
                            }
                        } else {
                            avel = avel + dt_s * anew;
                        }
                    }
                    // ---- jump (whfast.rs:495-556): before the Kepler drift in the second half, after it in the first
                    auto jump = [&]() {
                        S3 psum = zero3;
                        const S3 mine = m_s * avel;
                        if (N == 3) { const S3 other = x3<PAIR>(mine); psum = (psum + sel(first, mine, other)) + sel(first, other, mine); }
                        else psum = psum + mine;
                        apos = s3(apos.x + hdt_s * psum.x / rM, apos.y + hdt_s * psum.y / rM, apos.z + hdt_s * psum.z / rM);
                    };
                    if (!JAC && phase == 1) jump();
                    // (passive-planet build: a second pass through the same code drifts body 2; the coordinates swap places twice)
#pragma unroll 1
                    for (int kb = 0; kb < (PASSIVE ? 2 : 1); kb++) {
                        kepler_step(alive, apos, avel, sd(sl.get(kb ? K2_KMU : K_KMU)), hdt_s, st.tswarn, st.warnings);
                        if (PASSIVE) { const S3 tp = apos, tv = avel; apos = apos_o; avel = avel_o; apos_o = tp; avel_o = tv; }
                    }
                    if (PAIR) {
                        // the warning belongs to the system (whfast.rs:702-707)
                        unsigned int wv = st.warnings | (st.tswarn ? 0x80000000u : 0u);
                        wv |= __shfl_xor_sync(FULL, wv, 1);
                        st.warnings = wv & 0x7fffffffu; st.tswarn = (wv >> 31) != 0u;
                    }
                    spos = spos + hdt_s * svel;
                    if (!JAC && phase == 0) jump();
                    // ---- alternative -> inertial (whfast.rs:1026-1155)
                    S3 nr, nr0, nv = q.v, nv0 = q.v0, nr_o = zero3;
                    if (JAC) {
                        if (PAIR) { apos_o = x3<PAIR>(apos); if (phase == 1) avel_o = x3<PAIR>(avel); }
                        const S3 p1 = sel(first, apos, apos_o), p2 = sel(first, apos_o, apos), w1 = sel(first, avel, avel_o), w2 = sel(first, avel_o, avel);
                        const sd mtot = sd(sl.get(K_MTOT)), bei2 = sd(sl.get(J_MI)), et1 = sd(sl.get(J_ET1)), bei1 = sd(sl.get(J_BEI1)), et0 = sd(sl.get(J_ET0)),
                                 bmi = sd(sl.get(J_BMI));
                        S3 s = mtot * spos, sv = mtot * svel;
                        s = (s - m2 * p2) * bei2; sv = (sv - m2 * w2) * bei2;
                        const S3 r2 = p2 + s, v2 = w2 + sv;
                        s = s * et1; sv = sv * et1;
                        s = (s - m1 * p1) * bei1; sv = (sv - m1 * w1) * bei1;
                        const S3 r1 = p1 + s, v1 = w1 + sv;
                        s = s * et0; sv = sv * et0;
                        nr0 = s * bmi; nv0 = sv * bmi;
                        nr = sel(first, r1, r2); nr_o = sel(first, r2, r1);
                        if (phase == 1) nv = sel(first, v1, v2);
                        if (PASSIVE && alive) { q.r2 = r2; if (phase == 1) q.v2 = v2; }
                    } else {
                        const S3 term = (apos * m_s) / rT;
                        const S3 vterm = avel * sd(sl.get(K_BACKW));
                        S3 star_r = spos, star_v = svel;
                        if (N == 3) {
                            const S3 to = x3<PAIR>(term);
                            star_r = (star_r - sel(first, term, to)) - sel(first, to, term);
                            if (phase == 1) { const S3 vo_ = x3<PAIR>(vterm); star_v = (star_v - sel(first, vterm, vo_)) - sel(first, vo_, vterm); }
                            apos_o = x3<PAIR>(apos);
                            nr_o = apos_o + star_r;
                        } else {
                            star_r = star_r - term;
                            star_v = star_v - vterm;
                        }
                        nr0 = star_r; nr = apos + star_r;
                        if (phase == 1) { nv = avel + svel; nv0 = star_v; }
                    }
                    if (alive) { q.r = nr; q.r0 = nr0; if (phase == 1) { q.v = nv; q.v0 = nv0; } }
                    if (phase == 0) {
                        // ---- gravity with the Roche / collision / ejection checks (universe.rs:198-303)
                        int code = 0x7fffffff;
                        auto check = [&](int lo, int hi, double d2, int s_rr2, int s_rs2, bool host_pair) {
                            int fail = 0;
                            if (d2 <= sl.get(s_rr2)) fail = PB200_STATUS_ROCHE_DESTROYED;
                            if (!fail && d2 <= sl.get(s_rs2)) fail = PB200_STATUS_COLLISION;
                            if (!fail && host_pair && d2 > kMaxDistance2) fail = PB200_STATUS_EJECTED;
                            if (fail) { const int c = (lo << 8) | (hi << 4) | fail; code = c < code ? c : code; }
                        };
                        const S3 dh = q.r - q.r0;       // (host pair: the reference's d = r_host - r_planet, squares are the same)
                        const sd dh2 = dh.x * dh.x + dh.y * dh.y + dh.z * dh.z;
                        check(0, b, dh2.v, K_RR2H, K_RS2H, true);
                        S3 acc = zero3, acc0 = zero3;
                        if (PASSIVE) {
                            // the three pairs in one thread (Jacobi: the pair (host, body 1) is ignored, universe.rs:240-250)
                            const S3 d02 = q.r2 - q.r0, d12 = q.r2 - q.r;     // r_2 - r_0, r_2 - r_1
                            const sd d02s = d02.x * d02.x + d02.y * d02.y + d02.z * d02.z, d12s = d12.x * d12.x + d12.y * d12.y + d12.z * d12.z;
                            check(0, 2, d02s.v, K2_RR2H, K2_RS2H, true);
                            check(1, 2, d12s.v, K_RR2P, K_RS2P, false);
                            const sd di02 = ssqrt(d02s), di12 = ssqrt(d12s);
                            const sd g02 = sd(-kG) / (di02 * di02 * di02), g12 = sd(-kG) / (di12 * di12 * di12);
                            const sd pre02 = g02 * mp_s;           // host <- body 2, d = r_0 - r_2
                            acc0 = s3(acc0.x + pre02 * (-d02.x), acc0.y + pre02 * (-d02.y), acc0.z + pre02 * (-d02.z));
                            const sd pre12 = g12 * mp_s;           // body 1 <- body 2, d = r_1 - r_2
                            acc = s3(acc.x + pre12 * (-d12.x), acc.y + pre12 * (-d12.y), acc.z + pre12 * (-d12.z));
                            const sd pre20 = g02 * M_s, pre21 = g12 * m_s;   // body 2 <- host, then <- body 1
                            S3 a2 = zero3;
                            a2 = s3(a2.x + pre20 * d02.x, a2.y + pre20 * d02.y, a2.z + pre20 * d02.z);
                            a2 = s3(a2.x + pre21 * d12.x, a2.y + pre21 * d12.y, a2.z + pre21 * d12.z);
                            anew2 = a2;
                        } else if (N == 3) {
                            const S3 d = q.r - nr_o;     // r_b - r_partner
                            const sd d2 = d.x * d.x + d.y * d.y + d.z * d.z;
                            check(1, 2, d2.v, K_RR2P, K_RS2P, false);
                            const sd dist = ssqrt(d2);
                            const sd g = sd(-kG) / (dist * dist * dist);
                            if (JAC) {
                                // ignored: the pair (host, planet 1) (universe.rs:240-250)
                                const sd disth = ssqrt(dh2);
                                const sd gh = sd(-kG) / (disth * disth * disth);     // pair (host, this planet)
                                // planet 2's (host, 2) factor, needed by the host's acceleration in both lanes
                                const double gh_x = xd<PAIR>(gh.v);
                                const sd gh2 = sd(first ? gh_x : gh.v);
                                const S3 dh_2 = first ? (nr_o - q.r0) : dh;          // r_2 - r_0
                                const sd pre02 = gh2 * m2;                            // host's term from planet 2: -G / d^3 * m_2, d = r_0 - r_2
                                acc0 = s3(acc0.x + pre02 * (-dh_2.x), acc0.y + pre02 * (-dh_2.y), acc0.z + pre02 * (-dh_2.z));
                                if (first) {
                                    const sd pre = g * mp_s;
                                    acc = s3(acc.x + pre * d.x, acc.y + pre * d.y, acc.z + pre * d.z);
                                } else {
                                    const sd pre0 = gh * M_s, pre1 = g * mp_s;
                                    acc = s3(acc.x + pre0 * dh.x, acc.y + pre0 * dh.y, acc.z + pre0 * dh.z);
                                    acc = s3(acc.x + pre1 * d.x, acc.y + pre1 * d.y, acc.z + pre1 * d.z);
                                }
                            } else {
                                const sd pre = g * mp_s;
                                acc = s3(acc.x + pre * d.x, acc.y + pre * d.y, acc.z + pre * d.z);
                            }
                            const int o = __shfl_xor_sync(FULL, code, 1); code = o < code ? o : code;
                        }
                        anew = acc; anew0 = acc0;
                        const bool died = alive && code != 0x7fffffff;
                        if (alive && (leaves_now || died) && valid) {
                            const size_t i = (size_t)b * ns + sys;
                            P.acc[i] = anew.x.v; P.acc[i + cs] = anew.y.v; P.acc[i + 2 * cs] = anew.z.v;
                            if (writer) { P.acc[sys] = anew0.x.v; P.acc[sys + cs] = anew0.y.v; P.acc[sys + 2 * cs] = anew0.z.v; }
                            if (PASSIVE) { const size_t i2 = (size_t)2 * ns + sys; P.acc[i2] = anew2.x.v; P.acc[i2 + cs] = anew2.y.v; P.acc[i2 + 2 * cs] = anew2.z.v; }
                        }
                        if (died) {
                            st.status = code & 15; st.event_step = st.steps_done; alive = false;
                            store_state();
                        }
                    }
                }
            }
            // =================== implicit midpoint on v and L (whfast.rs:322-466) ===================
            {
                const bool evolution = half == 0;
                const bool save_t = half == 1 && save_tides;
                sl.set3(P_ORIG, plain(q.v)); sl.set3(P_ORIG + 3, q.L);
                sl.set3(H_ORIG, plain(q.v0)); sl.set3(H_ORIG + 3, q.L0);
                sl.set3(P_INCR, v3(0., 0., 0.)); sl.set3(P_INCR + 3, v3(0., 0., 0.));
                sl.set3(H_INCR, v3(0., 0., 0.)); sl.set3(H_INCR + 3, v3(0., 0., 0.));
                if (PASSIVE) {
                    sl.set3(P2_ORIG, plain(q.v2)); sl.set3(P2_ORIG + 3, sl.get3(K2_L));
                    sl.set3(P2_INCR, v3(0., 0., 0.)); sl.set3(P2_INCR + 3, v3(0., 0., 0.));
                }
                const S3 hr_s = q.r - q.r0;
                const V3 hr = plain(hr_s);
                const double inv_d = ARITH == 1 ? 0. : rsqrt(dot(hr, hr));
                const sd dist_s = ARITH ? ssqrt(hr_s.x * hr_s.x + hr_s.y * hr_s.y + hr_s.z * hr_s.z) : sd(1.);
                if (ARITH) { q.rs_s = sdot(hr_s, strict(q.s0)).v; q.rs_p = sdot(hr_s, strict(q.s)).v; }
                else { q.rs_s = dot(hr, q.s0); q.rs_p = dot(hr, q.s); }
                bool done = !alive;
                bool converged = false;
#pragma unroll 1
                for (int it = 0; it < 10; it++) {
                    if (!__any_sync(FULL, !done)) break;
                    if (has<FLAGS>(P, FLAG_EVO) && evolution && it == 0) evolve_all(st.t, alive);
                    // calculate_spin of body 2 (every evaluation refreshes every particle's spin, common.rs:3-15): L_2 never
                    // changes, so once per midpoint, after the inertia may have evolved
                    if (PASSIVE && it == 0 && !done) sl.set3(K2_S, plain(strict(sl.get3(K2_L)) / sd(sl.get(K2_I))));
                    const S3 hv_s = q.v - q.v0;
                    const bool save_now = save_t && !done;
                    const bool exact_now = ARITH == 1 || (ARITH == 2 && it >= 2);
                    S3 a_p, dl_p, a_h, dl_h;    // (plain doubles inside `sd` when they come from the fast forces)
                    if (exact_now) {
                        small_effects_exact<N, FLAGS, PASSIVE>(P, sl, ro, valid, b, sys, q, hr_s, dist_s, hv_s, a_p, dl_p, a_h, dl_h, save_now);
                    } else {
                        V3 fa_p, fdl_p, fa_h, fdl_h;
                        small_effects_fast<FLAGS, ARITH == 2>(P, sl, valid, b, sys, q, hr, inv_d, plain(hv_s), fa_p, fdl_p, fa_h, fdl_h, save_now);
                        if (PAIR) {
                            const V3 oa = x3<PAIR>(fa_h), od = x3<PAIR>(fdl_h);
                            fa_h = sel(first, fa_h, oa) + sel(first, oa, fa_h);
                            fdl_h = sel(first, fdl_h, od) + sel(first, od, fdl_h);
                        }
                        a_p = strict(fa_p); dl_p = strict(fdl_p); a_h = strict(fa_h); dl_h = strict(fdl_h);
                    }
                    // final = orig + (dt * a - err)   (whfast.rs:353-378)
                    const S3 vo = strict(sl.get3(P_ORIG)), Lo = strict(sl.get3(P_ORIG + 3)), vo0 = strict(sl.get3(H_ORIG)), Lo0 = strict(sl.get3(H_ORIG + 3));
                    const S3 ev = strict(sl.get3(P_ERR)), el = strict(sl.get3(P_ERR + 3)), ev0 = strict(sl.get3(H_ERR)), el0 = strict(sl.get3(H_ERR + 3));
                    const S3 ndv = s3(hdt_s * a_p.x - ev.x, hdt_s * a_p.y - ev.y, hdt_s * a_p.z - ev.z);
                    const S3 ndl = s3(hdt_s * dl_p.x - el.x, hdt_s * dl_p.y - el.y, hdt_s * dl_p.z - el.z);
                    const S3 ndv0 = s3(hdt_s * a_h.x - ev0.x, hdt_s * a_h.y - ev0.y, hdt_s * a_h.z - ev0.z);
                    const S3 ndl0 = s3(hdt_s * dl_h.x - el0.x, hdt_s * dl_h.y - el0.y, hdt_s * dl_h.z - el0.z);
                    const S3 vf = vo + ndv, Lf = Lo + ndl, vf0 = vo0 + ndv0, Lf0 = Lo0 + ndl0;
                    // passive-planet build: body 2 feels no additional effect — zero acceleration and torque through the same update
                    S3 vo2 = zero3, Lo2 = zero3, ndv2 = zero3, ndl2 = zero3, vf2 = zero3, Lf2 = zero3;
                    if (PASSIVE) {
                        vo2 = strict(sl.get3(P2_ORIG)); Lo2 = strict(sl.get3(P2_ORIG + 3));
                        const S3 ev2 = strict(sl.get3(P2_ERR)), el2 = strict(sl.get3(P2_ERR + 3));
                        ndv2 = s3(hdt_s * zero - ev2.x, hdt_s * zero - ev2.y, hdt_s * zero - ev2.z);
                        ndl2 = s3(hdt_s * zero - el2.x, hdt_s * zero - el2.y, hdt_s * zero - el2.z);
                        vf2 = vo2 + ndv2; Lf2 = Lo2 + ndl2;
                    }
                    bool conv_now = false;
                    if (it >= 2) {
                        // whfast.rs:424-451: sum(delta^2) / sum(total^2) < eps^2 decided as sum(delta_i^2 - eps^2 total_i^2) < 0
                        const S3 vf_old = vo + strict(sl.get3(P_INCR)), Lf_old = Lo + strict(sl.get3(P_INCR + 3));
                        const S3 vf0_old = vo0 + strict(sl.get3(H_INCR)), Lf0_old = Lo0 + strict(sl.get3(H_INCR + 3));
                        const V3 ddv = plain(vf - vf_old), ddl = plain(Lf - Lf_old), ddv0 = plain(vf0 - vf0_old), ddl0 = plain(Lf0 - Lf0_old);
                        const V3 vfp = plain(vf), Lfp = plain(Lf), vf0p = plain(vf0), Lf0p = plain(Lf0);
                        double c_v = valid ? dot(ddv, ddv) - kEps2 * dot(vfp, vfp) : 0.;
                        double c_l = valid ? dot(ddl, ddl) - kEps2 * dot(Lfp, Lfp) : 0.;
                        if (PAIR) { c_v += xd<PAIR>(c_v); c_l += xd<PAIR>(c_l); }
                        if (PASSIVE) {
                            const S3 vf2_old = vo2 + strict(sl.get3(P2_INCR)), Lf2_old = Lo2 + strict(sl.get3(P2_INCR + 3));
                            const V3 ddv2 = plain(vf2 - vf2_old), ddl2 = plain(Lf2 - Lf2_old), vf2p = plain(vf2), Lf2p = plain(Lf2);
                            c_v += dot(ddv2, ddv2) - kEps2 * dot(vf2p, vf2p);
                            c_l += dot(ddl2, ddl2) - kEps2 * dot(Lf2p, Lf2p);
                        }
                        c_v += dot(ddv0, ddv0) - kEps2 * dot(vf0p, vf0p);
                        c_l += dot(ddl0, ddl0) - kEps2 * dot(Lf0p, Lf0p);
                        conv_now = c_v < 0. && c_l < 0.;
                    }
                    if (!done) {
                        if (it > 0) {
                            sl.set3(P_INCR, plain(ndv)); sl.set3(P_INCR + 3, plain(ndl));
                            sl.set3(H_INCR, plain(ndv0)); sl.set3(H_INCR + 3, plain(ndl0));
                            if (PASSIVE) { sl.set3(P2_INCR, plain(ndv2)); sl.set3(P2_INCR + 3, plain(ndl2)); }
                        }
                        if (conv_now) { done = true; converged = true; }
                        else {
                            // average (whfast.rs:453-466)
                            const sd h = sd(0.5);
                            q.v = s3(h * (vo.x + vf.x), h * (vo.y + vf.y), h * (vo.z + vf.z));
                            q.v0 = s3(h * (vo0.x + vf0.x), h * (vo0.y + vf0.y), h * (vo0.z + vf0.z));
                            q.L = plain(s3(h * (Lo.x + Lf.x), h * (Lo.y + Lf.y), h * (Lo.z + Lf.z)));
                            q.L0 = plain(s3(h * (Lo0.x + Lf0.x), h * (Lo0.y + Lf0.y), h * (Lo0.z + Lf0.z)));
                            if (PASSIVE) {
                                q.v2 = s3(h * (vo2.x + vf2.x), h * (vo2.y + vf2.y), h * (vo2.z + vf2.z));
                                sl.set3(K2_L, plain(s3(h * (Lo2.x + Lf2.x), h * (Lo2.y + Lf2.y), h * (Lo2.z + Lf2.z))));
                            }
                        }
                    }
                }
                if (alive) {
                    if (!converged) st.warnings |= PB200_WARN_MIDPOINT_NOT_CONVERGED;
                    auto commit = [&](int s_orig, int s_incr, int s_err, S3& v_, V3& L_) {
                        const S3 vo = strict(sl.get3(s_orig)), Lo = strict(sl.get3(s_orig + 3)), dv = strict(sl.get3(s_incr)), dl = strict(sl.get3(s_incr + 3));
                        v_ = vo + dv;
                        const S3 Ln = Lo + dl;
                        L_ = plain(Ln);
                        sl.set3(s_err, plain((v_ - vo) - dv));
                        sl.set3(s_err + 3, plain((Ln - Lo) - dl));
                    };
                    commit(P_ORIG, P_INCR, P_ERR, q.v, q.L);
                    commit(H_ORIG, H_INCR, H_ERR, q.v0, q.L0);
                    if (PASSIVE) { V3 L2n; commit(P2_ORIG, P2_INCR, P2_ERR, q.v2, L2n); sl.set3(K2_L, L2n); }
                }
            }
        }
        if (alive) {
            st.t = __dadd_rn(st.t, P.dt);
            st.steps_done += 1;
            if (__dadd_rn(st.t, P.dt) > P.time_limit) {
                st.status = PB200_STATUS_COMPLETED; st.event_step = st.steps_done; alive = false;
                store_state();
            }
        }
    }
    if (alive) store_state();
    if (P.n_pieces > 1) {
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) { volatile unsigned int* flag = P.sched + 1 + group; *flag = piece + 1; }
    }
}

}  // namespace PB_NS
