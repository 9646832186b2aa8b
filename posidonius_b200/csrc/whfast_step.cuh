// whfast_step.cuh — the persistent-state step kernel: WHFast::iterate (integrator/whfast.rs:235-305) for
// n_steps steps of every system of the ensemble, state in registers (see whfast_kernel.cuh for the mapping).
//
// The core below is a strict-arithmetic transcription (strict.cuh) of the reference's operation order; sums over
// bodies that the reference accumulates serially are accumulated in the same order by walking the group with
// shuffles (every lane carries the running sum, so all lanes hold bit-identical copies).
#include "gr_variants.cuh"
#include "exact_effects.cuh"
#include "strict_gr_variants.cuh"

namespace PB_NS {
using namespace pb200;

#ifndef PB_BLOCK
#define PB_BLOCK 64
#endif
#define PB_HIST_FIELDS 17  // time, pos3, spin3, vel3, mass, radius, rg2, love_number, sigma, denergy_dt, lag_angle
#define PB_TIDE_SCRATCH 13

struct SysState {
    double t, last_hist;
    unsigned int steps_done, n_hist_new;   // relative to the launch start
    unsigned int event_step;               // steps_done when the status changed
    int status;
    unsigned int warnings;
    int hist_count;
    bool tswarn;
};

__device__ __forceinline__ void store_lane(const KParams& P, const Roles& ro, const Cold& cold, size_t sys, int b, const Lane& q,
                                           const SysState& st) {
    if (!ro.valid) return;
    const size_t ns = (size_t)P.n_sys;
    const size_t i = (size_t)b * ns + sys;
    const size_t cs = (size_t)PB_N(P) * ns;
    P.pos[i] = q.r.x.v; P.pos[i + cs] = q.r.y.v; P.pos[i + 2 * cs] = q.r.z.v;
    P.vel[i] = q.v.x.v; P.vel[i + cs] = q.v.y.v; P.vel[i + 2 * cs] = q.v.z.v;
    // (the Newtonian acceleration is written by the step that computed it: store_acc)
    P.L[i] = q.L.x; P.L[i + cs] = q.L.y; P.L[i + 2 * cs] = q.L.z;
    P.spin[i] = q.s.x; P.spin[i + cs] = q.s.y; P.spin[i + 2 * cs] = q.s.z;
    V3 ev, el;
    cold.get6(R_ERR, ev, el);
    P.verr[i] = ev.x; P.verr[i + cs] = ev.y; P.verr[i + 2 * cs] = ev.z;
    P.lerr[i] = el.x; P.lerr[i + cs] = el.y; P.lerr[i + 2 * cs] = el.z;
    // radius / radius of gyration / moment of inertia are written when they evolve (evolve_lane)
    if (b == 0) {
        P.t[sys] = st.t; P.last_hist[sys] = st.last_hist;
        unsigned long long it0 = ldm(P.iteration + sys);
        P.iteration[sys] = it0 + st.steps_done;
        P.n_hist[sys] = ldm(P.n_hist + sys) + st.n_hist_new;
        if (st.status != PB200_STATUS_OK) P.event_iteration[sys] = it0 + st.event_step;
        P.tswarn[sys] = st.tswarn ? 1ull : 0ull;
        P.status[sys] = st.status; P.warnings[sys] = st.warnings; P.hist_count[sys] = st.hist_count;
    }
}

// effects/evolution.rs:516-546 for this lane's body (radius, radius of gyration -> moment of inertia). Cold path:
// the evolving quantities live in global memory (rg2) and in the cold slots (R, I). Returns true when something changed.
__device__ __forceinline__ bool evolve_lane(const KParams& P, const Roles& ro, const Cold& cold, int b, size_t sys, double t, bool commit) {
    if (!ro.valid || !commit) return false;
    int ti = P.evo_table[b];
    if (ti < 0) return false;
    const DevTable& T = P.tables[ti];
    const size_t idx = (size_t)b * (size_t)P.n_sys + sys;
    const double R = cold.get(K_R), rg2 = ldm(P.rg2 + idx);
    int i = table_upper(T.time, T.n_rows, t);
    double nr = T.interp_radius ? table_interp(T.time, T.radius, T.n_rows, i, t) : R;
    double ng = ((P.evo_rg2 >> b) & 1u) ? table_interp(T.time, T.rg2, T.n_rows, i, t) : rg2;
    if (nr != R || ng != rg2) {
        double I = (sd(cold.get(K_M)) * sd(ng) * (sd(nr) * sd(nr))).v;
        cold.set(K_R, nr); cold.set(K_I, I);
        P.radius[idx] = nr; P.rg2[idx] = ng; P.moi[idx] = I;
        return true;
    }
    return false;
}

// Running sum over the NON-HOST bodies in index order, as the reference's serial loops accumulate
// (`particles_left.chain(particles_right)`). Every lane has left its own term in the exchange triple `slot`
// (cold.set3 + __syncwarp by the caller); all lanes get the same bits.
__device__ __forceinline__ S3 ordered_sum_others(const Cold& cold, S3 init, int slot, int n, int host) {
    S3 acc = init;
    for (int k = 0; k < n; k++) {
        if (k == host) continue;
        acc = acc + strict(cold.getk3(k, slot));
    }
    return acc;
}
__device__ __forceinline__ S3 ordered_diff_others(const Cold& cold, S3 init, int slot, int n, int host) {
    S3 acc = init;
    for (int k = 0; k < n; k++) {
        if (k == host) continue;
        acc = acc - strict(cold.getk3(k, slot));
    }
    return acc;
}

// Implicit midpoint on v and L (whfast.rs:322-466) around Universe::calculate_additional_effects.
// Registers across the evaluation: v, L, spin, heliocentric position and 1/r. Originals, increments and Kahan
// residuals sit in the cold slots and are touched once per iteration.
//
// ARITH selects the arithmetic of the perturbation forces:
//   0 (PB200_ARITH_FAST)    every evaluation with the fast forces (forces_fast.cuh);
//   1 (PB200_ARITH_STRICT)  every evaluation with the exact forces (exact_effects.cuh): bit-identical to the reference's arithmetic;
//   2 (PB200_ARITH_HYBRID)  iterations 0 and 1 with the fast forces, every later one — among them the one whose increments
//     are COMMITTED (the reference needs at least three iterations, whfast.rs:386) — with the exact forces. The committed
//     increment, hence the new v, L and their Kahan residuals, carry the reference's roundings whenever the two uncommitted
//     iterates rounded to the same doubles as the reference's, which they do except with probability ~ |dv| / |v| per
//     component (the iterates enter the exact evaluation only through v_orig + dv / 2, and a relative error of 1e-15 in a
//     dv of relative size 1e-8 .. 1e-6 rarely moves that sum across a rounding boundary; the angular momenta move faster
//     — dL / L up to 1e-5 per step — so the spin entering the exact evaluation is an ulp off now and then, which shows in
//     the Kahan residuals first and in r, v only much later: profiles/r2_hybrid_decay.txt).
template <int COORD, int GR, int ARITH>
__device__ __forceinline__ void midpoint(const KParams& P, const Roles& ro, const Cold& cold, int gb, int hl, int b, bool alive, Lane& q,
                                         double t, bool evolution, unsigned int& warnings, bool save_tides, size_t sys) {
    const int W = PB_W(P);
    const sd dt = sd(P.half_dt);
    // positions do not change inside the midpoint: heliocentric position and 1/r once (universe.rs:318-351)
    __syncwarp();   // the core's last exchange reads are done before these slots are rewritten
    cold.set3(S_RX, plain(q.r));
    cold.set6(R_ORIG, plain(q.v), q.L);
    cold.set6(R_INCR, v3(0., 0., 0.), v3(0., 0., 0.));
    cold.set3(M_0, plain(q.v)); cold.set3(E_S, q.s);   // the host's current v and last spin for the group
    __syncwarp();
    const S3 rh_s = strict(cold.getk3(PB_HOST(P), S_RX));
    // idle lanes (host slot, padding) get a unit dummy so that rsqrt/div stay on their fast paths for the whole warp
    const S3 hr_s = ro.planet ? q.r - rh_s : s3(sd(1.), sd(0.), sd(0.));
    const V3 hr = plain(hr_s);
    const double inv_d = ARITH == 1 ? 0. : rsqrt(dot(hr, hr));
    const sd dist_s = ARITH ? ssqrt(hr_s.x * hr_s.x + hr_s.y * hr_s.y + hr_s.z * hr_s.z) : sd(1.);   // universe.rs:328-330
    // Q3: previous spins, new position
    if (ARITH) { q.rs_s = sdot(hr_s, strict(cold.getk3(PB_HOST(P), E_S))).v; q.rs_p = sdot(hr_s, strict(q.s)).v; }
    else { q.rs_s = dot(hr, cold.getk3(PB_HOST(P), E_S)); q.rs_p = dot(hr, q.s); }
    // the first evaluation rewrites E_S with the fresh spins: every lane's read above comes first (the vote at the top of the
    // loop already keeps the lanes together; the barrier also orders the shared-memory accesses — compute-sanitizer racecheck)
    __syncwarp();
    bool done = !alive;  // group-uniform
    bool converged = false;
#pragma unroll 1
    for (int it = 0; it < 10; it++) {
#if PB_PHASE_BARRIER
        // block-uniform exit: the barrier keeps the warps of the CTA inside the same stretch of code (instruction cache)
        if (!__syncthreads_or(!done)) break;
#else
        if (!__any_sync(FULL, !done)) break;
#endif
        if (evolution && it == 0 && (PB_FLAGS(P) & FLAG_EVO)) {
            // once per step in evolving configurations: re-derive the constants unconditionally (warp-uniform control flow
            // around the shuffles inside make_consts)
            (void)evolve_lane(P, ro, cold, b, sys, t, alive);
            __syncwarp();
            make_consts(P, ro, cold, hl, b, sys);
        }
        const S3 vh_s = strict(cold.getk3(PB_HOST(P), M_0));
        const S3 hv_s = q.v - vh_s;
        V3 a, dldt;
        // the tidal internals of a step's last evaluation are kept for the next snapshot's denergy_dt (`!done`: a converged
        // system's later evaluations are discarded)
        const bool save_now = save_tides && !done;
        const bool evolve_now = evolution && it == 0;
        const bool exact_now = ARITH == 1 || (ARITH == 2 && it >= 2);   // warp-uniform
        if (exact_now) {
            // the idle lanes divide by |v|: a unit dummy
            const S3 hv_x = ro.planet ? hv_s : s3(sd(0.), sd(1.), sd(0.));
            additional_effects_exact<GR>(P, ro, cold, hl, b, sys, t, evolve_now, q, hr_s, dist_s, hv_x, a, dldt, save_now);
        } else {
            // the fast forces hold no division, and every term of a non-orbiting lane is multiplied by a zero constant
            // (make_consts), so the host / padding lanes need no dummy velocity (six selects less per evaluation)
            additional_effects<GR, ARITH == 2>(P, ro, cold, hl, b, sys, t, evolve_now, q, hr, inv_d, plain(hv_s), a, dldt, save_now);
        }
#if !PB_FIXED_N
        if (GR == PB200_GR_ANDERSON1975 || GR == PB200_GR_NEWHALL1983) {
            if (PB_FLAGS(P) & FLAG_GR) {
                V3 ag;
                Lane qq = q;
                qq.r = strict(cold.get3(S_RX));
                V3 acc_newton = cold.get3(S_AX);
                if (exact_now) {
                    if (GR == PB200_GR_ANDERSON1975) gr_anderson1975_strict(P, ro, cold, b, sys, qq, hr_s, strict(acc_newton), COORD == PB200_COORD_JACOBI, ag);
                    else gr_newhall1983_strict(P, ro, cold, b, qq, hr_s, strict(acc_newton), COORD == PB200_COORD_JACOBI, ag);
                    a = plain(strict(a) + strict(ag));
                } else {
                    if (GR == PB200_GR_ANDERSON1975) gr_anderson1975(P, ro, cold, gb, hl, b, sys, qq, hr, acc_newton, COORD == PB200_COORD_JACOBI, ag);
                    else gr_newhall1983(P, ro, cold, gb, hl, b, qq, hr, acc_newton, COORD == PB200_COORD_JACOBI, ag);
                    a = a + ag;
                }
            }
        }
#endif
        // final = orig + (dt * a - err)   (whfast.rs:353-378), with the previous final for the convergence test
        V3 vo_p, Lo, ev, el;
        cold.get6(R_ORIG, vo_p, Lo); cold.get6(R_ERR, ev, el);
        const S3 vo = strict(vo_p);
        S3 ndv = s3(dt * sd(a.x) - sd(ev.x), dt * sd(a.y) - sd(ev.y), dt * sd(a.z) - sd(ev.z));
        V3 ndl = v3((dt * sd(dldt.x) - sd(el.x)).v, (dt * sd(dldt.y) - sd(el.y)).v, (dt * sd(dldt.z) - sd(el.z)).v);
        S3 vf = vo + ndv;
        V3 Lf = PB_SPIN(P) ? v3(__dadd_rn(Lo.x, ndl.x), __dadd_rn(Lo.y, ndl.y), __dadd_rn(Lo.z, ndl.z)) : Lo;
        bool conv_now = false;
        if (it >= 2) {
            // whfast.rs:424-451 (sums over bodies by butterfly: only the branch decision depends on them); the previous
            // iterate's final values are rebuilt from the stored increments
            V3 dv_old, dl_old;
            cold.get6(R_INCR, dv_old, dl_old);
            const S3 vf_old = vo + strict(dv_old);
            const V3 Lf_old = v3(__dadd_rn(Lo.x, dl_old.x), __dadd_rn(Lo.y, dl_old.y), __dadd_rn(Lo.z, dl_old.z));
            V3 ddv = plain(vf - vf_old), ddl = Lf - Lf_old, vfp = plain(vf);
            // delta/total < eps^2 decided as sum(delta_i - eps^2 total_i) < 0: one group sum per test, no division; NaN
            // compares false either way. The two sides differ by factors of order one whenever an iterate moved by an ulp
            // (delta_i = ulp^2 <= eps^2 v_i^2), so the rounding of the merged sum cannot flip the decision.
            double c_v = ro.valid ? dot(ddv, ddv) - kEps2 * dot(vfp, vfp) : 0.;
            bool okv = group_sum(c_v, W) < 0.;
            bool okl = true;
            if (PB_SPIN(P)) {
                double c_l = ro.valid ? dot(ddl, ddl) - kEps2 * dot(Lf, Lf) : 0.;
                okl = group_sum(c_l, W) < 0.;
            }
            conv_now = okv && okl;
        }
        if (!done) {
            // the increments of iteration 0 are never read (the convergence test starts at it = 2 with those of it = 1, the
            // final update uses the last ones, and a live system always runs at least 3 iterations)
            if (it > 0) cold.set6(R_INCR, plain(ndv), ndl);
            if (conv_now) { done = true; converged = true; }
            else {
                // average (whfast.rs:453-466)
                q.v = s3(sd(0.5) * (vo.x + vf.x), sd(0.5) * (vo.y + vf.y), sd(0.5) * (vo.z + vf.z));
                if (PB_SPIN(P)) q.L = v3(__dmul_rn(0.5, __dadd_rn(Lo.x, Lf.x)), __dmul_rn(0.5, __dadd_rn(Lo.y, Lf.y)), __dmul_rn(0.5, __dadd_rn(Lo.z, Lf.z)));
            }
        }
        // publish the (possibly averaged) velocity for the next evaluation; the host has collected its totals by now
        __syncwarp();
        cold.set3(M_0, plain(q.v));
        __syncwarp();
    }
    q.r = strict(cold.get3(S_RX));
    if (alive) {
        if (!converged) warnings |= PB200_WARN_MIDPOINT_NOT_CONVERGED;
        V3 vo_p, Lo, dv_p, dl, ev_new, el_new;
        cold.get6(R_ORIG, vo_p, Lo); cold.get6(R_INCR, dv_p, dl);
        const S3 vo = strict(vo_p), dv = strict(dv_p);
        q.v = vo + dv;
        ev_new = plain((q.v - vo) - dv);
        if (PB_SPIN(P)) {
            q.L = v3(__dadd_rn(Lo.x, dl.x), __dadd_rn(Lo.y, dl.y), __dadd_rn(Lo.z, dl.z));
            el_new = v3(__dsub_rn(__dsub_rn(q.L.x, Lo.x), dl.x), __dsub_rn(__dsub_rn(q.L.y, Lo.y), dl.y), __dsub_rn(__dsub_rn(q.L.z, Lo.z), dl.z));
        } else {
            V3 unused;
            cold.get6(R_ERR, unused, el_new);   // the L residuals stay as they are (they share a cell with the v residuals)
        }
        cold.set6(R_ERR, ev_new, el_new);
    }
}

// particles/universe.rs:198-303 — Newtonian gravity with the WHFast ignore rules and the
// Roche / collision / ejection checks (panic! in the reference, status word here). Strict arithmetic.
template <int COORD>
__device__ __forceinline__ S3 gravity(const KParams& P, const Roles& ro, const Cold& cold, int gb, int b, size_t sys, const Lane& q, int& fail) {
    S3 acc = s3(sd(0.), sd(0.), sd(0.));
    const double q_R = cold.get(K_R);
    // upper bound of this body's squared Roche radii (K_ROCHE2, set when the CTA starts): the table itself stays in global
    // memory and is read only when a pair comes that close — no global load in the step loop, so the L1 that the
    // shared-memory carve-out leaves (28 KB at 12 warps/SM) does not matter
    const double roche_max2 = cold.get(K_ROCHE2);
    const int n = PB_N(P);
    const int first_other = PB_HOST(P) == 0 ? 1 : 0;
    fail = 0;
#if PB_FIXED_N
#pragma unroll
#else
#pragma unroll 1
#endif
    for (int j = 0; j < n; j++) {
        if (j == b || !ro.valid) continue;
        const S3 rj = strict(cold.getk3(j, E_R));   // published by the caller
        const sd mj = sd(cold.getk(j, K_M));
        const double Rj = cold.getk(j, K_R);
        S3 d = q.r - rj;
        sd d2 = d.x * d.x + d.y * d.y + d.z * d.z;
        if (b < j) {
            double rs = __dadd_rn(q_R, Rj);
            if (d2.v <= roche_max2) {
                const double rr = __ldg(P.roche + ((size_t)(b * n + j)) * (size_t)P.n_sys + sys);
                if (d2.v <= __dmul_rn(rr, rr)) { if (!fail) fail = PB200_STATUS_ROCHE_DESTROYED; }
            }
            if (d2.v <= __dmul_rn(rs, rs)) { if (!fail) fail = PB200_STATUS_COLLISION; }
            if (b == PB_HOST(P) && d2.v > kMaxDistance2) { if (!fail) fail = PB200_STATUS_EJECTED; }
        }
        bool skip;
        if (COORD == PB200_COORD_JACOBI) skip = (b == PB_HOST(P) && j == first_other) || (j == PB_HOST(P) && b == first_other);
        else skip = (b == PB_HOST(P) || j == PB_HOST(P));
        if (skip) continue;
        sd dist = ssqrt(d2);
        sd pre = sd(-kG) / (dist * dist * dist) * mj;
        acc.x = acc.x + pre * d.x; acc.y = acc.y + pre * d.y; acc.z = acc.z + pre * d.z;
    }
    return acc;
}

#if PB_FIXED_N == 8
// The same for exactly 8 bodies, host 0, democratic heliocentric coordinates: every unordered pair is evaluated ONCE.
// -G / |d|^3 is symmetric in the pair bit for bit (the components of d change sign exactly, their squares and the order of
// the sum do not), so planet b computes it for its next three neighbours on the ring of the seven planets, leaves it where
// both members of the pair will look for it (own column and the partner's, slot = rank of the other body among the
// partners in index order), and then accumulates its six terms in index order like the reference's inner loop — with the
// same roundings as gravity() above. The pair's Roche / collision checks are done by the lane that computes the pair; the
// host pairs (ignored by the acceleration, universe.rs:251-255) are checked by the planet. `code` packs (lower index,
// higher index, status): the group-wide minimum is the first failing pair of the reference's loop order.
__device__ __forceinline__ S3 gravity_n8_dh(const KParams& P, const Roles& ro, const Cold& cold, int b, size_t sys, const Lane& q, int& code) {
    const double q_R = cold.get(K_R);
    const double roche_max2 = cold.get(K_ROCHE2);
    const bool pl = ro.planet;
    const int bb = pl ? b : 1;   // host / padding lanes walk planet 1's pattern on their own (never published) values
    code = 0x7fffffff;
    auto check = [&](int lo, int hi, double d2, double Rj, bool host_pair) {
        int fail = 0;
        if (d2 <= roche_max2) {
            const double rr = __ldg(P.roche + ((size_t)(lo * 8 + hi)) * (size_t)P.n_sys + sys);
            if (d2 <= __dmul_rn(rr, rr)) fail = PB200_STATUS_ROCHE_DESTROYED;
        }
        const double rs = __dadd_rn(q_R, Rj);
        if (!fail && d2 <= __dmul_rn(rs, rs)) fail = PB200_STATUS_COLLISION;
        if (!fail && host_pair && d2 > kMaxDistance2) fail = PB200_STATUS_EJECTED;
        if (fail && pl) { const int c = (lo << 8) | (hi << 4) | fail; code = c < code ? c : code; }
    };
    {
        const S3 d = q.r - strict(cold.getk3(0, E_R));
        const sd d2 = d.x * d.x + d.y * d.y + d.z * d.z;
        check(0, bb, d2.v, cold.getk(0, K_R), true);
    }
#pragma unroll
    for (int k = 1; k <= 3; k++) {
        int j = bb + k; j = j > 7 ? j - 7 : j;
        const S3 d = q.r - strict(cold.getk3(j, E_R));
        const sd d2 = d.x * d.x + d.y * d.y + d.z * d.z;
        check(bb < j ? bb : j, bb < j ? j : bb, d2.v, cold.getk(j, K_R), false);
        const sd dist = ssqrt(d2);
        const sd g = sd(-kG) / (dist * dist * dist);
        if (pl) {
            // rank of the partner among this body's partners (index order, 0-based), and of this body among the partner's
            cold.set(E_A + (j > bb ? j - 2 : j - 1), g.v);
            cold.grp[j + (E_A + (bb > j ? bb - 2 : bb - 1)) * PB_BLOCK] = g.v;
        }
    }
    __syncwarp();
    S3 acc = s3(sd(0.), sd(0.), sd(0.));
#pragma unroll
    for (int k = 1; k <= 6; k++) {
        const int j = k + (k >= bb ? 1 : 0);
        const sd pre = sd(cold.get(E_A + k - 1)) * sd(cold.getk(j, K_M));
        const S3 d = q.r - strict(cold.getk3(j, E_R));
        acc.x = acc.x + pre * d.x; acc.y = acc.y + pre * d.y; acc.z = acc.z + pre * d.z;
    }
    if (!pl) acc = s3(sd(0.), sd(0.), sd(0.));
    return acc;
}
#endif

#ifndef PB_STEP_BARRIER
#define PB_STEP_BARRIER 0   // a block-wide barrier per step paid +2 % at 3 250 FP64 instructions per warp-step, costs 3 % at 2 700 (profiles/r1_variants.md)
#endif
#ifndef PB_PHASE_BARRIER
#define PB_PHASE_BARRIER 0   // block-wide barriers before every force evaluation, Kepler drift and gravity evaluation (instruction cache)
#endif
#ifndef PB_MIN_BLOCKS
#define PB_MIN_BLOCKS 5   // code-generation hint only: <= 168 registers/thread, no spills; six CTAs (12 warps) are resident (profiles/r1_variants.md)
#endif
#ifdef PB_MAXNREG
#define PB_KERNEL_ATTR __launch_bounds__(PB_BLOCK) __maxnreg__(PB_MAXNREG)
#else
#define PB_KERNEL_ATTR __launch_bounds__(PB_BLOCK, PB_MIN_BLOCKS)
#endif

template <int COORD, int GR, int ARITH>
__global__ void PB_KERNEL_ATTR whfast_steps_kernel(const __grid_constant__ KParams P, unsigned long long n_steps) {
    const int W = PB_W(P);
    const int n = PB_N(P);
    // ---- time slicing: which piece of which group of systems this CTA runs (KParams::sched). Tickets are handed out in
    // the order CTAs actually start, so the piece a CTA waits for has always started already: no deadlock.
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
    const int lane = threadIdx.x & 31;
    const int b = (int)(gtid & (size_t)(W - 1));
    const size_t sys = gtid >> PB_SHIFT(P);
    const int gb = lane & ~(W - 1);
    const int hl = gb + PB_HOST(P);
    const bool sys_ok = sys < (size_t)P.n_sys;
    Roles ro;
    ro.valid = sys_ok && b < n;
    ro.host = ro.valid && b == PB_HOST(P);
    ro.planet = ro.valid && !ro.host;
    ro.t_on = ro.planet && ((P.tides_orbiting >> b) & 1u);
    ro.f_on = ro.planet && ((P.flat_orbiting >> b) & 1u);
    ro.g_on = ro.planet && ((P.gr_orbiting >> b) & 1u);
    Cold cold;
    cold.base = pb_smem + threadIdx.x;
    cold.grp = cold.base - b;
    // a warp's pair cells stay inside the warp's own 32 columns (the warps of a CTA run unsynchronised): lanes 0-15 use the
    // first slot of a cell pair, lanes 16-31 the second, 16 bytes per lane
    cold.pair = pb_smem + (((threadIdx.x & 31u) >> 4) * PB_BLOCK + (threadIdx.x & ~31u) + 2u * (threadIdx.x & 15u));

    Lane q;
    SysState st;
    {
        const size_t ns = (size_t)P.n_sys;
        const size_t i = (size_t)b * ns + sys, cs = (size_t)n * ns;
        if (ro.valid) {
            auto ld3 = [&](const double* a) { return v3(ldm(a + i), ldm(a + i + cs), ldm(a + i + 2 * cs)); };
            q.r = strict(ld3(P.pos)); q.v = strict(ld3(P.vel));
            q.L = ld3(P.L); q.s = ld3(P.spin);
            cold.set6(R_ERR, ld3(P.verr), ld3(P.lerr));
#if !PB_FIXED_N
            cold.set3(S_AX, ld3(P.acc));
#endif
            cold.set(K_M, P.mass[i]); cold.set(K_R, ldm(P.radius + i)); cold.set(K_I, ldm(P.moi + i));
        } else {
            // padding lanes: finite, non-zero dummies (never read by live lanes, never stored)
            q.r = s3(sd(1. + b), sd(0.), sd(0.)); q.v = s3(sd(0.), sd(0.), sd(0.));
            q.L = v3(0., 0., 1.); q.s = v3(0., 0., 1.);
            cold.set6(R_ERR, v3(0., 0., 0.), v3(0., 0., 0.));
#if !PB_FIXED_N
            cold.set3(S_AX, v3(0., 0., 0.));
#endif
            cold.set(K_M, 1.); cold.set(K_R, 1.); cold.set(K_I, 1.);
        }
        if (sys_ok) {
            st.t = ldm(P.t + sys); st.last_hist = ldm(P.last_hist + sys);
            st.tswarn = ldm(P.tswarn + sys) != 0; st.status = ldm(P.status + sys); st.warnings = ldm(P.warnings + sys); st.hist_count = ldm(P.hist_count + sys);
        } else {
            st.t = 0.; st.last_hist = 0.; st.tswarn = true; st.status = PB200_STATUS_COMPLETED; st.warnings = 0; st.hist_count = 0;
        }
        st.steps_done = 0; st.n_hist_new = 0; st.event_step = 0;
    }
    {
        double rmax2 = 0.;
        if (ro.valid)
            for (int j = 0; j < n; j++) {
                if (j == b) continue;
                const int lo = b < j ? b : j, hi = b < j ? j : b;   // the table is filled for lo < hi (universe.rs:177-196)
                const double rr = __ldg(P.roche + ((size_t)(lo * n + hi)) * (size_t)P.n_sys + sys);
                rmax2 = fmax(rmax2, __dmul_rn(rr, rr));
            }
        // 8-body gravity: the host pairs are checked by the planets; the host lane's own (dummy) walk must not pass the
        // pre-test, or it fetches a Roche radius from global memory on every step (its distance to itself is zero)
        if (PB_FIXED_N == 8 && COORD == PB200_COORD_DEMOCRATIC_HELIOCENTRIC && !ro.planet) rmax2 = -1.;
        cold.set(K_ROCHE2, rmax2);
    }
    bool alive = sys_ok && st.status == PB200_STATUS_OK;
    make_consts(P, ro, cold, hl, b, sys);

    // constants of the transforms, all strict and in the reference's order; parked in the cold slots
    {
        __syncwarp();
        // (the gravitational masses are only read here and in make_consts: they stay in global memory)
        const size_t ns_ = (size_t)P.n_sys, sys_ = sys_ok ? sys : 0;
        const sd m_s = sd(cold.get(K_M)), mg_s = sd(ro.valid ? P.mass_g[(size_t)b * ns_ + sys_] : 1.);
        const sd M_s = sd(cold.getk(PB_HOST(P), K_M));       // m0
        const sd Mg_s = sd(P.mass_g[(size_t)PB_HOST(P) * ns_ + sys_]);
        // total mass as inertial_to_*_posvel accumulate it: host first, then the others in index order
        sd mtot = M_s;
        sd eta_k = sd(0.), mu_k = sd(0.);        // Jacobi: cumulative mass / mass_g up to and including this body
        sd mu = Mg_s;
        if (COORD != PB200_COORD_JACOBI) mtot = sd(0.) + M_s;
        for (int k = 0; k < n; k++) {
            if (k == PB_HOST(P)) continue;
            mtot = mtot + sd(cold.getk(k, K_M));
            mu = mu + sd(P.mass_g[(size_t)k * ns_ + sys_]);
            if (k == b) { eta_k = mtot; mu_k = mu; }
        }
        // per-body constants of the heliocentric transforms (whfast.rs:1015, 1101, 1112-1114), divided once
        sd back_w = sd(1.), whds_f = sd(1.);
        if (COORD == PB200_COORD_WHDS) { whds_f = (M_s + m_s) / M_s; back_w = m_s / (M_s + m_s); }
        if (COORD == PB200_COORD_DEMOCRATIC_HELIOCENTRIC) back_w = m_s / M_s;
        const sd kepler_mu = COORD == PB200_COORD_JACOBI ? mu_k : (COORD == PB200_COORD_WHDS ? Mg_s + mg_s : Mg_s);
        cold.set(K_KMU, kepler_mu.v);
        // the host's own columns of these three carry the per-system values (see ColdSlot)
        cold.set(K_ETAK, b == PB_HOST(P) ? mtot.v : eta_k.v);
        cold.set(K_BACKW, b == PB_HOST(P) ? make_rcp(mtot).y : back_w.v);
        cold.set(K_WHDSF, b == PB_HOST(P) ? make_rcp(M_s).y : whds_f.v);
        __syncwarp();
    }
    const int first_other = PB_HOST(P) == 0 ? 1 : 0;
    const sd zero = sd(0.), one = sd(1.);
#ifdef PB_NO_DIST
    constexpr bool DIST = false;
#else
    constexpr bool DIST = PB_DIST && COORD == PB200_COORD_DEMOCRATIC_HELIOCENTRIC;   // distributed ordered sums (8-body build)
#endif
    const S3 zero3 = s3(zero, zero, zero);
    const S3 one3 = s3(one, one, one);

#pragma unroll 1
    for (unsigned long long step = step_begin; step < step_end; step++) {
#if PB_STEP_BARRIER || PB_PHASE_BARRIER
        // one block-wide barrier per step keeps the warps of a block in the same region of the (large) loop body, so that
        // they share instruction-cache lines; it also makes the loop exit block-uniform
        if (!__syncthreads_or(alive)) break;
#else
        if (!__any_sync(FULL, alive)) break;
#endif
        // ---- historic snapshot (whfast.rs:237-261, output.rs:119-163)
        {
            bool first = st.last_hist < 0.;
            bool due = __dadd_rn(st.last_hist, P.hist_period) <= st.t;
            bool snap = alive && (first || due);
            if (__any_sync(FULL, snap)) {
                // refresh: evolving quantities and spin = L / I (universe.rs:305-316); it changes the live state too
                if (PB_FLAGS(P) & FLAG_EVO) {
                    (void)evolve_lane(P, ro, cold, b, sys, st.t, snap);
                    __syncwarp();
                    make_consts(P, ro, cold, hl, b, sys);
                }
                if (snap) { sd I = sd(cold.get(K_I)); q.s = v3((sd(q.L.x) / I).v, (sd(q.L.y) / I).v, (sd(q.L.z) / I).v); }   // spin = L / I (common.rs:9-11)
#if !PB_FIXED_N
                if ((PB_FLAGS(P) & FLAG_DYN) && (PB_FLAGS(P) & FLAG_EVO)) {
                    const S3 ss = strict(q.s);
                    update_lag_angle(P, ro, b, sys, st.t, (ss.x * ss.x) + (ss.y * ss.y) + (ss.z * ss.z), snap);
                }
#endif
                if (snap && ro.valid && st.hist_count < P.hist_capacity) {
                    const size_t ns = (size_t)P.n_sys;
                    const size_t i = (size_t)b * ns + sys, cs = (size_t)n * ns;
                    double denergy = 0.;
                    if ((PB_FLAGS(P) & FLAG_TIDES) && ro.t_on) {
                        // tides/common.rs:263-279 with the internals left by the last evaluation and the fresh spin
                        double ts[PB_TIDE_SCRATCH];
                        for (int k = 0; k < PB_TIDE_SCRATCH; k++) ts[k] = ldm(P.tide_scratch + i + k * cs);
                        V3 tp = v3(ts[0], ts[1], ts[2]), tv = v3(ts[3], ts[4], ts[5]);
                        double dist = ts[6], radvel = ts[7], orth_p = ts[8], diss_pm = ts[9];
                        V3 tdl = v3(ts[10], ts[11], ts[12]);
                        double factor2 = orth_p / dist;
                        V3 wxr = cross(q.s, tp);
                        denergy = -((1.0 / dist * (diss_pm + factor2 * radvel)) * dot(tp, tv)
                                    + factor2 * ((wxr.x - tv.x) * tv.x + (wxr.y - tv.y) * tv.y + (wxr.z - tv.z) * tv.z))
                                  - dot(tdl, q.s);
                    }
                    double* h = P.hist + (size_t)st.hist_count * PB_HIST_FIELDS * cs + i;
                    h[0 * cs] = st.t;
                    h[1 * cs] = q.r.x.v; h[2 * cs] = q.r.y.v; h[3 * cs] = q.r.z.v;
                    h[4 * cs] = q.s.x; h[5 * cs] = q.s.y; h[6 * cs] = q.s.z;
                    h[7 * cs] = q.v.x.v; h[8 * cs] = q.v.y.v; h[9 * cs] = q.v.z.v;
                    h[10 * cs] = cold.get(K_M); h[11 * cs] = cold.get(K_R); h[12 * cs] = ldm(P.rg2 + i);
                    h[13 * cs] = P.k2t[i]; h[14 * cs] = P.sigma[i]; h[15 * cs] = denergy;
                    h[16 * cs] = (PB_FLAGS(P) & FLAG_DYN) ? ldm(P.lag + i) : 0.;
                }
                if (snap) {
                    if (!first) st.last_hist = __dadd_rn(st.last_hist, P.hist_period); else st.last_hist = 0.;
                    st.n_hist_new += 1;
                    if (st.hist_count < P.hist_capacity) st.hist_count += 1;
                    else st.warnings |= PB200_WARN_HISTORY_DROPPED;   // cannot happen through pb200_ensemble_step's guard
                }
            }
        }
        // last step of this CTA's piece, or the step after which the system completes (whfast.rs:300, the same roundings)
        const bool leaves_now = (step + 1 == step_end) || (__dadd_rn(__dadd_rn(st.t, P.dt), P.dt) > P.time_limit);
        // internals needed by the NEXT snapshot's denergy_dt are those of this step's last evaluation
        const bool save_tides = (step + 1 == step_end) || (__dadd_rn(st.last_hist, P.hist_period) <= __dadd_rn(st.t, P.dt));

        // One code instance of the midpoint serves both halves of the step (whfast.rs:278 and :293); the drift-kick-drift
        // core runs between them. Likewise one instance of the Kepler solver serves both drifts.
#pragma unroll 1
        for (int half = 0; half < 2; half++) {
            if (half == 1) {
                const sd m_s = sd(cold.get(K_M)), M_s = sd(cold.getk(PB_HOST(P), K_M)), mtot = sd(cold.getk(PB_HOST(P), K_ETAK));
                const sd dt_s = sd(P.dt), hdt_s = sd(P.half_dt);
                // step-invariant divisors with their refined reciprocals (strict.cuh): 3 instructions per division
                srcp rM, rT;
                rM.b = M_s.v; rM.y = cold.getk(PB_HOST(P), K_WHDSF); rT.b = mtot.v; rT.y = cold.getk(PB_HOST(P), K_BACKW);
                S3 apos, avel;       // this body's alternative coordinates
                S3 spos, svel;       // the host slot of the alternative coordinates (centre of mass), replicated in the group
                // 8-body build: the scalar of (spos, svel) this lane owns (lane 0-2: spos, 3-5: svel) and, for lanes 0-2, the
                // matching component of svel; spos itself is not kept
                sd my_com = zero, my_sv = zero;
                S3 anew_s = zero3;
                const bool kwork = ro.planet && alive;
                // ---- inertial -> alternative coordinates (whfast.rs:881-1023)
                __syncwarp();   // the midpoint's reads of its slots are done: they become exchange space
                if (COORD == PB200_COORD_JACOBI) {
                    cold.set3(E_R, plain(q.r)); cold.set3(E_A, plain(q.v));
                    __syncwarp();
                    sd eta = M_s;
                    S3 s = eta * strict(cold.getk3(PB_HOST(P), E_R)), sv = eta * strict(cold.getk3(PB_HOST(P), E_A));
                    apos = one3; avel = zero3;
                    for (int k = 0; k < n; k++) {
                        if (k == PB_HOST(P)) continue;
                        sd mk = sd(cold.getk(k, K_M));
                        S3 rk = strict(cold.getk3(k, E_R)), vk = strict(cold.getk3(k, E_A));
                        sd ei = one / eta;
                        eta = eta + mk;
                        sd pme = eta * ei;
                        S3 pk = rk - s * ei, wk = vk - sv * ei;
                        if (b == k) { apos = pk; avel = wk; }
                        s = s * pme + mk * pk; sv = sv * pme + mk * wk;
                    }
                    sd mi = one / eta;
                    spos = s * mi; svel = sv * mi;
                } else {
                    // host first, then the others (whfast.rs:986-995)
                    S3 mr = q.r * m_s, mv = q.v * m_s;
                    if (DIST) {
                        // scalars 0-2: sum m r / M_tot (centre of mass), 3-5: sum m v / M_tot; lane c owns scalar c for the step
                        dist_put3(cold, E_A, 0, plain(mr)); dist_put3(cold, E_A, 3, plain(mv)); cold.set3(E_R, plain(q.r));
                        __syncwarp();
                        const unsigned row = dist_row(cold, E_A, b, 6);
                        my_com = dist_walk<false>(row, zero + sd(dist_ld(row, 0))) / rT;
                        cold.set(M_0, my_com.v);
                        __syncwarp();
                        svel = strict(v3(cold.getk(3, M_0), cold.getk(4, M_0), cold.getk(5, M_0)));
                        my_sv = sd(cold.getk(b < 3 ? b + 3 : 3, M_0));
                    } else {
                    cold.set3(E_A, plain(mr)); cold.set3(E_B, plain(mv)); cold.set3(E_R, plain(q.r));
                    __syncwarp();
                    S3 sr = ordered_sum_others(cold, zero3 + strict(cold.getk3(PB_HOST(P), E_A)), E_A, n, PB_HOST(P));
                    S3 sv = ordered_sum_others(cold, zero3 + strict(cold.getk3(PB_HOST(P), E_B)), E_B, n, PB_HOST(P));
                    spos = sr / rT; svel = sv / rT;
                    }
                    apos = q.r - strict(cold.getk3(PB_HOST(P), E_R));
                    avel = q.v - svel;
                    if (COORD == PB200_COORD_WHDS) avel = avel * sd(cold.get(K_WHDSF));
                }
#pragma unroll 1
                for (int phase = 0; phase < 2; phase++) {
                    if (phase == 1) {
                        // ---- kick (whfast.rs:558-625)
                        if (COORD == PB200_COORD_JACOBI) {
                            // inertial_to_jacobi_acc (whfast.rs:935-963) + jacobi_interaction_step (:566-592)
                            cold.set3(E_B, plain(anew_s));
                            __syncwarp();
                            sd eta = M_s;
                            S3 sa = eta * strict(cold.getk3(PB_HOST(P), E_B));
                            S3 aacc = zero3;
                            for (int k = 0; k < n; k++) {
                                if (k == PB_HOST(P)) continue;
                                sd mk = sd(cold.getk(k, K_M));
                                S3 ak = strict(cold.getk3(k, E_B));
                                sd ei = one / eta;
                                eta = eta + mk;
                                sd pme = eta * ei;
                                S3 ck = ak - sa * ei;
                                if (b == k) aacc = ck;
                                sa = sa * pme + mk * ck;
                            }
                            avel = avel + dt_s * aacc;
                            if (b != first_other) {
                                sd rj2i = one / (apos.x * apos.x + apos.y * apos.y + apos.z * apos.z + sd(1e-12));
                                sd rji = ssqrt(rj2i);
                                sd rj3im = rji * rj2i * sd(kG) * sd(cold.get(K_ETAK));
                                sd prefac = dt_s * rj3im;
                                avel = avel + prefac * apos;
                            }
                        } else if (COORD == PB200_COORD_DEMOCRATIC_HELIOCENTRIC) {
                            avel = avel + dt_s * anew_s;
                        } else {
                            sd f = M_s + m_s;
                            avel = s3(avel.x + dt_s * f * anew_s.x / rM, avel.y + dt_s * f * anew_s.y / rM, avel.z + dt_s * f * anew_s.z / rM);
                        }
                    }
                    // ---- jump (whfast.rs:495-556): before the Kepler drift in the second half, after it in the first
                    for (int jump_slot = 0; jump_slot < 2; jump_slot++) {
                        if (jump_slot == 1) {
#if PB_PHASE_BARRIER
                            __syncthreads();
#endif
                            kepler_step(kwork, apos, avel, sd(cold.get(K_KMU)), hdt_s, st.tswarn, st.warnings);
                            {
                                // WHFast.timestep_warning and the warning itself belong to the system, not to the planet whose
                                // drift raised it (whfast.rs:702-707): share them in the group (the host lane stores them)
                                unsigned int wv = st.warnings | (st.tswarn ? 0x80000000u : 0u);
                                for (int off = W >> 1; off > 0; off >>= 1) wv |= __shfl_xor_sync(FULL, wv, off);
                                st.warnings = wv & 0x7fffffffu; st.tswarn = (wv >> 31) != 0u;
                            }
                            if (DIST) { if (b < 3) my_com = my_com + hdt_s * my_sv; }
                            else spos = spos + hdt_s * svel;
                        }
                        if (COORD != PB200_COORD_JACOBI && jump_slot != phase) {
                            if (DIST) {
                                dist_put3(cold, E_C, 0, plain(m_s * avel));
                                __syncwarp();
                                const sd p = dist_walk<false>(dist_row(cold, E_C, b, 3), zero);
                                cold.set(M_0, (hdt_s * p / rM).v);
                                __syncwarp();
                                apos = s3(apos.x + sd(cold.getk(0, M_0)), apos.y + sd(cold.getk(1, M_0)), apos.z + sd(cold.getk(2, M_0)));
                            } else if (COORD == PB200_COORD_DEMOCRATIC_HELIOCENTRIC) {
                                cold.set3(E_C, plain(m_s * avel));
                                __syncwarp();
                                S3 p = ordered_sum_others(cold, zero3, E_C, n, PB_HOST(P));
                                apos = s3(apos.x + hdt_s * p.x / rM, apos.y + hdt_s * p.y / rM, apos.z + hdt_s * p.z / rM);
                            } else {
                                sd f = M_s + m_s;
                                S3 term = s3(m_s * avel.x / f, m_s * avel.y / f, m_s * avel.z / f);
                                cold.set3(E_C, plain(term));
                                __syncwarp();
                                S3 p = ordered_sum_others(cold, zero3, E_C, n, PB_HOST(P));
                                apos = s3(apos.x + hdt_s * (p.x - term.x), apos.y + hdt_s * (p.y - term.y), apos.z + hdt_s * (p.z - term.z));
                            }
                        }
                    }
                    // ---- alternative -> inertial (whfast.rs:1026-1155)
                    if (COORD == PB200_COORD_JACOBI) {
                        cold.set3(E_D, plain(apos)); cold.set3(E_C, plain(avel));
                        __syncwarp();
                        sd et = mtot;
                        S3 s = et * spos, sv = et * svel;
                        S3 nr = q.r, nv = q.v;
                        for (int k = n - 1; k >= 0; k--) {
                            if (k == PB_HOST(P)) continue;
                            sd mk = sd(cold.getk(k, K_M));
                            S3 pk = strict(cold.getk3(k, E_D)), wk = strict(cold.getk3(k, E_C));
                            sd ei = one / et;
                            s = (s - mk * pk) * ei; sv = (sv - mk * wk) * ei;
                            if (b == k) { nr = pk + s; nv = wk + sv; }
                            et = et - mk;
                            s = s * et; sv = sv * et;
                        }
                        if (ro.host) { sd mi = one / et; nr = s * mi; nv = sv * mi; }
                        if (alive) { q.r = nr; q.v = nv; }
                    } else {
                        // positions (whfast.rs:1128-1155); the host lane divides a dummy instead of its zero vector
                        S3 num = ro.planet ? apos * m_s : one3;
                        S3 term = num / rT;
                        const S3 vterm = avel * sd(cold.get(K_BACKW));
                        if (DIST) {
                            // scalars 0-2: star position, 3-5: star velocity (read in the second pass only)
                            dist_put3(cold, E_C, 0, plain(term));
                            if (phase == 1) dist_put3(cold, E_C, 3, plain(vterm));
                            __syncwarp();
                            cold.set(M_0, dist_walk<true>(dist_row(cold, E_C, b, 6), my_com).v);
                            __syncwarp();
                            const S3 star_r = strict(v3(cold.getk(0, M_0), cold.getk(1, M_0), cold.getk(2, M_0)));
                            const S3 nr = ro.host ? star_r : apos + star_r;
                            if (alive) q.r = nr;
                            if (phase == 1) {
                                S3 nv = avel + svel;
                                if (ro.host) nv = strict(v3(cold.getk(3, M_0), cold.getk(4, M_0), cold.getk(5, M_0)));
                                if (alive) q.v = nv;
                            }
                        } else {
                        cold.set3(E_D, plain(term));
                        if (phase == 1) cold.set3(E_A, plain(vterm));
                        __syncwarp();
                        S3 star_r = ordered_diff_others(cold, spos, E_D, n, PB_HOST(P));
                        S3 nr = ro.host ? star_r : apos + star_r;
                        if (alive) q.r = nr;
                        if (phase == 1) {
                            // velocities (whfast.rs:1090-1126); those of the first drift are dead (the kick overwrites them)
                            S3 nv = (COORD == PB200_COORD_WHDS ? avel / sd(cold.get(K_WHDSF)) : avel) + svel;
                            S3 star_v = ordered_diff_others(cold, svel, E_A, n, PB_HOST(P));
                            if (ro.host) nv = star_v;
                            if (alive) q.v = nv;
                        }
                        }
                    }
                    if (phase == 0) {
                        // ---- gravity (whfast.rs:281)
                        int fail;
                        cold.set3(E_R, plain(q.r));
#if PB_PHASE_BARRIER
                        __syncthreads();
#else
                        __syncwarp();
#endif
                        int code;
#if PB_FIXED_N == 8
                        if (COORD == PB200_COORD_DEMOCRATIC_HELIOCENTRIC) {
                            anew_s = gravity_n8_dh(P, ro, cold, b, sys, q, code);
                            (void)fail;
                        } else
#endif
                        {
                            anew_s = gravity<COORD>(P, ro, cold, gb, b, sys, q, fail);
                            // group-wide failure: lowest body index wins, like the reference's loop order
                            code = (ro.valid && fail) ? ((b << 4) | fail) : 0x7fffffff;
                        }
                        for (int off = W >> 1; off > 0; off >>= 1) { int o = __shfl_xor_sync(FULL, code, off); code = o < code ? o : code; }
                        const bool died = alive && code != 0x7fffffff;
#if !PB_FIXED_N
                        if (alive) cold.set3(S_AX, plain(anew_s));
#endif
                        // The Newtonian acceleration is part of the state image (Particle.inertial_acceleration) but not of the
                        // on-chip state: the step that may be the system's last of this launch writes it out
                        if (alive && (leaves_now || died) && ro.valid) {
                            const size_t ns = (size_t)P.n_sys;
                            const size_t i = (size_t)b * ns + sys, cs = (size_t)n * ns;
                            P.acc[i] = anew_s.x.v; P.acc[i + cs] = anew_s.y.v; P.acc[i + 2 * cs] = anew_s.z.v;
                        }
                        if (died) {
                            st.status = code & 15; st.event_step = st.steps_done; alive = false;
                            store_lane(P, ro, cold, sys, b, q, st);
                        }
                    }
                }
            }
            midpoint<COORD, GR, ARITH>(P, ro, cold, gb, hl, b, alive, q, st.t, half == 0, st.warnings, half == 1 && save_tides, sys);
        }

        if (alive) {
            st.t = __dadd_rn(st.t, P.dt);
            st.steps_done += 1;
            if (__dadd_rn(st.t, P.dt) > P.time_limit) {
                st.status = PB200_STATUS_COMPLETED; st.event_step = st.steps_done; alive = false;
                store_lane(P, ro, cold, sys, b, q, st);
            }
        }
    }
    if (alive) store_lane(P, ro, cold, sys, b, q, st);
    if (P.n_pieces > 1) {
        // hand the group over to the CTA that runs its next piece: state first, then the flag
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) { volatile unsigned int* flag = P.sched + 1 + group; *flag = piece + 1; }
    }
}

}  // namespace PB_NS
