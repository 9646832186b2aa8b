// cold_slots.cuh — the map of the per-thread shared-memory slots (`Cold`, whfast_kernel.cuh) of one geometry build.
// Included once per build inside its own namespace (PB_NS); no include guard on purpose.
//
// 73 slots x 64 threads x 8 B = 36.5 KB per CTA: six CTAs (12 warps) per SM, the same residency that the 168-register
// budget allows. Every arithmetic mode (fast, exact, hybrid) lives in the same 73 slots:
//   * the constants of the FAST forces are folded products (every step-invariant division and power done once);
//   * the EXACT forces (exact_effects.cuh) rebuild the reference's products from the base quantities (m, R^5, R^10, sigma,
//     k2, k2f of the body and of the host, read from the host's column) in the reference's association order — about 30
//     FP64 instructions per evaluation instead of 16 more slots — and share with the fast forces whatever can be rounded
//     the reference's way without cost: the tidal numerators (C_AS, C_AP), 1 / m, mu_host + mu and the 13 polynomials in the
//     GR factor (G_0 ..).
// The run-time-geometry build alone carries three more slots for the Newtonian acceleration, which only the Anderson1975 /
// Newhall1983 variants read back inside the midpoint (77 slots: five CTAs per SM).
namespace PB_NS {

enum ColdSlot : int {
    // Kahan residuals (whfast.rs:117-119) and the midpoint's working set (whfast.rs:333-337)
    S_EVX, S_EVY, S_EVZ, S_ELX, S_ELY, S_ELZ,
    S_VOX, S_VOY, S_VOZ, S_LOX, S_LOY, S_LOZ,
    S_DVX, S_DVY, S_DVZ, S_DLX, S_DLY, S_DLZ,
    S_RX, S_RY, S_RZ,            // inertial position while the midpoint runs
    // body parameters and the base quantities of the exact forces (the host's are read from the host's column)
    K_M, K_R, K_I, K_SIG, K_K2T, K_K2F, K_R5, K_R10,
    // constants of the fast perturbation forces (every division with step-invariant operands is done once)
    // (host-body quantities — its mass, inertia, 1/M — are read from the host's own column with getk, they have no slot)
    // (C_AS .. C_MGS are one region of 16-byte pair cells: Cold::get2 / set2)
    C_INVI, C_INVM, C_AS, C_AP, C_KS, C_KP, C_ZP, C_ZH, C_DP1, C_DS1, C_MFA, C_SXS, C_BK, C_MGS,
    // constants of the coordinate transforms (strict)
    // The host's columns of the last three are meaningless for the host body itself and carry the per-system values:
    // K_ETAK <- total mass, K_BACKW <- refined reciprocal of the total mass, K_WHDSF <- refined reciprocal of the host
    // mass (strict.cuh, srcp).
    K_KMU, K_BACKW, K_WHDSF, K_ETAK,
    // 13 polynomials in the GR factor (general_relativity.rs:197-205, 256-268), rounded like the reference: six pair cells
    // and a single. Builds whose GR variant is Anderson1975 / Newhall1983 never read them: there the region is the
    // exchange space of those variants (X_0 ..).
    G_0, G_1, G_2, G_3, G_4, G_5, G_6, G_7, G_8, G_9, G_10, G_11, G_12,
    // exact forces: mass_factor * (M + m) and the reduced mass (general_relativity.rs:321, 383)
    Z_MFM, Z_MURED,
    // spin exchange (the host's fresh spin for the group)
    E_S, E_S1, E_S2,
    // exchange space of the midpoint: six contributions to the host sums / their totals, |spin|^2, Roche bound
    M_0, M_1, M_2, M_3, M_4, M_5, M_6, M_7,
#if !PB_FIXED_N
    S_AX, S_AY, S_AZ,            // Newtonian acceleration of the last gravity evaluation (read by the Anderson / Newhall variants)
#endif
    N_COLD_SLOTS,
    K_ROCHE2 = M_7,   // max over j > b of the squared Roche radius of the pair (b, j): the cheap pre-test of gravity()
    X_0 = G_0, X_3 = G_3, X_6 = G_6
};
// pair regions of the midpoint (six slots each): Kahan residuals (v, L), originals (v, L), increments (v, L)
enum : int { R_ERR = S_EVX, R_ORIG = S_VOX, R_INCR = S_DVX };
// exchange triples of the core (dead midpoint slots)
enum : int { E_A = S_VOX, E_B = S_LOX, E_C = S_DVX, E_D = S_DLX, E_R = S_RX };

constexpr size_t kSmemBytes = (size_t)N_COLD_SLOTS * PB_BLOCK * sizeof(double);

}  // namespace PB_NS
