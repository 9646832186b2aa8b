// ensemble_host.hpp — host-side state of one device-resident ensemble and the launch helper shared by the translation
// units of the library (pb200_api.cu: C ABI; kernels_tu.cu: one object per geometry build x arithmetic mode of the step kernel).
#pragma once
#include <cuda_runtime.h>
#include <cmath>
#include <cstdlib>
#include <string>
#include <vector>
#include "whfast_kernel.cuh"

struct pb200_ensemble {
    int device = 0;
    size_t n_sys = 0;
    int n_bodies = 0;
    pb200::KParams P{};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_ms = 0.f;
    bool timing_pending = false;
    uint64_t launches = 0;
    std::vector<void*> allocations;
    pb200_case_t tmpl{};                // structure + uniform scalars (system 0 image at creation)
    std::vector<pb200_case_t> cases;    // per-system images when n_cases == n_systems (params may differ), else 1
    int coord = 0, gr = PB200_GR_DISABLED;
    // device arrays that are not part of KParams constness
    double *d_mass = nullptr, *d_mass_g = nullptr, *d_sigma = nullptr, *d_k2t = nullptr, *d_k2f = nullptr, *d_roche = nullptr;
    double *d_energy = nullptr, *d_angmom = nullptr;
    double *d_wind_k = nullptr, *d_wind_sat = nullptr, *d_diss = nullptr, *d_diss_scale = nullptr;
    double* d_gather = nullptr;   // staging of pb200_ensemble_get_case
    unsigned int* d_records = nullptr;
    size_t records_capacity = 0;
    double recovery_snapshot_period = 0.;
    int arithmetic = PB200_ARITH_HYBRID;
    int sm_count = 0;
    bool perturbed = false;       // built by pb200_ensemble_create_perturbed: heliocentric fields of the image are per member
    bool narrow_blocks = false;   // PB200_NARROW_BLOCKS=1: never use the 384-thread build of the 8-body kernel (A/B tests)
    bool pair_lanes = false;      // PB200_PAIR_LANES=1: 3-body Jacobi systems always on the two-lane build (A/B tests of the passive-planet build)
    bool force_generic = false;   // PB200_FORCE_GENERIC=1 in the environment: bypass the compile-time geometry builds (A/B tests)
    // host mirror of the ensemble clock (exact snapshot counting without a device round trip); invalid after an upload of
    // current_time (uniform_clock = false: the device is asked instead)
    bool uniform_clock = true;
    bool clock_dirty = false;     // current_time was uploaded: the mirror is re-read from the device before the next step
    double clock_t = 0., clock_last_hist = -1.;
    size_t hist_pending_host = 0;
    unsigned last_pieces = 1;     // time slices of the last step launch (diagnostics)
    const char* last_kernel = "";  // build that ran the last step launch (diagnostics)
};

// Wave quantisation: `grid` CTAs of equal length on `slots` resident CTAs leave the last wave partly empty (65536
// TRAPPIST-1 systems = 8192 CTAs on 888 slots = 9.23 waves: 7.7 % of the GPU-time idle; the 8192 systems that one GPU of
// eight holds = 1024 CTAs = 1.15 waves: 42 % idle). Cutting every CTA's steps into k consecutive pieces makes the unit of
// scheduling k times shorter: the ticket order (kernel prologue) turns the launch into a work queue of grid x k items.
// Returns the number of pieces (1 = plain launch). PB200_PIECES in the environment overrides (experiments).
// k is the smallest count whose wave efficiency is within 1 % of the best one reachable with pieces of >= 25 steps (the state
// crosses HBM once per piece: ~0.4 KB per body against ~5 kflop per body-step).
inline unsigned pb200_plan_pieces(unsigned grid, unsigned slots, unsigned long long n_steps) {
    if (const char* f = getenv("PB200_PIECES")) { int k = atoi(f); if (k >= 1 && (unsigned long long)k <= n_steps) return (unsigned)k; }
    if (slots == 0 || grid <= slots) return 1;   // everything is resident at once: pieces of a group would only serialise
    auto eff = [&](unsigned k) { double w = (double)grid * k / slots; return w / std::ceil(w); };
    const unsigned k_max = (unsigned)std::min<unsigned long long>(64ull, std::max<unsigned long long>(1ull, n_steps / 25ull));
    double best_eff = eff(1);
    for (unsigned k = 2; k <= k_max; k++) best_eff = std::max(best_eff, eff(k));
    for (unsigned k = 1; k <= k_max; k++)
        if (eff(k) >= best_eff - 0.01) return k;
    return 1;
}

// `threads` = systems x lanes per system; `block` = the CTA size the kernel was compiled for (PB_BLOCK of its translation unit).
template <class K>
inline cudaError_t pb200_launch_sliced(pb200_ensemble* e, K kernel, size_t smem_bytes, int block, int& configured_device, int& blocks_per_sm,
                                       size_t threads, unsigned long long n) {
    const unsigned grid = (unsigned)((threads + (size_t)block - 1) / (size_t)block);
    // PB200_SMEM_PAD_KB (experiments only): extra dynamic shared memory per CTA, to lower the residency of the same binary
    static const size_t pad = []() { const char* v = getenv("PB200_SMEM_PAD_KB"); return v ? (size_t)atoi(v) * 1024 : (size_t)0; }();
    const size_t smem = smem_bytes + pad;
    if (configured_device != e->device) {
        // the cold slots need more than the default 48 KB of dynamic shared memory
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kernel, block, smem);
        if (err != cudaSuccess) return err;
        configured_device = e->device;
    }
    e->P.n_groups = grid;
    e->P.n_pieces = pb200_plan_pieces(grid, (unsigned)(blocks_per_sm * e->sm_count), n);
    e->last_pieces = e->P.n_pieces;
    if (e->P.n_pieces > 1) {
        cudaError_t err = cudaMemsetAsync(e->P.sched, 0, (size_t)(grid + 1) * sizeof(unsigned int), e->stream);
        if (err != cudaSuccess) return err;
    }
    kernel<<<grid * e->P.n_pieces, block, smem, e->stream>>>(e->P, n);
    return cudaGetLastError();
}

// One entry per translation unit of kernels_tu.cu. The run-time-geometry entries (one per arithmetic mode) dispatch on
// e->coord / e->gr; the compile-time geometry entries dispatch on e->arithmetic (the caller, pb200_ensemble_step, checks
// that the ensemble has the build's geometry and effect set).
cudaError_t pb200_launch_generic_fast(pb200_ensemble* e, size_t threads, unsigned long long n);
cudaError_t pb200_launch_generic_strict(pb200_ensemble* e, size_t threads, unsigned long long n);
cudaError_t pb200_launch_generic_hybrid(pb200_ensemble* e, size_t threads, unsigned long long n);
cudaError_t pb200_launch_n8(pb200_ensemble* e, size_t threads, unsigned long long n);        // 8 bodies, DH, tides + flattening + Kidder
cudaError_t pb200_launch_n8w(pb200_ensemble* e, size_t threads, unsigned long long n);       // the same in 384-thread CTAs (one per SM)
// lane = planet builds (small_step.cuh, kernels_small_tu.cu): N - 1 lanes per system, the host body replicated in every lane
cudaError_t pb200_launch_s2(pb200_ensemble* e, unsigned long long n);    // 2 bodies, DH, tides + flattening + Kidder          (config 1)
cudaError_t pb200_launch_s2t(pb200_ensemble* e, unsigned long long n);   // 2 bodies, DH, tides only                            (config 2)
cudaError_t pb200_launch_s3(pb200_ensemble* e, unsigned long long n);    // 3 bodies, DH, tides + flattening + Kidder          (config 3)
cudaError_t pb200_launch_s3e(pb200_ensemble* e, unsigned long long n);   // 3 bodies, DH, the same + evolution                 (config 3 evolving)
cudaError_t pb200_launch_s3j(pb200_ensemble* e, unsigned long long n);   // 3 bodies, Jacobi, the same + evolution
cudaError_t pb200_launch_s2any(pb200_ensemble* e, unsigned long long n);   // 2 bodies, DH, any subset of tides / flattening / Kidder / evolution (run-time flags)
cudaError_t pb200_launch_s3any(pb200_ensemble* e, unsigned long long n);   // 3 bodies, DH, the same
cudaError_t pb200_launch_s3jany(pb200_ensemble* e, unsigned long long n);  // 3 bodies, Jacobi, the same
cudaError_t pb200_launch_s3p(pb200_ensemble* e, unsigned long long n);   // the same with body 2 outside every effect: thread = system (config 5)
