// kernels_small_tu.cu — one translation unit per build of the lane = planet step kernel (small_step.cuh) for the 2- and
// 3-body BASELINE configurations. build.py compiles this file once per -DPB_TU_SMALL=<id>, all three arithmetic modes each:
//   2   2 bodies, democratic heliocentric, tides + flattening + GR Kidder1995              (config 1)
//   20  2 bodies, democratic heliocentric, tides only                                      (config 2)
//   3   3 bodies, democratic heliocentric, tides + flattening + GR Kidder1995              (config 3)
//   30  3 bodies, democratic heliocentric, the same + evolution tables                     (config 3 evolving)
//   31  3 bodies, Jacobi, the same + evolution tables, two lanes per system
//   32  the same with body 2 outside every effect: one thread per system                   (config 5)
//   29 / 39 / 38   catch-all builds (2 bodies DH / 3 bodies DH / 3 bodies Jacobi): every effect compiled in, the ensemble's
//       flag word selects at run time — any subset of tides, flattening, GR Kidder1995, evolution tables
#if PB_TU_SMALL == 2
#define PB_NS pbs2
#define PB_S_N 2
#define PB_S_COORD PB200_COORD_DEMOCRATIC_HELIOCENTRIC
#define PB_S_FLAGS (pb200::FLAG_TIDES | pb200::FLAG_FLAT | pb200::FLAG_GR)
#define PB_S_ENTRY pb200_launch_s2
#elif PB_TU_SMALL == 20
#define PB_NS pbs2t
#define PB_S_N 2
#define PB_S_COORD PB200_COORD_DEMOCRATIC_HELIOCENTRIC
#define PB_S_FLAGS (pb200::FLAG_TIDES)
#define PB_S_ENTRY pb200_launch_s2t
#elif PB_TU_SMALL == 3
#define PB_NS pbs3
#define PB_S_N 3
#define PB_S_COORD PB200_COORD_DEMOCRATIC_HELIOCENTRIC
#define PB_S_FLAGS (pb200::FLAG_TIDES | pb200::FLAG_FLAT | pb200::FLAG_GR)
#define PB_S_ENTRY pb200_launch_s3
#elif PB_TU_SMALL == 30
#define PB_NS pbs3e
#define PB_S_N 3
#define PB_S_COORD PB200_COORD_DEMOCRATIC_HELIOCENTRIC
#define PB_S_FLAGS (pb200::FLAG_TIDES | pb200::FLAG_FLAT | pb200::FLAG_GR | pb200::FLAG_EVO)
#define PB_S_ENTRY pb200_launch_s3e
#elif PB_TU_SMALL == 31
#define PB_NS pbs3j
#define PB_S_N 3
#define PB_S_COORD PB200_COORD_JACOBI
#define PB_S_FLAGS (pb200::FLAG_TIDES | pb200::FLAG_FLAT | pb200::FLAG_GR | pb200::FLAG_EVO)
#define PB_S_ENTRY pb200_launch_s3j
#elif PB_TU_SMALL == 29
#define PB_NS pbs2any
#define PB_S_N 2
#define PB_S_COORD PB200_COORD_DEMOCRATIC_HELIOCENTRIC
#define PB_S_FLAGS (pb200::FLAG_TIDES | pb200::FLAG_FLAT | pb200::FLAG_GR | pb200::FLAG_EVO | PB_NS::SMALL_RT)
#define PB_S_ENTRY pb200_launch_s2any
#elif PB_TU_SMALL == 39
#define PB_NS pbs3any
#define PB_S_N 3
#define PB_S_COORD PB200_COORD_DEMOCRATIC_HELIOCENTRIC
#define PB_S_FLAGS (pb200::FLAG_TIDES | pb200::FLAG_FLAT | pb200::FLAG_GR | pb200::FLAG_EVO | PB_NS::SMALL_RT)
#define PB_S_ENTRY pb200_launch_s3any
#elif PB_TU_SMALL == 38
#define PB_NS pbs3jany
#define PB_S_N 3
#define PB_S_COORD PB200_COORD_JACOBI
#define PB_S_FLAGS (pb200::FLAG_TIDES | pb200::FLAG_FLAT | pb200::FLAG_GR | pb200::FLAG_EVO | PB_NS::SMALL_RT)
#define PB_S_ENTRY pb200_launch_s3jany
#elif PB_TU_SMALL == 32
#define PB_NS pbs3p
#define PB_S_N 3
#define PB_S_COORD PB200_COORD_JACOBI
#define PB_S_FLAGS (pb200::FLAG_TIDES | pb200::FLAG_FLAT | pb200::FLAG_GR | pb200::FLAG_EVO)
#define PB_S_PASSIVE true
#define PB_S_ENTRY pb200_launch_s3p
#else
#error "kernels_small_tu.cu: define PB_TU_SMALL=<2|20|3|30|31|32|29|39|38>"
#endif
#ifndef PB_S_PASSIVE
#define PB_S_PASSIVE false
#endif
#include "ensemble_host.hpp"
#include "small_step.cuh"

namespace {

// CTA size: one CTA of 8 warps (3 bodies) or two of 4 warps (2 bodies) per SM once the ensemble fills three quarters of the
// GPU that way (warps that start together share instruction-cache lines), else single-warp CTAs that spread over every SM.
// PB200_SMALL_BLOCK=32 in the environment forces the small CTAs (A/B runs).
constexpr int kLanes = PB_S_PASSIVE ? 1 : PB_S_N - 1;   // lanes per system
// (the passive-planet build holds 140 slots per thread: six warps per SM, as one CTA)
constexpr int kBigBlock = PB_S_PASSIVE ? 192 : kLanes == 2 ? 256 : 128;

template <int ARITH, int BLK>
cudaError_t launch_blk(pb200_ensemble* e, size_t threads, unsigned long long n) {
    static thread_local int configured_device = -1, blocks_per_sm = 0;
    return pb200_launch_sliced(e, PB_NS::small_steps_kernel<PB_S_N, PB_S_COORD, PB_S_FLAGS, ARITH, BLK, PB_S_PASSIVE>,
                               PB_NS::small_smem_bytes<PB_S_COORD, BLK, PB_S_PASSIVE>(), BLK,
                               configured_device, blocks_per_sm, threads, n);
}

template <int ARITH>
cudaError_t launch_one(pb200_ensemble* e, unsigned long long n) {
    // lane = planet: N - 1 lanes per system (one when body 2 rides in body 1's thread)
    const size_t threads = e->n_sys * (size_t)kLanes;
    static const bool force_small = []() { const char* v = getenv("PB200_SMALL_BLOCK"); return v && atoi(v) == 32; }();
    const bool big = !force_small && 4 * threads >= 3 * (size_t)kBigBlock * (size_t)e->sm_count * (size_t)(256 / kBigBlock);
    return big ? launch_blk<ARITH, kBigBlock>(e, threads, n) : launch_blk<ARITH, 32>(e, threads, n);
}

}  // namespace

cudaError_t PB_S_ENTRY(pb200_ensemble* e, unsigned long long n) {
    switch (e->arithmetic) {
        case PB200_ARITH_FAST: return launch_one<PB200_ARITH_FAST>(e, n);
        case PB200_ARITH_STRICT: return launch_one<PB200_ARITH_STRICT>(e, n);
        default: return launch_one<PB200_ARITH_HYBRID>(e, n);
    }
}
