// kernels_tu.cu — one translation unit per build of the step kernel. build.py compiles this file several times in parallel:
//   -DPB_TU_GENERIC=<arith>            run-time geometry (any bodies / host / coordinates / GR variant), one arithmetic mode
//   -DPB_TU_FIXED=8 [-DPB_TU_WIDE=1]    compile-time geometry build, all three arithmetic modes: 8 bodies, host 0, tides +
//                                       flattening + GR Kidder1995, democratic heliocentric (config 4, TRAPPIST-1), in 64- or
//                                       384-thread CTAs
// (the 2- and 3-body configurations have their own lane = planet kernel: kernels_small_tu.cu)
#if defined(PB_TU_GENERIC)
#define PB_NS pbgen
#define PB_FIXED_N 0
#define PB_FIXED_W 0
#define PB_FIXED_SHIFT 0
#define PB_FIXED_FLAGS 0
#elif PB_TU_FIXED == 8
#ifdef PB_TU_WIDE
// One 384-thread CTA per SM instead of six 64-thread ones: the same 12 warps, but they start together and stay loosely in
// phase, so that an instruction-cache line fetched for one warp serves the others (instruction cache).
#define PB_NS pbn8w
#define PB_BLOCK 384
#define PB_MIN_BLOCKS 1
#else
#define PB_NS pbn8
#endif
#define PB_FIXED_N 8
#define PB_FIXED_W 8
#define PB_FIXED_SHIFT 3
#define PB_FIXED_FLAGS (pb200::FLAG_TIDES | pb200::FLAG_FLAT | pb200::FLAG_GR)
#else
#error "kernels_tu.cu: define PB_TU_GENERIC=<arith> or PB_TU_FIXED=8"
#endif
#include "ensemble_host.hpp"
#include "whfast_step.cuh"

namespace {

template <int COORD, int GR, int ARITH>
cudaError_t launch_one(pb200_ensemble* e, size_t threads, unsigned long long n) {
    static thread_local int configured_device = -1, blocks_per_sm = 0;
    return pb200_launch_sliced(e, PB_NS::whfast_steps_kernel<COORD, GR, ARITH>, PB_NS::kSmemBytes, PB_BLOCK, configured_device, blocks_per_sm, threads, n);
}

#if defined(PB_TU_GENERIC)
template <int COORD>
cudaError_t launch_gr(pb200_ensemble* e, size_t threads, unsigned long long n) {
    switch (e->gr) {
        case PB200_GR_KIDDER1995: return launch_one<COORD, PB200_GR_KIDDER1995, PB_TU_GENERIC>(e, threads, n);
        case PB200_GR_ANDERSON1975: return launch_one<COORD, PB200_GR_ANDERSON1975, PB_TU_GENERIC>(e, threads, n);
        case PB200_GR_NEWHALL1983: return launch_one<COORD, PB200_GR_NEWHALL1983, PB_TU_GENERIC>(e, threads, n);
        default: return launch_one<COORD, PB200_GR_DISABLED, PB_TU_GENERIC>(e, threads, n);
    }
}
cudaError_t launch_generic(pb200_ensemble* e, size_t threads, unsigned long long n) {
    switch (e->coord) {
        case PB200_COORD_JACOBI: return launch_gr<PB200_COORD_JACOBI>(e, threads, n);
        case PB200_COORD_DEMOCRATIC_HELIOCENTRIC: return launch_gr<PB200_COORD_DEMOCRATIC_HELIOCENTRIC>(e, threads, n);
        default: return launch_gr<PB200_COORD_WHDS>(e, threads, n);
    }
}
#else
template <int COORD, int GR>
cudaError_t launch_arith(pb200_ensemble* e, size_t threads, unsigned long long n) {
    switch (e->arithmetic) {
        case PB200_ARITH_FAST: return launch_one<COORD, GR, PB200_ARITH_FAST>(e, threads, n);
        case PB200_ARITH_STRICT: return launch_one<COORD, GR, PB200_ARITH_STRICT>(e, threads, n);
        default: return launch_one<COORD, GR, PB200_ARITH_HYBRID>(e, threads, n);
    }
}
#endif

}  // namespace

#if defined(PB_TU_GENERIC)
#if PB_TU_GENERIC == 0
cudaError_t pb200_launch_generic_fast(pb200_ensemble* e, size_t threads, unsigned long long n) { return launch_generic(e, threads, n); }
#elif PB_TU_GENERIC == 1
cudaError_t pb200_launch_generic_strict(pb200_ensemble* e, size_t threads, unsigned long long n) { return launch_generic(e, threads, n); }
#else
cudaError_t pb200_launch_generic_hybrid(pb200_ensemble* e, size_t threads, unsigned long long n) { return launch_generic(e, threads, n); }
#endif
#elif PB_TU_FIXED == 8
#ifdef PB_TU_WIDE
cudaError_t pb200_launch_n8w(pb200_ensemble* e, size_t threads, unsigned long long n) {
#else
cudaError_t pb200_launch_n8(pb200_ensemble* e, size_t threads, unsigned long long n) {
#endif
    return launch_arith<PB200_COORD_DEMOCRATIC_HELIOCENTRIC, PB200_GR_KIDDER1995>(e, threads, n);
}
#endif
