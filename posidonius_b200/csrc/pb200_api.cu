// pb200_api.cu — C ABI of the B200 ensemble integrator (include/posidonius_b200.h).
// Host side: case validation (no CPU fallback: unsupported effects are rejected), SoA packing,
// device residency, launches of the step kernel, status / state / history transfer.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "whfast_kernel.cuh"

// The step kernels themselves are compiled in their own translation units (kernels_tu.cu, one object per geometry build
// and arithmetic mode, built in parallel by build.py); this file sees the run-time-geometry device headers only for the
// helpers that the small kernels below share with them (update_lag_angle).
#define PB_NS pbgen
#define PB_FIXED_N 0
#define PB_FIXED_W 0
#define PB_FIXED_SHIFT 0
#define PB_FIXED_FLAGS 0
#include "whfast_step.cuh"
#include "ensemble_host.hpp"

using namespace pb200;

static thread_local std::string g_last_error;
static int set_error(int code, const std::string& msg) { g_last_error = msg; return code; }
int pb200_set_error_message(int code, const std::string& msg) { return set_error(code, msg); }   // for host/case_io.cpp

#define CUDA_TRY(expr)                                                                                         \
    do {                                                                                                        \
        cudaError_t _e = (expr);                                                                                \
        if (_e != cudaSuccess) return set_error(PB200_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

// ---------------------------------------------------------------------------------------------
// small device kernels around the step kernel
namespace pb200 {

// Integrator::initialize_physical_values (whfast.rs:226-233): evolving quantities at t, spin = L/I
// (universe.rs:305-316) and the Roche radii table (universe.rs:177-196).
// evolved (radius, rg2) of body b at time t (effects/evolution.rs:458-476); inputs returned unchanged when NonEvolving
__device__ __forceinline__ void evolved_values(const KParams& P, int b, double t, double& R, double& rg2) {
    int ti = P.evo_table[b];
    if (ti < 0) return;
    const DevTable& T = P.tables[ti];
    int i = table_upper(T.time, T.n_rows, t);
    if (T.interp_radius) R = table_interp(T.time, T.radius, T.n_rows, i, t);
    if ((P.evo_rg2 >> b) & 1u) rg2 = table_interp(T.time, T.rg2, T.n_rows, i, t);
}

__global__ void init_physical_kernel(const __grid_constant__ KParams P) {
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t ns = (size_t)P.n_sys;
    if (gtid >= ns * (size_t)P.n_bodies) return;
    const size_t sys = gtid % ns;
    const int b = (int)(gtid / ns);
    const size_t i = (size_t)b * ns + sys, cs = (size_t)P.n_bodies * ns;
    const double m = P.mass[i];
    const double R0 = P.radius[i], g0 = P.rg2[i];
    double R = R0, g = g0, I = P.moi[i];
    if (P.flags & FLAG_EVO) evolved_values(P, b, P.t[sys], R, g);
    if (R != R0 || g != g0) I = (sd(m) * sd(g) * (sd(R) * sd(R))).v;   // evolution.rs:527-531
    // calculate_spin (particles/common.rs:9-12) and the spin-dependent evolving quantities (evolution.rs:548-567)
    const sd sx = sd(P.L[i]) / sd(I), sy = sd(P.L[i + cs]) / sd(I), sz = sd(P.L[i + 2 * cs]) / sd(I);
    P.spin[i] = sx.v; P.spin[i + cs] = sy.v; P.spin[i + 2 * cs] = sz.v;
    if ((P.flags & FLAG_DYN) && (P.flags & FLAG_EVO)) {
        pbgen::Roles ro{};
        ro.valid = true;
        pbgen::update_lag_angle(P, ro, b, sys, P.t[sys], (sx * sx) + (sy * sy) + (sz * sz), true);
    }
    // Roche radii use the radii AFTER the evolution update (whfast.rs:231-232). Body j's evolved radius is recomputed here:
    // an evolving body's radius depends on t only and a non-evolving one is never rewritten, so it does not matter
    // whether j's thread has already stored its value — no ordering between threads is needed.
    double* roche = const_cast<double*>(P.roche);
    for (int j = 0; j < P.n_bodies; j++) {
        if (j == b) continue;
        const size_t ij = (size_t)j * ns + sys;
        double mj = P.mass[ij], Rj = P.radius[ij], gj = 1.;
        if (P.flags & FLAG_EVO) evolved_values(P, j, P.t[sys], Rj, gj);
        double rr;
        if (m > mj) rr = (Rj / 0.462) * cbrt(m / mj);
        else rr = (R / 0.462) * cbrt(mj / m);
        roche[((size_t)(b * P.n_bodies + j)) * ns + sys] = rr;
    }
    P.radius[i] = R; P.rg2[i] = g; P.moi[i] = I;
}

// Universe::compute_total_energy / compute_total_angular_momentum (universe.rs:625-658) after a heliocentric refresh.
__global__ void summary_kernel(const __grid_constant__ KParams P, double* energy, double* angmom) {
    const size_t sys = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t ns = (size_t)P.n_sys;
    if (sys >= ns) return;
    const int n = P.n_bodies;
    const size_t cs = (size_t)n * ns;
    double hx[PB200_MAX_PARTICLES][3], hv[PB200_MAX_PARTICLES][3], m[PB200_MAX_PARTICLES], mg[PB200_MAX_PARTICLES];
    const size_t ih = (size_t)P.host * ns + sys;
    for (int b = 0; b < n; b++) {
        size_t i = (size_t)b * ns + sys;
        for (int c = 0; c < 3; c++) {
            hx[b][c] = b == P.host ? 0. : P.pos[i + c * cs] - P.pos[ih + c * cs];
            hv[b][c] = b == P.host ? 0. : P.vel[i + c * cs] - P.vel[ih + c * cs];
        }
        m[b] = P.mass[i]; mg[b] = P.mass_g[i];
    }
    double ekin = 0., epot = 0.;
    for (int b = 0; b < n; b++) ekin += 0.5 * m[b] * (hv[b][0] * hv[b][0] + hv[b][1] * hv[b][1] + hv[b][2] * hv[b][2]);
    for (int a = 0; a < n; a++)
        for (int b = a + 1; b < n; b++) {
            double dx = hx[a][0] - hx[b][0], dy = hx[a][1] - hx[b][1], dz = hx[a][2] - hx[b][2];
            epot -= mg[b] * m[a] / sqrt(dx * dx + dy * dy + dz * dz);
        }
    energy[sys] = ekin + epot;
    double lx = 0., ly = 0., lz = 0.;
    for (int b = 0; b < n; b++) {
        lx += m[b] * (hx[b][1] * hv[b][2] - hx[b][2] * hv[b][1]);
        ly += m[b] * (hx[b][2] * hv[b][0] - hx[b][0] * hv[b][2]);
        lz += m[b] * (hx[b][0] * hv[b][1] - hx[b][1] * hv[b][0]);
    }
    angmom[sys] = sqrt(lx * lx + ly * ly + lz * lz);
}

// SoA history planes -> the reference's 156-byte records (output.rs:119-163), system-major.
// One CTA = 32 consecutive systems x one snapshot: the planes are read with the system index fastest (every load a full
// 256-byte row), the records are assembled in shared memory and written out as each system's contiguous run of
// n_bodies x 156 bytes (the doubles of a record sit at byte 20, hence 4-byte words throughout).
#define PB_REC_WORDS (PB200_HISTORIC_RECORD_BYTES / 4)
__global__ void pack_history_kernel(const __grid_constant__ KParams P, int n_snap, double dt, unsigned int* out) {
    extern __shared__ unsigned int tile[];   // [32 systems][row] words, row = n_bodies * 39 padded to an odd count
    const size_t ns = (size_t)P.n_sys;
    const int n = P.n_bodies;
    const int row = (n * PB_REC_WORDS) | 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const size_t sys = (size_t)blockIdx.x * 32 + lane;
    const int k = blockIdx.y;
    const bool ok = sys < ns;
    const bool have = ok && k < P.hist_count[sys];   // systems that stopped early have fewer snapshots: zero records
    const size_t cs = (size_t)n * ns;
    const double* h = P.hist + (size_t)k * PB_HIST_FIELDS * cs;
    unsigned int* mine = tile + lane * row;
    // (body, plane) pairs over the warps; plane -> word of the record: time | pos, spin, vel, mass, radius, rg2, love, sigma |
    // denergy_dt | lag_angle (record order: ..., sigma, lag_angle, denergy_dt, migration_timescale)
    for (int item = warp; item < n * PB_HIST_FIELDS; item += n_warps) {
        const int b = item / PB_HIST_FIELDS, p = item % PB_HIST_FIELDS;
        const double v = have ? h[(size_t)p * cs + (size_t)b * ns + sys] : 0.;
        const int word = p == 0 ? 0 : (p <= 14 ? 5 + 2 * (p - 1) : (p == 15 ? 5 + 2 * 15 : 5 + 2 * 14));
        const unsigned long long u = (unsigned long long)__double_as_longlong(v);
        mine[b * PB_REC_WORDS + word] = (unsigned int)(u & 0xffffffffull);
        mine[b * PB_REC_WORDS + word + 1] = (unsigned int)(u >> 32);
    }
    for (int b = warp; b < n; b += n_warps) {
        const unsigned long long u = (unsigned long long)__double_as_longlong(dt);
        mine[b * PB_REC_WORDS + 2] = (unsigned int)(u & 0xffffffffull); mine[b * PB_REC_WORDS + 3] = (unsigned int)(u >> 32);   // time_step
        mine[b * PB_REC_WORDS + 4] = (unsigned int)b;                                                                     // particle id (i32)
        mine[b * PB_REC_WORDS + 5 + 2 * 16] = 0u; mine[b * PB_REC_WORDS + 5 + 2 * 16 + 1] = 0u;                            // disk migration_timescale
    }
    __syncthreads();
    // each warp writes whole systems: n * 39 consecutive words
    for (int s = warp; s < 32; s += n_warps) {
        const size_t gs = (size_t)blockIdx.x * 32 + s;
        if (gs >= ns) break;
        unsigned int* dst = out + ((gs * (size_t)n_snap + (size_t)k) * (size_t)n) * PB_REC_WORDS;
        const unsigned int* src = tile + s * row;
        for (int w = lane; w < n * PB_REC_WORDS; w += 32) dst[w] = src[w];
    }
}

// Device-side construction of a perturbed ensemble (pb200_ensemble_create_perturbed): one thread per member.
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long& s) {
    unsigned long long z = (s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
struct PerturbBase { double hpos[PB200_MAX_PARTICLES][3], hvel[PB200_MAX_PARTICLES][3], mass[PB200_MAX_PARTICLES]; };
__global__ void perturb_kernel(const __grid_constant__ KParams P, const __grid_constant__ PerturbBase B, unsigned long long seed, double amp,
                               unsigned long long first_member) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t ns = (size_t)P.n_sys;
    const unsigned long long member = first_member + (unsigned long long)k;   // index in the global ensemble
    if (member == 0 || k >= ns) return;   // member 0 is the base case (already uploaded)
    const int n = P.n_bodies;
    unsigned long long s = seed * 0x100000001b3ull + member;
    sd hp[PB200_MAX_PARTICLES][3], hv[PB200_MAX_PARTICLES][3];
    for (int b = 0; b < n; b++) {
        for (int c = 0; c < 3; c++) { hp[b][c] = sd(B.hpos[b][c]); hv[b][c] = sd(B.hvel[b][c]); }
        if (b == P.host) continue;
        // (2 u - 1) a with u = (z >> 11) 2^-53, then x (1 + delta): every product and sum rounded like the host's doubles
        for (int c = 0; c < 3; c++) {
            const sd u = sd((double)(splitmix64(s) >> 11)) * sd(1.0 / 9007199254740992.0);
            hp[b][c] = hp[b][c] * (sd(1.0) + (sd(2.0) * u - sd(1.0)) * sd(amp));
        }
        for (int c = 0; c < 3; c++) {
            const sd u = sd((double)(splitmix64(s) >> 11)) * sd(1.0 / 9007199254740992.0);
            hv[b][c] = hv[b][c] * (sd(1.0) + (sd(2.0) * u - sd(1.0)) * sd(amp));
        }
    }
    // calculate_center_of_mass (universe.rs:663-697): running pairwise centre of mass in body order
    sd cp[3] = {sd(0.), sd(0.), sd(0.)}, cv[3] = {sd(0.), sd(0.), sd(0.)}, cm = sd(0.);
    for (int b = 0; b < n; b++) {
        const sd m = sd(B.mass[b]);
        for (int c = 0; c < 3; c++) { cp[c] = cp[c] * cm + hp[b][c] * m; cv[c] = cv[c] * cm + hv[b][c] * m; }
        const sd nm = cm + m;
        if (nm.v > 0.) for (int c = 0; c < 3; c++) { cp[c] = div_ieee(cp[c], nm); cv[c] = div_ieee(cv[c], nm); }
        cm = nm;
    }
    const size_t cs = (size_t)n * ns;
    for (int b = 0; b < n; b++)
        for (int c = 0; c < 3; c++) {
            P.pos[(size_t)c * cs + (size_t)b * ns + k] = (hp[b][c] - cp[c]).v;
            P.vel[(size_t)c * cs + (size_t)b * ns + k] = (hv[b][c] - cv[c]).v;
        }
}

// One system's image gathered into a contiguous staging buffer (pb200_ensemble_get_case: one copy instead of ~100):
// per body 27 doubles [pos3 vel3 acc3 L3 spin3 verr3 lerr3 | radius rg2 moi | lag pair_h pair_p], then the Roche table
// (n x n), then t, last_hist, and iteration / n_hist / tswarn as bit patterns.
#define PB_GATHER_PER_BODY 27
__global__ void gather_case_kernel(const __grid_constant__ KParams P, size_t s, const double* roche, double* out) {
    const size_t ns = (size_t)P.n_sys;
    const int n = P.n_bodies;
    const size_t cs = (size_t)n * ns;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int item = threadIdx.x; item < n * PB_GATHER_PER_BODY; item += blockDim.x) {
        const int b = item / PB_GATHER_PER_BODY, f = item % PB_GATHER_PER_BODY;
        const size_t i = (size_t)b * ns + s;
        double v;
        if (f < 21) {
            const double* arr[7] = {P.pos, P.vel, P.acc, P.L, P.spin, P.verr, P.lerr};
            v = arr[f / 3][(size_t)(f % 3) * cs + i];
        } else if (f == 21) v = P.radius[i];
        else if (f == 22) v = P.rg2[i];
        else if (f == 23) v = P.moi[i];
        else if (f == 24) v = (P.flags & FLAG_DYN) ? P.lag[i] : 0.;
        else if (f == 25) v = (P.flags & FLAG_DYN) ? P.pair_h[i] : nan;
        else v = (P.flags & FLAG_DYN) ? P.pair_p[i] : nan;
        out[item] = v;
    }
    double* tail = out + n * PB_GATHER_PER_BODY;
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) tail[i] = roche[(size_t)i * ns + s];
    if (threadIdx.x == 0) {
        double* sys = tail + n * n;
        sys[0] = P.t[s]; sys[1] = P.last_hist[s];
        sys[2] = __longlong_as_double((long long)P.iteration[s]);
        sys[3] = __longlong_as_double((long long)P.n_hist[s]);
        sys[4] = __longlong_as_double((long long)P.tswarn[s]);
    }
}

// DFMA chain microbenchmark: 8 independent chains per thread.
__global__ void dfma_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-9, x1 = x0 + 1., x2 = x0 + 2., x3 = x0 + 3., x4 = x0 + 4., x5 = x0 + 5., x6 = x0 + 6., x7 = x0 + 7.;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

}  // namespace pb200

// ---------------------------------------------------------------------------------------------
template <class T>
static int dev_alloc(pb200_ensemble* e, T** p, size_t count) {
    void* q = nullptr;
    cudaError_t err = cudaMalloc(&q, count * sizeof(T) + 16);
    if (err != cudaSuccess) return set_error(PB200_E_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(err));
    err = cudaMemset(q, 0, count * sizeof(T) + 16);
    if (err != cudaSuccess) return set_error(PB200_E_CUDA, std::string("cudaMemset: ") + cudaGetErrorString(err));
    e->allocations.push_back(q);
    *p = (T*)q;
    return PB200_OK;
}

static bool is_dynamical_tide_evolution(const pb200_body_t& b) {
    return b.evolution_type == PB200_EVO_GALLETBOLMONT2017 || b.evolution_type == PB200_EVO_BOLMONTMATHIS2016 ||
           (b.evolution_type == PB200_EVO_LECONTECHABRIER2013 && b.evolution_parameter != 0.);
}

extern "C" {

const char* pb200_version(void) { return "posidonius_b200 0.1.0 (sm_100a)"; }
const char* pb200_last_error(void) { return g_last_error.c_str(); }

int pb200_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { set_error(PB200_E_CUDA, cudaGetErrorString(e)); return -1; }
    return n;
}

int pb200_case_validate(const pb200_case_t* c, const pb200_table_t* tables, size_t n_tables) {
    if (!c) return set_error(PB200_E_INVALID, "null case");
    const int n = c->n_particles;
    if (n < 2 || n > PB200_MAX_PARTICLES) return set_error(PB200_E_INVALID, "n_particles must be in [2, 10]");
    if (c->coordinates_type < 0 || c->coordinates_type > 2) return set_error(PB200_E_INVALID, "unknown coordinates type");
    if (!(c->time_step > 0.) || c->half_time_step != 0.5 * c->time_step) return set_error(PB200_E_INVALID, "time_step must be > 0 and half_time_step = time_step / 2");
    if (c->consider_disk) return set_error(PB200_E_UNSUPPORTED, "disk interaction is outside the B200 hot path (no CPU fallback)");
    const int h = c->host_most_massive;
    if (h < 0 || h >= n) return set_error(PB200_E_INVALID, "host_most_massive out of range");
    for (int i = 0; i < n; i++) {
        if (i != h && !(c->bodies[i].mass <= c->bodies[h].mass)) return set_error(PB200_E_INVALID, "host_most_massive is not the most massive body");
        if (!(c->bodies[i].moment_of_inertia > 0.)) return set_error(PB200_E_INVALID, "moment of inertia must be > 0 (particles/common.rs:5-7)");
        if (!(c->bodies[i].mass > 0.)) return set_error(PB200_E_INVALID, "mass must be > 0");
    }
    if (c->consider_tides) {
        if (c->host_tides != h) return set_error(PB200_E_UNSUPPORTED, "the tidal host must be the most massive body on the B200 path");
        if (c->bodies[h].tides_role != PB200_ROLE_CENTRAL) return set_error(PB200_E_INVALID, "tides enabled without a CentralBody host");
    }
    if (c->consider_rotational_flattening) {
        if (c->host_rotational_flattening != h) return set_error(PB200_E_UNSUPPORTED, "the rotational-flattening host must be the most massive body on the B200 path");
        if (c->bodies[h].flattening_role != PB200_ROLE_CENTRAL) return set_error(PB200_E_INVALID, "rotational flattening enabled without a CentralBody host");
    }
    if (c->consider_general_relativity) {
        if (c->host_general_relativity != h) return set_error(PB200_E_INVALID, "the central body for General Relativity should be the most massive one (universe.rs:925-927)");
        if (c->bodies[h].general_relativity_role != PB200_ROLE_CENTRAL) return set_error(PB200_E_INVALID, "general relativity enabled without a CentralBody host");
        int g = c->general_relativity_implementation;
        if (g != PB200_GR_KIDDER1995 && g != PB200_GR_ANDERSON1975 && g != PB200_GR_NEWHALL1983)
            return set_error(PB200_E_INVALID, "general relativity enabled with implementation Disabled");
    }
    for (int i = 0; i < n; i++) {
        const pb200_body_t& b = c->bodies[i];
        if (i != h) {
            if (b.tides_role == PB200_ROLE_CENTRAL && c->consider_tides) return set_error(PB200_E_INVALID, "only one central body is allowed for tidal effects");
            if (b.flattening_role == PB200_ROLE_CENTRAL && c->consider_rotational_flattening) return set_error(PB200_E_INVALID, "only one central body is allowed for rotational flattening effects");
            if (b.general_relativity_role == PB200_ROLE_CENTRAL && c->consider_general_relativity) return set_error(PB200_E_INVALID, "only one central body is allowed for general relativity effects");
        }
        if (c->consider_evolution && b.evolution_type != PB200_EVO_NONEVOLVING) {
            if (b.evolution_table < 0 || (size_t)b.evolution_table >= n_tables || !tables)
                return set_error(PB200_E_INVALID, "evolving body without an evolution table");
            const pb200_table_t& t = tables[b.evolution_table];
            if (t.n_rows < 1 || !t.time || !t.radius) return set_error(PB200_E_INVALID, "evolution table needs time and radius columns");
            bool need_rg2 = b.evolution_type == PB200_EVO_BARAFFE2015 || b.evolution_type == PB200_EVO_LECONTE2011 ||
                            b.evolution_type == PB200_EVO_LECONTECHABRIER2013;
            if (need_rg2 && !t.radius_of_gyration_2) return set_error(PB200_E_INVALID, "evolution table needs a radius_of_gyration_2 column");
            if (is_dynamical_tide_evolution(b) && !t.inverse_tidal_q_factor) return set_error(PB200_E_INVALID, "evolution table needs an inverse_tidal_q_factor column");
        }
    }
    // particle ids index the flattened HashMap of pair-dependent dissipation factors: in range and distinct
    {
        bool seen[PB200_MAX_PARTICLES] = {false};
        for (int i = 0; i < n; i++) {
            const int id = c->bodies[i].id;
            if (id < 0 || id >= PB200_MAX_PARTICLES) return set_error(PB200_E_INVALID, "particle id out of range [0, MAX_PARTICLES)");
            if (seen[id]) return set_error(PB200_E_INVALID, "duplicate particle id");
            seen[id] = true;
        }
    }
    // Q4: universe.rs:335-337 reads the host's stale heliocentric velocity; the kernel takes it as zero.
    const double* hv = c->bodies[h].heliocentric_velocity;
    if (hv[0] != 0. || hv[1] != 0. || hv[2] != 0.)
        return set_error(PB200_E_UNSUPPORTED, "the host's heliocentric velocity must be zero (stale-read quirk of universe.rs:335-337 is not reproduced)");
    return PB200_OK;
}

static int same_structure(const pb200_case_t& a, const pb200_case_t& b) {
    if (a.n_particles != b.n_particles || a.coordinates_type != b.coordinates_type || a.time_step != b.time_step ||
        a.time_limit != b.time_limit || a.historic_snapshot_period != b.historic_snapshot_period ||
        a.consider_tides != b.consider_tides || a.consider_rotational_flattening != b.consider_rotational_flattening ||
        a.consider_general_relativity != b.consider_general_relativity || a.consider_evolution != b.consider_evolution ||
        a.consider_wind != b.consider_wind ||
        a.general_relativity_implementation != b.general_relativity_implementation || a.host_most_massive != b.host_most_massive)
        return 0;
    for (int i = 0; i < a.n_particles; i++) {
        const pb200_body_t &x = a.bodies[i], &y = b.bodies[i];
        if (x.tides_role != y.tides_role || x.flattening_role != y.flattening_role ||
            x.general_relativity_role != y.general_relativity_role || x.evolution_type != y.evolution_type ||
            x.evolution_table != y.evolution_table || x.wind_role != y.wind_role || x.evolution_parameter != y.evolution_parameter)
            return 0;
    }
    return 1;
}

void pb200_ensemble_destroy(pb200_ensemble_t* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    for (void* p : e->allocations) cudaFree(p);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

// Everything of KParams that is uniform across the ensemble and known from the case alone (no device involved): geometry,
// effect flags, per-body role masks. Shared by pb200_ensemble_create and pb200_case_step_kernel.
static void fill_uniform_params(KParams& P, const pb200_case_t& c0, size_t n_systems) {
    const int n = c0.n_particles;
    P.n_sys = (int)n_systems; P.n_bodies = n;
    int W = 2; while (W < n) W <<= 1;
    P.W = W; P.shift = 0; while ((1 << P.shift) < W) P.shift++;
    P.host = c0.host_most_massive;
    P.flags = (c0.consider_tides ? FLAG_TIDES : 0) | (c0.consider_rotational_flattening ? FLAG_FLAT : 0) |
              (c0.consider_general_relativity ? FLAG_GR : 0) | (c0.consider_evolution ? FLAG_EVO : 0);
    P.spin_on = c0.consider_tides || c0.consider_rotational_flattening || c0.consider_evolution ||
                (c0.consider_general_relativity && c0.general_relativity_implementation == PB200_GR_KIDDER1995);
    P.dt = c0.time_step; P.half_dt = c0.half_time_step; P.time_limit = c0.time_limit; P.hist_period = c0.historic_snapshot_period;
    P.tides_orbiting = P.flat_orbiting = P.gr_orbiting = P.gr_enabled = 0;
    for (int i = 0; i < n; i++) {
        const pb200_body_t& b = c0.bodies[i];
        if (b.tides_role == PB200_ROLE_ORBITING) P.tides_orbiting |= 1u << i;
        if (b.flattening_role == PB200_ROLE_ORBITING) P.flat_orbiting |= 1u << i;
        if (b.general_relativity_role == PB200_ROLE_ORBITING) P.gr_orbiting |= 1u << i;
        if (b.general_relativity_role != PB200_ROLE_DISABLED) P.gr_enabled |= 1u << i;
        P.evo_table[i] = (c0.consider_evolution && b.evolution_type != PB200_EVO_NONEVOLVING) ? b.evolution_table : -1;
    }
    for (int i = n; i < PB200_MAX_PARTICLES; i++) P.evo_table[i] = -1;
    P.wind_on = P.dyn_evo = 0;
    for (int i = 0; i < n; i++) {
        const pb200_body_t& b = c0.bodies[i];
        if (c0.consider_wind && b.wind_role == 0) P.wind_on |= 1u << i;
        // the reference matches on Particle.evolution whatever consider_effects.evolution says (constant_time_lag.rs:27, 96)
        if (c0.consider_tides && is_dynamical_tide_evolution(b) && (i == P.host || b.tides_role == PB200_ROLE_ORBITING)) P.dyn_evo |= 1u << i;
    }
    if (P.wind_on) P.flags |= FLAG_WIND;
    if (P.dyn_evo) P.flags |= FLAG_DYN;
    P.tides_host_central = c0.bodies[P.host].tides_role == PB200_ROLE_CENTRAL;
    P.flat_host_central = c0.bodies[P.host].flattening_role == PB200_ROLE_CENTRAL;
}

int pb200_ensemble_create(const pb200_case_t* cases, size_t n_cases, size_t n_systems, const pb200_table_t* tables,
                          size_t n_tables, int device, pb200_ensemble_t** out) {
    if (!cases || !out || n_systems == 0) return set_error(PB200_E_INVALID, "null argument or empty ensemble");
    if (n_cases != 1 && n_cases != n_systems) return set_error(PB200_E_INVALID, "n_cases must be 1 or n_systems");
    if (n_tables > PB200_MAX_PARTICLES) return set_error(PB200_E_INVALID, "too many evolution tables");
    int rc = pb200_case_validate(&cases[0], tables, n_tables);
    if (rc != PB200_OK) return rc;
    for (size_t s = 1; s < n_cases; s++) {
        if (!same_structure(cases[0], cases[s])) return set_error(PB200_E_INVALID, "all cases of an ensemble must share structure (bodies, coordinates, effects, dt, limits)");
        rc = pb200_case_validate(&cases[s], tables, n_tables);
        if (rc != PB200_OK) return rc;
    }
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return set_error(PB200_E_INVALID, "device index out of range");
    CUDA_TRY(cudaSetDevice(device));
    pb200_ensemble* e = new pb200_ensemble();
    e->device = device;
    e->n_sys = n_systems;
    const pb200_case_t& c0 = cases[0];
    const int n = c0.n_particles;
    e->n_bodies = n;
    e->tmpl = c0;
    e->cases.assign(cases, cases + n_cases);
    e->coord = c0.coordinates_type;
    e->gr = c0.consider_general_relativity ? c0.general_relativity_implementation : PB200_GR_DISABLED;
    e->recovery_snapshot_period = c0.recovery_snapshot_period;
    e->clock_t = c0.current_time; e->clock_last_hist = c0.last_historic_snapshot_time;
    { const char* fg = getenv("PB200_FORCE_GENERIC"); e->force_generic = fg && fg[0] == '1'; }
    { const char* nb = getenv("PB200_NARROW_BLOCKS"); e->narrow_blocks = nb && nb[0] == '1'; }
    { const char* pl = getenv("PB200_PAIR_LANES"); e->pair_lanes = pl && pl[0] == '1'; }
    if (cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) e->sm_count = 0;
    for (size_t s = 1; s < n_cases; s++)
        if (cases[s].current_time != c0.current_time || cases[s].last_historic_snapshot_time != c0.last_historic_snapshot_time) e->uniform_clock = false;
    KParams& P = e->P;
    fill_uniform_params(P, c0, n_systems);

#define TRY(x) do { int _r = (x); if (_r != PB200_OK) { pb200_ensemble_destroy(e); return _r; } } while (0)
#define CUDA_TRY_E(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { pb200_ensemble_destroy(e); return set_error(PB200_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } } while (0)
    CUDA_TRY_E(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    CUDA_TRY_E(cudaEventCreate(&e->ev0));
    CUDA_TRY_E(cudaEventCreate(&e->ev1));
    // evolution tables (replicated per GPU, read through the read-only path)
    for (int i = 0; i < PB200_MAX_PARTICLES; i++) P.tables[i] = DevTable{nullptr, nullptr, nullptr, 0, 0, 0, nullptr};
    // Every column a table provides is uploaded, so that bodies of different EvolutionTypes may share one table; which
    // columns a body interpolates is decided per body (evo_rg2 / dyn_evo masks; pb200_case_validate checked the columns).
    P.evo_rg2 = 0;
    for (int i = 0; i < n; i++) {
        int ti = P.evo_table[i];
        if (ti < 0) continue;
        const pb200_body_t& b = c0.bodies[i];
        if (b.evolution_type == PB200_EVO_BARAFFE2015 || b.evolution_type == PB200_EVO_LECONTE2011 || b.evolution_type == PB200_EVO_LECONTECHABRIER2013)
            P.evo_rg2 |= 1u << i;
        if (P.tables[ti].time) continue;
        const pb200_table_t& t = tables[ti];
        double *dt_ = nullptr, *dr = nullptr, *dg = nullptr, *dq = nullptr;
        TRY(dev_alloc(e, &dt_, t.n_rows));
        TRY(dev_alloc(e, &dr, t.n_rows));
        CUDA_TRY_E(cudaMemcpy(dt_, t.time, t.n_rows * sizeof(double), cudaMemcpyHostToDevice));
        CUDA_TRY_E(cudaMemcpy(dr, t.radius, t.n_rows * sizeof(double), cudaMemcpyHostToDevice));
        if (t.radius_of_gyration_2) {
            TRY(dev_alloc(e, &dg, t.n_rows));
            CUDA_TRY_E(cudaMemcpy(dg, t.radius_of_gyration_2, t.n_rows * sizeof(double), cudaMemcpyHostToDevice));
        }
        if (t.inverse_tidal_q_factor) {
            TRY(dev_alloc(e, &dq, t.n_rows));
            CUDA_TRY_E(cudaMemcpy(dq, t.inverse_tidal_q_factor, t.n_rows * sizeof(double), cudaMemcpyHostToDevice));
        }
        P.tables[ti] = DevTable{dt_, dr, dg, (int)t.n_rows, 1, 0, dq};
    }
    const size_t ns = n_systems, nb = (size_t)n;
    TRY(dev_alloc(e, &P.pos, 3 * nb * ns)); TRY(dev_alloc(e, &P.vel, 3 * nb * ns)); TRY(dev_alloc(e, &P.acc, 3 * nb * ns));
    TRY(dev_alloc(e, &P.L, 3 * nb * ns)); TRY(dev_alloc(e, &P.spin, 3 * nb * ns)); TRY(dev_alloc(e, &P.verr, 3 * nb * ns));
    TRY(dev_alloc(e, &P.lerr, 3 * nb * ns));
    TRY(dev_alloc(e, &P.radius, nb * ns)); TRY(dev_alloc(e, &P.rg2, nb * ns)); TRY(dev_alloc(e, &P.moi, nb * ns));
    TRY(dev_alloc(e, &e->d_mass, nb * ns)); TRY(dev_alloc(e, &e->d_mass_g, nb * ns)); TRY(dev_alloc(e, &e->d_sigma, nb * ns));
    TRY(dev_alloc(e, &e->d_k2t, nb * ns)); TRY(dev_alloc(e, &e->d_k2f, nb * ns)); TRY(dev_alloc(e, &e->d_roche, nb * nb * ns));
    P.mass = e->d_mass; P.mass_g = e->d_mass_g; P.sigma = e->d_sigma; P.k2t = e->d_k2t; P.k2f = e->d_k2f; P.roche = e->d_roche;
    TRY(dev_alloc(e, &P.t, ns)); TRY(dev_alloc(e, &P.last_hist, ns));
    TRY(dev_alloc(e, &P.iteration, ns)); TRY(dev_alloc(e, &P.n_hist, ns)); TRY(dev_alloc(e, &P.event_iteration, ns));
    TRY(dev_alloc(e, &P.tswarn, ns)); TRY(dev_alloc(e, &P.status, ns)); TRY(dev_alloc(e, &P.warnings, ns));
    TRY(dev_alloc(e, &P.hist_count, ns));
    TRY(dev_alloc(e, &P.tide_scratch, (size_t)PB_TIDE_SCRATCH * nb * ns));
    TRY(dev_alloc(e, &e->d_energy, ns)); TRY(dev_alloc(e, &e->d_angmom, ns));
    TRY(dev_alloc(e, &P.sched, (ns * (size_t)P.W + 31) / 32 + 2));   // one flag per CTA of the smallest build (32 threads) + the ticket counter
    P.n_groups = 0; P.n_pieces = 1;
    if (P.flags & FLAG_WIND) { TRY(dev_alloc(e, &e->d_wind_k, nb * ns)); TRY(dev_alloc(e, &e->d_wind_sat, nb * ns)); }
    if (P.flags & FLAG_DYN) {
        TRY(dev_alloc(e, &e->d_diss, nb * ns)); TRY(dev_alloc(e, &e->d_diss_scale, nb * ns)); TRY(dev_alloc(e, &P.lag, nb * ns));
        TRY(dev_alloc(e, &P.pair_h, nb * ns)); TRY(dev_alloc(e, &P.pair_p, nb * ns));
    }
    P.wind_k = e->d_wind_k; P.wind_sat = e->d_wind_sat; P.diss = e->d_diss; P.diss_scale = e->d_diss_scale;
    // history planes: keep the buffer below ~256 MB
    {
        size_t per_slot = (size_t)PB_HIST_FIELDS * nb * ns * sizeof(double);
        size_t slots = (256ull << 20) / per_slot;
        if (slots < 2) slots = 2;
        if (slots > 64) slots = 64;
        if (const char* hs = getenv("PB200_HISTORY_SLOTS")) { int v = atoi(hs); if (v >= 2 && (size_t)v < slots) slots = (size_t)v; }   // tests: a small buffer
        P.hist_capacity = (int)slots;
        TRY(dev_alloc(e, &P.hist, (size_t)PB_HIST_FIELDS * nb * ns * slots));
    }
    // pack the SoA on the host and upload
    {
        std::vector<double> buf(3 * nb * ns);
        auto up3 = [&](double* dst, auto get) -> int {
            for (size_t s = 0; s < ns; s++) {
                const pb200_case_t& cs_ = cases[n_cases == 1 ? 0 : s];
                for (size_t b = 0; b < nb; b++) {
                    const double* v = get(cs_, (int)b);
                    for (int c = 0; c < 3; c++) buf[((size_t)c * nb + b) * ns + s] = v[c];
                }
            }
            cudaError_t err = cudaMemcpy(dst, buf.data(), 3 * nb * ns * sizeof(double), cudaMemcpyHostToDevice);
            return err == cudaSuccess ? PB200_OK : set_error(PB200_E_CUDA, cudaGetErrorString(err));
        };
        auto up1 = [&](double* dst, auto get) -> int {
            for (size_t s = 0; s < ns; s++) {
                const pb200_case_t& cs_ = cases[n_cases == 1 ? 0 : s];
                for (size_t b = 0; b < nb; b++) buf[b * ns + s] = get(cs_, (int)b);
            }
            cudaError_t err = cudaMemcpy(dst, buf.data(), nb * ns * sizeof(double), cudaMemcpyHostToDevice);
            return err == cudaSuccess ? PB200_OK : set_error(PB200_E_CUDA, cudaGetErrorString(err));
        };
        TRY(up3(P.pos, [](const pb200_case_t& c, int b) { return c.bodies[b].inertial_position; }));
        TRY(up3(P.vel, [](const pb200_case_t& c, int b) { return c.bodies[b].inertial_velocity; }));
        TRY(up3(P.acc, [](const pb200_case_t& c, int b) { return c.bodies[b].inertial_acceleration; }));
        TRY(up3(P.L, [](const pb200_case_t& c, int b) { return c.bodies[b].angular_momentum; }));
        TRY(up3(P.spin, [](const pb200_case_t& c, int b) { return c.bodies[b].spin; }));
        TRY(up3(P.verr, [](const pb200_case_t& c, int b) { return (const double*)c.inertial_velocity_errors[b]; }));
        TRY(up3(P.lerr, [](const pb200_case_t& c, int b) { return (const double*)c.particle_angular_momentum_errors[b]; }));
        TRY(up1(P.radius, [](const pb200_case_t& c, int b) { return c.bodies[b].radius; }));
        TRY(up1(P.rg2, [](const pb200_case_t& c, int b) { return c.bodies[b].radius_of_gyration_2; }));
        TRY(up1(P.moi, [](const pb200_case_t& c, int b) { return c.bodies[b].moment_of_inertia; }));
        TRY(up1(e->d_mass, [](const pb200_case_t& c, int b) { return c.bodies[b].mass; }));
        TRY(up1(e->d_mass_g, [](const pb200_case_t& c, int b) { return c.bodies[b].mass_g; }));
        TRY(up1(e->d_sigma, [](const pb200_case_t& c, int b) { return c.bodies[b].tides_scaled_dissipation_factor; }));
        TRY(up1(e->d_k2t, [](const pb200_case_t& c, int b) { return c.bodies[b].tides_role != PB200_ROLE_DISABLED ? c.bodies[b].tides_love_number : 0.; }));
        TRY(up1(e->d_k2f, [](const pb200_case_t& c, int b) { return c.bodies[b].flattening_role != PB200_ROLE_DISABLED ? c.bodies[b].flattening_love_number : 0.; }));
        if (P.flags & FLAG_WIND) {
            TRY(up1(e->d_wind_k, [](const pb200_case_t& c, int b) { return c.bodies[b].wind_k_factor; }));
            TRY(up1(e->d_wind_sat, [](const pb200_case_t& c, int b) { return c.bodies[b].wind_rotation_saturation; }));
        }
        if (P.flags & FLAG_DYN) {
            const int host = P.host;
            TRY(up1(e->d_diss, [](const pb200_case_t& c, int b) { return c.bodies[b].tides_dissipation_factor; }));
            TRY(up1(e->d_diss_scale, [](const pb200_case_t& c, int b) { return c.bodies[b].tides_dissipation_factor_scale; }));
            TRY(up1(P.lag, [](const pb200_case_t& c, int b) { return c.bodies[b].tides_lag_angle; }));
            // the HashMap entries (host, b) and (b, host), keyed by particle id (constant_time_lag.rs:152-160)
            TRY(up1(P.pair_h, [host](const pb200_case_t& c, int b) { return c.pair_dependent_scaled_dissipation_factor[c.bodies[host].id * PB200_MAX_PARTICLES + c.bodies[b].id]; }));
            TRY(up1(P.pair_p, [host](const pb200_case_t& c, int b) { return c.pair_dependent_scaled_dissipation_factor[c.bodies[b].id * PB200_MAX_PARTICLES + c.bodies[host].id]; }));
        }
        // roche [i][j][s]
        {
            std::vector<double> rb(nb * nb * ns);
            for (size_t s = 0; s < ns; s++) {
                const pb200_case_t& cs_ = cases[n_cases == 1 ? 0 : s];
                for (size_t i = 0; i < nb * nb; i++) rb[i * ns + s] = cs_.roche_radiuses[i];
            }
            CUDA_TRY_E(cudaMemcpy(e->d_roche, rb.data(), rb.size() * sizeof(double), cudaMemcpyHostToDevice));
        }
        std::vector<double> t(ns), lh(ns);
        std::vector<unsigned long long> it(ns), nh(ns), tw(ns);
        for (size_t s = 0; s < ns; s++) {
            const pb200_case_t& cs_ = cases[n_cases == 1 ? 0 : s];
            t[s] = cs_.current_time; lh[s] = cs_.last_historic_snapshot_time; it[s] = cs_.current_iteration;
            nh[s] = cs_.n_historic_snapshots; tw[s] = cs_.timestep_warning;
        }
        CUDA_TRY_E(cudaMemcpy(P.t, t.data(), ns * sizeof(double), cudaMemcpyHostToDevice));
        CUDA_TRY_E(cudaMemcpy(P.last_hist, lh.data(), ns * sizeof(double), cudaMemcpyHostToDevice));
        CUDA_TRY_E(cudaMemcpy(P.iteration, it.data(), ns * sizeof(unsigned long long), cudaMemcpyHostToDevice));
        CUDA_TRY_E(cudaMemcpy(P.n_hist, nh.data(), ns * sizeof(unsigned long long), cudaMemcpyHostToDevice));
        CUDA_TRY_E(cudaMemcpy(P.tswarn, tw.data(), ns * sizeof(unsigned long long), cudaMemcpyHostToDevice));
    }
#undef TRY
#undef CUDA_TRY_E
    *out = e;
    return PB200_OK;
}

int pb200_ensemble_create_perturbed(const pb200_case_t* base, size_t n_systems, uint64_t seed, double amplitude,
                                    const pb200_table_t* tables, size_t n_tables, int device, pb200_ensemble_t** out) {
    return pb200_ensemble_create_perturbed_range(base, 0, n_systems, seed, amplitude, tables, n_tables, device, out);
}

int pb200_ensemble_create_perturbed_range(const pb200_case_t* base, uint64_t first_member, size_t n_systems, uint64_t seed, double amplitude,
                                          const pb200_table_t* tables, size_t n_tables, int device, pb200_ensemble_t** out) {
    if (!base || !out) return set_error(PB200_E_INVALID, "null argument");
    pb200_ensemble_t* e = nullptr;
    int rc = pb200_ensemble_create(base, 1, n_systems, tables, n_tables, device, &e);   // every member starts as the base case
    if (rc != PB200_OK) return rc;
    PerturbBase B;
    std::memset(&B, 0, sizeof B);
    for (int b = 0; b < base->n_particles; b++) {
        for (int c = 0; c < 3; c++) { B.hpos[b][c] = base->bodies[b].heliocentric_position[c]; B.hvel[b][c] = base->bodies[b].heliocentric_velocity[c]; }
        B.mass[b] = base->bodies[b].mass;
    }
    perturb_kernel<<<(unsigned)((n_systems + 127) / 128), 128, 0, e->stream>>>(e->P, B, (unsigned long long)seed, amplitude,
                                                                                 (unsigned long long)first_member);
    e->launches++;
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
    if (err != cudaSuccess) { pb200_ensemble_destroy(e); return set_error(PB200_E_CUDA, std::string("perturb kernel: ") + cudaGetErrorString(err)); }
    e->perturbed = true;
    *out = e;
    return PB200_OK;
}

int pb200_ensemble_n_particles(const pb200_ensemble_t* e) { return e ? e->n_bodies : 0; }
size_t pb200_ensemble_n_systems(const pb200_ensemble_t* e) { return e ? e->n_sys : 0; }

int pb200_ensemble_set_time_limit(pb200_ensemble_t* e, double time_limit) {
    if (!e) return set_error(PB200_E_INVALID, "null ensemble");
    // whfast.rs:187-208: ignored unless > 0 and different
    if (time_limit > 0. && e->P.time_limit != time_limit) { e->P.time_limit = time_limit; e->tmpl.time_limit = time_limit; }
    return PB200_OK;
}
int pb200_ensemble_set_snapshot_periods(pb200_ensemble_t* e, double historic, double recovery) {
    if (!e) return set_error(PB200_E_INVALID, "null ensemble");
    if (historic > 0. && e->P.hist_period != historic) { e->P.hist_period = historic; e->tmpl.historic_snapshot_period = historic; }
    if (recovery > 0. && e->recovery_snapshot_period != recovery) { e->recovery_snapshot_period = recovery; e->tmpl.recovery_snapshot_period = recovery; }
    return PB200_OK;
}

int pb200_ensemble_set_arithmetic(pb200_ensemble_t* e, int mode) {
    if (!e) return set_error(PB200_E_INVALID, "null ensemble");
    if (mode != PB200_ARITH_FAST && mode != PB200_ARITH_STRICT && mode != PB200_ARITH_HYBRID) return set_error(PB200_E_INVALID, "unknown arithmetic mode");
    e->arithmetic = mode;
    return PB200_OK;
}

int pb200_ensemble_initialize_physical_values(pb200_ensemble_t* e) {
    if (!e) return set_error(PB200_E_INVALID, "null ensemble");
    CUDA_TRY(cudaSetDevice(e->device));
    for (const auto& c : e->cases)
        if (c.current_time != 0.) return set_error(PB200_E_INVALID, "Physical values cannot be initialized on a resumed simulation (whfast.rs:227-229)");
    size_t total = e->n_sys * (size_t)e->n_bodies;
    init_physical_kernel<<<(unsigned)((total + 127) / 128), 128, 0, e->stream>>>(e->P);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    return PB200_OK;
}

// Re-reads the per-system clocks after an upload of current_time (pb200_ensemble_upload / run_host): the host mirror that
// counts upcoming snapshots is valid only while every live system shares one clock.
static int refresh_clock_mirror(pb200_ensemble* e) {
    const size_t ns = e->n_sys;
    std::vector<double> t(ns), lh(ns);
    std::vector<int> hc(ns), st(ns);
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    CUDA_TRY(cudaMemcpy(t.data(), e->P.t, ns * sizeof(double), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(lh.data(), e->P.last_hist, ns * sizeof(double), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(hc.data(), e->P.hist_count, ns * sizeof(int), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(st.data(), e->P.status, ns * sizeof(int), cudaMemcpyDeviceToHost));
    bool uniform = true, have = false;
    double t0 = 0., lh0 = -1.;
    int pending = 0;
    for (size_t s = 0; s < ns; s++) {
        if (hc[s] > pending) pending = hc[s];
        if (st[s] != PB200_STATUS_OK) continue;   // stopped systems take no further snapshots
        if (!have) { t0 = t[s]; lh0 = lh[s]; have = true; }
        else if (t[s] != t0 || lh[s] != lh0) uniform = false;
    }
    e->uniform_clock = uniform;
    e->clock_t = t0; e->clock_last_hist = lh0;
    e->hist_pending_host = (size_t)pending;
    e->clock_dirty = false;
    return PB200_OK;
}

// ---- Which build of the step kernel integrates an ensemble (geometry builds). A pure function of the case's
// uniform parameters, the ensemble size and the SM count, so that it can be asked without a device (pb200_case_step_kernel).
enum StepBuild { BUILD_GENERIC, BUILD_N8, BUILD_N8W, BUILD_S2, BUILD_S2T, BUILD_S2ANY, BUILD_S3, BUILD_S3E, BUILD_S3ANY, BUILD_S3J, BUILD_S3P, BUILD_S3JANY };
static const char* build_name(StepBuild b) {
    static const char* const names[] = {"generic", "n8", "n8w", "s2", "s2t", "s2any", "s3", "s3e", "s3any", "s3j", "s3p", "s3jany"};
    return names[b];
}
static StepBuild select_build(const KParams& P, int coord, int gr, int arithmetic, size_t n_sys, int sm_count, bool force_generic, bool narrow_blocks,
                              bool pair_lanes) {
    // compile-time geometry builds: host at index 0
    const bool fixed_ok = P.host == 0 && !force_generic;
    const bool dh = coord == PB200_COORD_DEMOCRATIC_HELIOCENTRIC, kidder = gr == PB200_GR_KIDDER1995;
    const int tfg = FLAG_TIDES | FLAG_FLAT | FLAG_GR, flags = P.flags;
    // 8 bodies, tides + flattening + Kidder1995, democratic heliocentric (config 4): 384-thread CTAs once there is at least one of
    // them per SM (below that the 64-thread build spreads the work over more SMs; the all-exact arithmetic gains nothing from
    // them: 2.36e8 against 2.49e8 system-steps/s)
    if (fixed_ok && P.n_bodies == 8 && dh && kidder && flags == tfg)
        return (n_sys * 8 >= (size_t)384 * (size_t)sm_count && !narrow_blocks && arithmetic != PB200_ARITH_STRICT) ? BUILD_N8W : BUILD_N8;
    // 2 and 3 bodies: lane = planet (small_step.cuh) — host 0, democratic heliocentric (or Jacobi with 3 bodies), spin
    // integrated, any subset of tides / flattening / GR Kidder1995 / evolution tables, no wind, no dynamical tides. The BASELINE
    // configurations (1, 2, 3, 3-evolving, 5) have compile-time effect sets; every other subset takes the catch-all build of
    // its geometry (flag word read at run time). PB200_FORCE_GENERIC=1 gives the lane = body kernel.
    const bool small_ok = fixed_ok && P.spin_on && (flags & ~(tfg | FLAG_EVO)) == 0 && (!(flags & FLAG_GR) || kidder);
    if (small_ok && P.n_bodies == 2 && dh) return flags == tfg ? BUILD_S2 : flags == FLAG_TIDES ? BUILD_S2T : BUILD_S2ANY;
    if (small_ok && P.n_bodies == 3 && dh) return flags == tfg ? BUILD_S3 : flags == (tfg | FLAG_EVO) ? BUILD_S3E : BUILD_S3ANY;
    if (small_ok && P.n_bodies == 3 && coord == PB200_COORD_JACOBI) {
        if (flags != (tfg | FLAG_EVO)) return BUILD_S3JANY;
        // body 2 an OrbitingBody of no effect (the circumbinary planet of config 5): it rides in body 1's thread once the ensemble
        // fills the GPU that way (six warps per SM as one 192-thread CTA); a smaller ensemble keeps two lanes per system (twice the
        // warps: 8192 members 4.4e8 against 4.2e8 system-steps/s)
        const bool passive = (((P.tides_orbiting | P.flat_orbiting | P.gr_orbiting) >> 2) & 1u) == 0;
        return passive && !pair_lanes && 4 * n_sys >= 3 * (size_t)192 * (size_t)sm_count ? BUILD_S3P : BUILD_S3J;
    }
    return BUILD_GENERIC;
}

const char* pb200_case_step_kernel(const pb200_case_t* c, size_t n_systems, int sm_count, int arithmetic) {
    if (!c || c->n_particles < 2 || c->n_particles > PB200_MAX_PARTICLES) return "";
    KParams P{};
    fill_uniform_params(P, *c, n_systems);
    const int gr = c->consider_general_relativity ? c->general_relativity_implementation : PB200_GR_DISABLED;
    return build_name(select_build(P, c->coordinates_type, gr, arithmetic, n_systems, sm_count, false, false, false));
}

int pb200_ensemble_step(pb200_ensemble_t* e, uint64_t n_steps) {
    if (!e) return set_error(PB200_E_INVALID, "null ensemble");
    if (n_steps == 0) return PB200_OK;
    CUDA_TRY(cudaSetDevice(e->device));
    const size_t threads = e->n_sys * (size_t)e->P.W;
    // The history planes hold hist_capacity snapshots; refuse a call that could overflow them (the caller drains and
    // steps in smaller calls: pb200_ensemble_history_capacity). The count is exact when the ensemble shares one clock
    // (host mirror, advanced with the kernel's own roundings), else a bound from a device query.
    if (e->clock_dirty) {
        int rc = refresh_clock_mirror(e);
        if (rc != PB200_OK) return rc;
    }
    {
        size_t upcoming;
        if (e->uniform_clock) {
            upcoming = 0;
            double t = e->clock_t, lh = e->clock_last_hist;
            for (uint64_t k = 0; k < n_steps; k++) {
                bool first = lh < 0.;
                if (first || lh + e->P.hist_period <= t) { upcoming++; if (!first) lh += e->P.hist_period; else lh = 0.; }
                t += e->P.dt;
                if (t + e->P.dt > e->P.time_limit) break;
            }
            if (e->hist_pending_host + upcoming > (size_t)e->P.hist_capacity)
                return set_error(PB200_E_INVALID, "history buffer would overflow: call pb200_ensemble_history_drain and step in smaller calls");
            e->clock_t = t; e->clock_last_hist = lh;
            e->hist_pending_host += upcoming;
        } else {
            upcoming = (size_t)std::floor((double)n_steps * e->P.dt / e->P.hist_period) + 2;
            if (upcoming > n_steps) upcoming = (size_t)n_steps;
            if (pb200_ensemble_history_pending(e) + upcoming > (size_t)e->P.hist_capacity)
                return set_error(PB200_E_INVALID, "history buffer would overflow: call pb200_ensemble_history_drain and step in smaller calls");
        }
    }
    CUDA_TRY(cudaEventRecord(e->ev0, e->stream));
    {
        const StepBuild build = select_build(e->P, e->coord, e->gr, e->arithmetic, e->n_sys, e->sm_count, e->force_generic, e->narrow_blocks, e->pair_lanes);
        e->last_kernel = build_name(build);
        cudaError_t err;
        switch (build) {
            case BUILD_N8W: err = pb200_launch_n8w(e, threads, n_steps); break;
            case BUILD_N8: err = pb200_launch_n8(e, threads, n_steps); break;
            case BUILD_S2: err = pb200_launch_s2(e, n_steps); break;
            case BUILD_S2T: err = pb200_launch_s2t(e, n_steps); break;
            case BUILD_S2ANY: err = pb200_launch_s2any(e, n_steps); break;
            case BUILD_S3: err = pb200_launch_s3(e, n_steps); break;
            case BUILD_S3E: err = pb200_launch_s3e(e, n_steps); break;
            case BUILD_S3ANY: err = pb200_launch_s3any(e, n_steps); break;
            case BUILD_S3J: err = pb200_launch_s3j(e, n_steps); break;
            case BUILD_S3P: err = pb200_launch_s3p(e, n_steps); break;
            case BUILD_S3JANY: err = pb200_launch_s3jany(e, n_steps); break;
            default:
                err = e->arithmetic == PB200_ARITH_FAST ? pb200_launch_generic_fast(e, threads, n_steps)
                      : e->arithmetic == PB200_ARITH_STRICT ? pb200_launch_generic_strict(e, threads, n_steps) : pb200_launch_generic_hybrid(e, threads, n_steps);
        }
        e->launches++;
        if (err != cudaSuccess) return set_error(PB200_E_CUDA, std::string("step kernel launch: ") + cudaGetErrorString(err));
    }
    CUDA_TRY(cudaEventRecord(e->ev1, e->stream));
    e->timing_pending = true;
    return PB200_OK;
}

int pb200_ensemble_synchronize(pb200_ensemble_t* e) {
    if (!e) return set_error(PB200_E_INVALID, "null ensemble");
    CUDA_TRY(cudaSetDevice(e->device));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return PB200_OK;
}

int pb200_ensemble_last_step_ms(pb200_ensemble_t* e, float* ms) {
    if (!e || !ms) return set_error(PB200_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(e->device));
    if (e->timing_pending) {
        CUDA_TRY(cudaEventSynchronize(e->ev1));
        CUDA_TRY(cudaEventElapsedTime(&e->last_ms, e->ev0, e->ev1));
        e->timing_pending = false;
    }
    *ms = e->last_ms;
    return PB200_OK;
}

uint64_t pb200_ensemble_launch_count(const pb200_ensemble_t* e) { return e ? e->launches : 0; }
unsigned pb200_ensemble_last_pieces(const pb200_ensemble_t* e) { return e ? e->last_pieces : 0; }
const char* pb200_ensemble_last_kernel(const pb200_ensemble_t* e) { return e ? e->last_kernel : ""; }
size_t pb200_ensemble_history_capacity(const pb200_ensemble_t* e) { return e ? (size_t)e->P.hist_capacity : 0; }

int pb200_ensemble_status(pb200_ensemble_t* e, int32_t* status, uint32_t* warnings, uint64_t* iteration_of_event) {
    if (!e) return set_error(PB200_E_INVALID, "null ensemble");
    CUDA_TRY(cudaSetDevice(e->device));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    if (status) CUDA_TRY(cudaMemcpy(status, e->P.status, e->n_sys * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (warnings) CUDA_TRY(cudaMemcpy(warnings, e->P.warnings, e->n_sys * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (iteration_of_event) CUDA_TRY(cudaMemcpy(iteration_of_event, e->P.event_iteration, e->n_sys * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return PB200_OK;
}

static int copy_state(pb200_ensemble* e, const pb200_state_view_t* v, bool to_host) {
    const size_t ns = e->n_sys, nb = (size_t)e->n_bodies;
    struct Item { double* host; double* dev; size_t count; };
    Item items[] = {
        {v->position, e->P.pos, 3 * nb * ns}, {v->velocity, e->P.vel, 3 * nb * ns}, {v->acceleration, e->P.acc, 3 * nb * ns},
        {v->angular_momentum, e->P.L, 3 * nb * ns}, {v->spin, e->P.spin, 3 * nb * ns}, {v->velocity_errors, e->P.verr, 3 * nb * ns},
        {v->angular_momentum_errors, e->P.lerr, 3 * nb * ns}, {v->radius, e->P.radius, nb * ns},
        {v->radius_of_gyration_2, e->P.rg2, nb * ns}, {v->moment_of_inertia, e->P.moi, nb * ns}, {v->current_time, e->P.t, ns},
    };
    if (!to_host && v->current_time) e->clock_dirty = true;
    for (const Item& it : items) {
        if (!it.host) continue;
        cudaError_t err = to_host ? cudaMemcpyAsync(it.host, it.dev, it.count * sizeof(double), cudaMemcpyDeviceToHost, e->stream)
                                  : cudaMemcpyAsync(it.dev, it.host, it.count * sizeof(double), cudaMemcpyHostToDevice, e->stream);
        if (err != cudaSuccess) return set_error(PB200_E_CUDA, std::string("state copy: ") + cudaGetErrorString(err));
    }
    return PB200_OK;
}

int pb200_ensemble_download(pb200_ensemble_t* e, const pb200_state_view_t* dst) {
    if (!e || !dst) return set_error(PB200_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(e->device));
    int rc = copy_state(e, dst, true);
    if (rc != PB200_OK) return rc;
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return PB200_OK;
}

int pb200_ensemble_upload(pb200_ensemble_t* e, const pb200_state_view_t* src) {
    if (!e || !src) return set_error(PB200_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(e->device));
    int rc = copy_state(e, src, false);
    if (rc != PB200_OK) return rc;
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return PB200_OK;
}

int pb200_ensemble_run_host(pb200_ensemble_t* e, const pb200_state_view_t* io, uint64_t n_steps) {
    if (!e || !io) return set_error(PB200_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(e->device));
    int rc = copy_state(e, io, false);
    if (rc != PB200_OK) return rc;
    rc = pb200_ensemble_step(e, n_steps);
    if (rc != PB200_OK) return rc;
    rc = copy_state(e, io, true);
    if (rc != PB200_OK) return rc;
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return PB200_OK;
}

int pb200_ensemble_get_case(pb200_ensemble_t* e, size_t s, pb200_case_t* out) {
    if (!e || !out || s >= e->n_sys) return set_error(PB200_E_INVALID, "bad argument");
    CUDA_TRY(cudaSetDevice(e->device));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    *out = e->cases.size() == 1 ? e->cases[0] : e->cases[s];
    out->time_limit = e->P.time_limit;
    out->historic_snapshot_period = e->P.hist_period;
    out->recovery_snapshot_period = e->recovery_snapshot_period;
    const size_t nb = (size_t)e->n_bodies;
    const size_t count = nb * PB_GATHER_PER_BODY + nb * nb + 5;
    if (!e->d_gather) { int rc = dev_alloc(e, &e->d_gather, (size_t)PB200_MAX_PARTICLES * PB_GATHER_PER_BODY + PB200_MAX_PARTICLES * PB200_MAX_PARTICLES + 5); if (rc != PB200_OK) return rc; }
    gather_case_kernel<<<1, 128, 0, e->stream>>>(e->P, s, e->d_roche, e->d_gather);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    std::vector<double> g(count);
    CUDA_TRY(cudaMemcpyAsync(g.data(), e->d_gather, count * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    for (size_t b = 0; b < nb; b++) {
        pb200_body_t& B = out->bodies[b];
        const double* v = g.data() + b * PB_GATHER_PER_BODY;
        for (int c = 0; c < 3; c++) {
            B.inertial_position[c] = v[c]; B.inertial_velocity[c] = v[3 + c]; B.inertial_acceleration[c] = v[6 + c];
            B.angular_momentum[c] = v[9 + c]; B.spin[c] = v[12 + c];
            out->inertial_velocity_errors[b][c] = v[15 + c]; out->particle_angular_momentum_errors[b][c] = v[18 + c];
        }
        B.radius = v[21]; B.radius_of_gyration_2 = v[22]; B.moment_of_inertia = v[23];
        if (e->P.flags & FLAG_DYN) {
            B.tides_lag_angle = v[24];
            const int hid = out->bodies[e->P.host].id;
            if ((int)b != e->P.host && B.id >= 0 && B.id < PB200_MAX_PARTICLES && hid >= 0 && hid < PB200_MAX_PARTICLES) {
                out->pair_dependent_scaled_dissipation_factor[hid * PB200_MAX_PARTICLES + B.id] = v[25];
                out->pair_dependent_scaled_dissipation_factor[B.id * PB200_MAX_PARTICLES + hid] = v[26];
            }
        }
    }
    if (e->perturbed) {
        // members built on the device have no host image of their own: heliocentric coordinates as inertial_to_heliocentric
        // (universe.rs:318-351) leaves them
        const pb200_body_t& H = out->bodies[e->P.host];
        for (size_t b = 0; b < nb; b++)
            for (int c = 0; c < 3; c++) {
                out->bodies[b].heliocentric_position[c] = (int)b == e->P.host ? 0. : out->bodies[b].inertial_position[c] - H.inertial_position[c];
                out->bodies[b].heliocentric_velocity[c] = (int)b == e->P.host ? 0. : out->bodies[b].inertial_velocity[c] - H.inertial_velocity[c];
            }
    }
    const double* tail = g.data() + nb * PB_GATHER_PER_BODY;
    for (size_t i = 0; i < nb * nb; i++) out->roche_radiuses[i] = tail[i];
    const double* sysv = tail + nb * nb;
    out->current_time = sysv[0];
    out->last_historic_snapshot_time = sysv[1];
    uint64_t u;
    std::memcpy(&u, &sysv[2], 8); out->current_iteration = u;
    std::memcpy(&u, &sysv[3], 8); out->n_historic_snapshots = u;
    std::memcpy(&u, &sysv[4], 8); out->timestep_warning = u;
    return PB200_OK;
}

size_t pb200_ensemble_history_pending(pb200_ensemble_t* e) {
    if (!e) return 0;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    // all live systems snapshot in lock step; take the maximum
    std::vector<int> hc(e->n_sys);
    if (cudaMemcpy(hc.data(), e->P.hist_count, e->n_sys * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    int mx = 0;
    for (int v : hc) if (v > mx) mx = v;
    return (size_t)mx;
}

int pb200_ensemble_history_drain(pb200_ensemble_t* e, void* dst, size_t dst_bytes) {
    if (!e) return set_error(PB200_E_INVALID, "null ensemble");
    CUDA_TRY(cudaSetDevice(e->device));
    size_t n_snap = pb200_ensemble_history_pending(e);
    if (n_snap == 0) return PB200_OK;
    const size_t nrec = e->n_sys * n_snap * (size_t)e->n_bodies;
    const size_t bytes = nrec * PB200_HISTORIC_RECORD_BYTES;
    if (!dst || dst_bytes < bytes) return set_error(PB200_E_INVALID, "history destination too small");
    if (e->records_capacity < bytes) {
        int rc = dev_alloc(e, &e->d_records, bytes / 4 + 1);
        if (rc != PB200_OK) return rc;
        e->records_capacity = bytes;
    }
    {
        const size_t smem = 32 * (size_t)((e->n_bodies * PB_REC_WORDS) | 1) * sizeof(unsigned int);   // <= 50 KB (10 bodies)
        static thread_local int configured_device = -1;
        if (configured_device != e->device) {
            CUDA_TRY(cudaFuncSetAttribute(pack_history_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            configured_device = e->device;
        }
        dim3 grid((unsigned)((e->n_sys + 31) / 32), (unsigned)n_snap);
        pack_history_kernel<<<grid, 256, smem, e->stream>>>(e->P, (int)n_snap, e->P.dt, e->d_records);
    }
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(dst, e->d_records, bytes, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaMemsetAsync(e->P.hist_count, 0, e->n_sys * sizeof(int), e->stream));
    e->hist_pending_host = 0;
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return PB200_OK;
}

int pb200_ensemble_summary(pb200_ensemble_t* e, double* energy, double* angular_momentum) {
    if (!e) return set_error(PB200_E_INVALID, "null ensemble");
    CUDA_TRY(cudaSetDevice(e->device));
    summary_kernel<<<(unsigned)((e->n_sys + 127) / 128), 128, 0, e->stream>>>(e->P, e->d_energy, e->d_angmom);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    if (energy) CUDA_TRY(cudaMemcpyAsync(energy, e->d_energy, e->n_sys * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    if (angular_momentum) CUDA_TRY(cudaMemcpyAsync(angular_momentum, e->d_angmom, e->n_sys * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return PB200_OK;
}

int pb200_measure_fp64_peak(int device, double ms_target, double* flops_per_s) {
    if (!flops_per_s) return set_error(PB200_E_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    const int threads = 256, blocks = prop.multiProcessorCount * 8;
    double* out = nullptr;
    CUDA_TRY(cudaMalloc(&out, (size_t)threads * blocks * sizeof(double)));
    cudaEvent_t a, b;
    CUDA_TRY(cudaEventCreate(&a)); CUDA_TRY(cudaEventCreate(&b));
    int iters = 256;
    double best = 0.;
    float ms = 0.f;
    for (int rep = 0; rep < 12; rep++) {
        CUDA_TRY(cudaEventRecord(a));
        dfma_peak_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        CUDA_TRY(cudaEventRecord(b));
        CUDA_TRY(cudaEventSynchronize(b));
        CUDA_TRY(cudaEventElapsedTime(&ms, a, b));
        double flops = 2.0 * 8 * 16 * (double)iters * threads * blocks;
        double rate = flops / (ms * 1e-3);
        if (rep > 0 && rate > best) best = rate;
        if (ms < ms_target && iters < (1 << 24)) iters *= 2;
    }
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(out);
    *flops_per_s = best;
    return PB200_OK;
}

}  // extern "C"
