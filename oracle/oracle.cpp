// oracle.cpp — C entry points of the CPU oracle (see oracle_core.hpp for scope and parity status).
// TEST INFRASTRUCTURE ONLY: loaded by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
#include <chrono>
#include <thread>
#include <vector>
#include "oracle_core.hpp"

namespace pb200_oracle {

// Counting scalar: every + - * / sqrt of the restatement counts as one flop (SURVEY §8d counting rule).
struct Counted {
    double v;
    static thread_local uint64_t n_add, n_mul, n_div, n_sqrt;
    Counted() : v(0.) {}
    Counted(double x) : v(x) {}
    Counted(int x) : v(x) {}
};
thread_local uint64_t Counted::n_add = 0, Counted::n_mul = 0, Counted::n_div = 0, Counted::n_sqrt = 0;
inline Counted operator+(Counted a, Counted b) { Counted::n_add++; return Counted(a.v + b.v); }
inline Counted operator-(Counted a, Counted b) { Counted::n_add++; return Counted(a.v - b.v); }
inline Counted operator*(Counted a, Counted b) { Counted::n_mul++; return Counted(a.v * b.v); }
inline Counted operator/(Counted a, Counted b) { Counted::n_div++; return Counted(a.v / b.v); }
inline Counted operator-(Counted a) { return Counted(-a.v); }
inline Counted o_sqrt(Counted x) { Counted::n_sqrt++; return Counted(std::sqrt(x.v)); }
inline Counted o_abs(Counted x) { return Counted(std::fabs(x.v)); }
inline Counted o_floor(Counted x) { return Counted(std::floor(x.v)); }
inline Counted o_pow(Counted x, double y) { Counted::n_mul++; return Counted(std::pow(x.v, y)); }
inline double o_val(Counted x) { return x.v; }

}  // namespace pb200_oracle

using namespace pb200_oracle;
typedef System<double> Sys;

extern "C" {

void* pb200_oracle_create(const pb200_case_t* c, const pb200_table_t* tables, size_t n_tables) {
    Sys* s = new Sys();
    s->load(*c, tables, n_tables);
    return s;
}
void pb200_oracle_destroy(void* h) { delete (Sys*)h; }
int pb200_oracle_initialize_physical_values(void* h) { return ((Sys*)h)->initialize_physical_values(); }
// Calls Integrator::iterate up to n_steps times; returns the number of steps actually taken.
uint64_t pb200_oracle_iterate(void* h, uint64_t n_steps) {
    Sys* s = (Sys*)h;
    uint64_t done = 0;
    for (uint64_t k = 0; k < n_steps; k++) {
        if (s->status != PB200_STATUS_OK) break;
        uint64_t before = s->current_iteration;
        s->iterate();
        done += s->current_iteration - before;
    }
    return done;
}
int pb200_oracle_status(void* h, uint32_t* warnings, uint64_t* iteration) {
    Sys* s = (Sys*)h;
    if (warnings) *warnings = s->warnings;
    if (iteration) *iteration = s->event_iteration;
    return s->status;
}
void pb200_oracle_store(void* h, pb200_case_t* out) { ((Sys*)h)->store(*out); }
size_t pb200_oracle_history_bytes(void* h) { return ((Sys*)h)->history.size(); }
size_t pb200_oracle_history_drain(void* h, void* dst, size_t cap) {
    Sys* s = (Sys*)h;
    size_t nb = s->history.size() < cap ? s->history.size() : cap;
    std::memcpy(dst, s->history.data(), nb);
    s->history.clear();
    return nb;
}
void pb200_oracle_summary(void* h, double* e, double* l) { ((Sys*)h)->summary(*e, *l); }
int pb200_oracle_last_midpoint_iterations(void* h) { return ((Sys*)h)->last_midpoint_iterations; }
// calls of kepler_individual_step by the branch they took: Newton converged, quartic solver, bisection fallback, hyperbolic
void pb200_oracle_kepler_branches(void* h, uint64_t* out) { for (int k = 0; k < 4; k++) out[k] = ((Sys*)h)->kepler_branches[k]; }
// The Universe::calculate_additional_effects evaluation alone (for unit parity of accelerations/torques):
// out_acc / out_dldt: 3 * n doubles, body-major [b][c].
void pb200_oracle_additional_effects(void* h, double* out_acc, double* out_dldt) {
    Sys* s = (Sys*)h;
    s->inertial_to_heliocentric();
    bool gr_spin = s->c_gr && s->gr_impl == PB200_GR_KIDDER1995;
    bool integrate_spin = s->c_tides || s->c_flat || s->c_evo || gr_spin;
    s->calculate_additional_effects(s->current_time, true, integrate_spin, true, s->ignore_terms());
    for (int i = 0; i < s->n; i++) {
        out_acc[3 * i + 0] = s->p[i].iadd.x; out_acc[3 * i + 1] = s->p[i].iadd.y; out_acc[3 * i + 2] = s->p[i].iadd.z;
        out_dldt[3 * i + 0] = s->p[i].dLdt.x; out_dldt[3 * i + 1] = s->p[i].dLdt.y; out_dldt[3 * i + 2] = s->p[i].dLdt.z;
    }
}

// Runs n independent systems for n_steps each on n_threads host threads (the CPU baseline:
// one process-equivalent per system, like `posidonius start` per member). cases[i] -> out[i].
// n_cases is 1 (replicated) or n. Returns the wall-clock seconds of the stepping phase.
int pb200_oracle_run_ensemble(const pb200_case_t* cases, size_t n_cases, size_t n, const pb200_table_t* tables,
                              size_t n_tables, uint64_t n_steps, int init_physical, int n_threads,
                              pb200_case_t* out, int32_t* status, double* seconds) {
    if (n_threads < 1) n_threads = 1;
    std::vector<Sys*> sys(n);
    for (size_t i = 0; i < n; i++) {
        sys[i] = new Sys();
        sys[i]->load(cases[n_cases == 1 ? 0 : i], tables, n_tables);
        if (init_physical) sys[i]->initialize_physical_values();
    }
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) {
        th.emplace_back([&, t]() {
            for (size_t i = (size_t)t; i < n; i += (size_t)n_threads) {
                Sys* s = sys[i];
                for (uint64_t k = 0; k < n_steps; k++) {
                    if (!s->iterate()) break;
                }
                s->history.clear();
            }
        });
    }
    for (auto& x : th) x.join();
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    for (size_t i = 0; i < n; i++) {
        if (out) { out[i] = cases[n_cases == 1 ? 0 : i]; sys[i]->store(out[i]); }
        if (status) status[i] = sys[i]->status;
        delete sys[i];
    }
    return 0;
}

// Exact operation counts of n_steps steps of one system under the counting rule of SURVEY §8(d):
// counts = {add/sub, mul, div, sqrt, midpoint evaluations, stumpff evaluations}.
int pb200_oracle_count_flops(const pb200_case_t* c, const pb200_table_t* tables, size_t n_tables, uint64_t n_steps,
                             int init_physical, uint64_t* counts) {
    System<Counted>* s = new System<Counted>();
    s->load(*c, tables, n_tables);
    if (init_physical) s->initialize_physical_values();
    // one untimed step so that the first-snapshot refresh is not counted
    s->iterate();
    Counted::n_add = Counted::n_mul = Counted::n_div = Counted::n_sqrt = 0;
    s->kepler_stumpff_calls = 0;
    uint64_t evals = 0;
    for (uint64_t k = 0; k < n_steps; k++) {
        if (!s->iterate()) break;
        evals += 0;
    }
    counts[0] = Counted::n_add; counts[1] = Counted::n_mul; counts[2] = Counted::n_div; counts[3] = Counted::n_sqrt;
    counts[4] = evals; counts[5] = s->kepler_stumpff_calls;
    delete s;
    return 0;
}

const char* pb200_oracle_version(void) { return "posidonius_b200 oracle 0.1 (CPU restatement, test infrastructure)"; }

}  // extern "C"
