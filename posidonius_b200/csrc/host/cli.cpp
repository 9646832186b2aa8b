// cli.cpp — `posidonius-b200 start | resume | ensemble`: the reference's command line (src/main.rs:10-184) over the
// C ABI of libposidonius_b200.so, plus the new ensemble subcommand.
//
//   start  <case.json> <recovery.bin> <history.bin> [-l|--limit seconds] [-s|--silent]
//   resume <recovery.bin> <history.bin> [-l seconds] [-s] [--historic-snapshot-period days]
//          [--recovery-snapshot-period days] [--time-limit days]
//   ensemble <case.json> <out_dir> --systems N [--seed S] [--amplitude A] [--steps K] [--device D] [--arithmetic hybrid|strict|fast]
//
// Same semantics as the reference: `start` refuses to overwrite existing outputs (main.rs:144-148); the history file is
// truncated to what the recovery snapshot knows (output.rs:91-117); a recovery snapshot is written when iterate() asks
// for it, or — when --limit is given — only once the wall-clock limit is hit (main.rs:158-176); completion prints
// "Simulation completed"; physical failures end the process like the reference's panic! (exit code 101).
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>
#include <sys/stat.h>
#include <unistd.h>
#include "../../../include/posidonius_b200.h"

static std::string stamp() {
    time_t now = time(nullptr);
    struct tm g;
    gmtime_r(&now, &g);
    char b[64];
    snprintf(b, sizeof b, "%04d.%02d.%02d %02d:%02d:%02d UTC", g.tm_year + 1900, g.tm_mon + 1, g.tm_mday, g.tm_hour, g.tm_min, g.tm_sec);
    return b;
}
#define INFO(...) do { printf("[INFO %s] ", stamp().c_str()); printf(__VA_ARGS__); printf("\n"); fflush(stdout); } while (0)
#define WARN(...) do { printf("[WARNING %s] ", stamp().c_str()); printf(__VA_ARGS__); printf("\n"); fflush(stdout); } while (0)
[[noreturn]] static void panic(const std::string& msg) {
    printf("[PANIC %s] %s\n", stamp().c_str(), msg.c_str());
    fflush(stdout);
    exit(101);   // the exit code of a Rust panic
}
static void check(int rc, const char* what) {
    if (rc != PB200_OK) panic(std::string(what) + ": " + pb200_last_error());
}
static bool exists(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0; }

struct Args {
    std::vector<std::string> pos;
    double limit_s = 0., hist_period = -1., rec_period = -1., time_limit = -1., amplitude = 1e-3;
    bool silent = false, strict = false;
    int arithmetic = -1;   // -1: the library default (hybrid)
    long long systems = 0, seed = 20261017, steps = -1;
    int device = 0;
};

static Args parse(int argc, char** argv, int first) {
    Args a;
    for (int i = first; i < argc; i++) {
        std::string s = argv[i];
        auto val = [&](const char* name) -> std::string {
            if (i + 1 >= argc) panic(std::string("option ") + name + " needs a value");
            return argv[++i];
        };
        if (s == "-l" || s == "--limit") a.limit_s = atof(val("--limit").c_str());
        else if (s == "-s" || s == "--silent") a.silent = true;
        else if (s == "--historic-snapshot-period") a.hist_period = atof(val(s.c_str()).c_str());
        else if (s == "--recovery-snapshot-period") a.rec_period = atof(val(s.c_str()).c_str());
        else if (s == "--time-limit") a.time_limit = atof(val(s.c_str()).c_str());
        else if (s == "--systems") a.systems = atoll(val(s.c_str()).c_str());
        else if (s == "--seed") a.seed = atoll(val(s.c_str()).c_str());
        else if (s == "--amplitude") a.amplitude = atof(val(s.c_str()).c_str());
        else if (s == "--steps") a.steps = atoll(val(s.c_str()).c_str());
        else if (s == "--device") a.device = atoi(val(s.c_str()).c_str());
        else if (s == "--strict") a.strict = true;
        else if (s == "--arithmetic") {
            std::string v = val(s.c_str());
            if (v == "fast") a.arithmetic = PB200_ARITH_FAST; else if (v == "strict") a.arithmetic = PB200_ARITH_STRICT;
            else if (v == "hybrid") a.arithmetic = PB200_ARITH_HYBRID; else panic("--arithmetic takes fast, strict or hybrid");
        }
        else if (!s.empty() && s[0] == '-') panic("unknown option " + s);
        else a.pos.push_back(s);
    }
    return a;
}

static const char* failure_text(int status) {
    switch (status) {
        case PB200_STATUS_ROCHE_DESTROYED: return "A particle was destroyed by another one due to a close encounter!";
        case PB200_STATUS_COLLISION: return "Collision between two particles!";
        case PB200_STATUS_EJECTED: return "A particle has been ejected!";
        case PB200_STATUS_ZERO_INERTIA: return "Moment of inertia of a particle is zero!";
        default: return "unknown failure";
    }
}

// Steps until (and including) the iterate() call that returns Ok(true): first snapshot ever, or
// last_recovery + period <= t at the start of the step (whfast.rs:237-239, 303). `triggered` = found within `cap` steps.
static uint64_t steps_to_recovery_trigger(const pb200_case_t& c, uint64_t cap, bool& triggered) {
    triggered = true;
    if (c.last_historic_snapshot_time < 0.) return 1;
    double t = c.current_time;
    for (uint64_t k = 0; k < cap; k++) {
        if (c.last_recovery_snapshot_time + c.recovery_snapshot_period <= t) return k + 1;
        t += c.time_step;
    }
    triggered = false;
    return cap;
}

static int run_single(const Args& a, bool resume) {
    const std::string first = a.pos[0];
    const std::string recovery = resume ? a.pos[0] : a.pos[1];
    const std::string history = resume ? a.pos[1] : a.pos[2];
    auto t_start = std::chrono::steady_clock::now();
    pb200_case_t c;
    pb200_table_store_t* store = nullptr;
    if (pb200_case_load(first.c_str(), &c, &store) != PB200_OK)
        panic(std::string(resume ? "It was not possible to resume the simulation: " : "It was not possible to start the simulation: ") + pb200_last_error());
    if (c.current_time == 0.) INFO("Created new simulation based on '%s'.", first.c_str());
    else { INFO("Restored previous simulation from '%s'.", first.c_str()); INFO("Continuing from year %.0f (%.1e).", c.current_time / 365.25, c.current_time / 365.25); }
    pb200_ensemble_t* e = nullptr;
    check(pb200_ensemble_create(&c, 1, 1, pb200_table_store_tables(store), pb200_table_store_count(store), a.device, &e), "cannot create the GPU integrator");
    if (a.strict) check(pb200_ensemble_set_arithmetic(e, PB200_ARITH_STRICT), "strict arithmetic");
    else if (a.arithmetic >= 0) check(pb200_ensemble_set_arithmetic(e, a.arithmetic), "arithmetic mode");
    if (c.current_time == 0.) check(pb200_ensemble_initialize_physical_values(e), "initialize_physical_values");
    // set_snapshot_periods / set_time_limit (whfast.rs:187-224)
    if (a.hist_period > 0. && a.hist_period != c.historic_snapshot_period) INFO("The historic snapshot period changed from %g to %g days", c.historic_snapshot_period, a.hist_period);
    else INFO("A historic snapshot will be saved every %g days", c.historic_snapshot_period);
    if (a.rec_period > 0. && a.rec_period != c.recovery_snapshot_period) INFO("The recovery snapshot period changed from %g to %g days", c.recovery_snapshot_period, a.rec_period);
    else INFO("A recovery snapshot will be saved every %g days", c.recovery_snapshot_period);
    check(pb200_ensemble_set_snapshot_periods(e, a.hist_period, a.rec_period), "set_snapshot_periods");
    if (a.time_limit > 0. && a.time_limit != c.time_limit) {
        if (a.time_limit < c.current_time) panic("Your new time limit is smaller than the current time");
        INFO("The time limit changed from %g to %g days", c.time_limit, a.time_limit);
    }
    check(pb200_ensemble_set_time_limit(e, a.time_limit), "set_time_limit");
    if (!resume && exists(recovery)) panic("File '" + recovery + "' already exists.");
    if (!resume && exists(history)) panic("File '" + history + "' already exists.");
    // get_universe_history_writer (output.rs:91-117)
    const int n = c.n_particles;
    const uint64_t expected = c.n_historic_snapshots * (uint64_t)PB200_HISTORIC_RECORD_BYTES * (uint64_t)n;
    FILE* hf = fopen(history.c_str(), "ab");
    if (!hf) panic("File error: cannot open " + history);
    {
        struct stat st;
        stat(history.c_str(), &st);
        if ((uint64_t)st.st_size < expected) panic("Historic snapshots do not contain all the expected history as indicated by the recovery snapshot");
        if (truncate(history.c_str(), (off_t)expected) != 0) panic("cannot truncate " + history);
    }
    std::vector<unsigned char> records;
    const bool limited = a.limit_s > 0.;
    bool completed = false;
    pb200_case_t img;
    check(pb200_ensemble_get_case(e, 0, &img), "get_case");
    // WHFast.last_recovery_snapshot_time lives on the host only (write_recovery_snapshot sets it, whfast.rs:307-308): the image
    // that get_case returns carries the value of the loaded file, so it is tracked here and re-applied after every get_case
    double last_recovery = img.last_recovery_snapshot_time;
    const bool trace = getenv("PB200_CLI_TRACE") != nullptr;
    const size_t capacity = pb200_ensemble_history_capacity(e);
    while (!completed) {
        // never more historic snapshots per launch than the device history buffer holds, never past the next recovery trigger
        double per = img.historic_snapshot_period / img.time_step;
        uint64_t cap = (uint64_t)std::max(1.0, std::min(20000.0, per >= 1. ? (double)(capacity - 1) * per : (double)capacity));
        bool trigger = false;
        uint64_t k = cap;
        if (!limited) k = steps_to_recovery_trigger(img, cap, trigger);
        check(pb200_ensemble_step(e, k), "step");
        size_t pending = pb200_ensemble_history_pending(e);
        if (pending) {
            records.resize(pending * (size_t)n * PB200_HISTORIC_RECORD_BYTES);
            check(pb200_ensemble_history_drain(e, records.data(), records.size()), "history_drain");
            if (fwrite(records.data(), 1, records.size(), hf) != records.size()) panic("write failed on " + history);
        }
        int32_t status = 0;
        check(pb200_ensemble_status(e, &status, nullptr, nullptr), "status");
        check(pb200_ensemble_get_case(e, 0, &img), "get_case");
        img.last_recovery_snapshot_time = last_recovery;
        if (!a.silent) { printf("Year: %.0f (%.1e) | Time step: %.3f days                    \r", img.current_time / 365.25, img.current_time / 365.25, img.time_step); fflush(stdout); }
        if (status == PB200_STATUS_COMPLETED) { INFO("Simulation completed '%s'.", first.c_str()); completed = true; break; }
        if (status != PB200_STATUS_OK) { fflush(hf); printf("\n\n"); panic(failure_text(status)); }
        bool write_recovery = trigger;
        if (limited) {
            double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
            if (el >= a.limit_s) write_recovery = true;
        }
        if (write_recovery) {
            // Integrator::write_recovery_snapshot (whfast.rs:307-316)
            fflush(hf);
            last_recovery = img.last_recovery_snapshot_time = img.current_time;
            if (trace) { printf("[TRACE] recovery snapshot at t = %.17g (iteration %llu)\n", img.current_time, (unsigned long long)img.current_iteration); fflush(stdout); }
            check(pb200_case_save(recovery.c_str(), &img, pb200_table_store_tables(store), pb200_table_store_count(store)), "write_recovery_snapshot");
            if (limited) { WARN("Reached execution time limit before simulation completion"); break; }
        }
    }
    fclose(hf);
    double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    if (!resume) INFO("Execution time: %g seconds", el);
    else INFO("Execution time since last resume: %g seconds", el);
    pb200_ensemble_destroy(e);
    pb200_table_store_free(store);
    return 0;
}

// ---- ensemble: N perturbed members of one case, built on the device from a deterministic SplitMix64
// stream (pb200_ensemble_create_perturbed): no per-member images on the host
static int run_ensemble(const Args& a) {
    if (a.pos.size() < 2 || a.systems <= 0) panic("usage: posidonius-b200 ensemble <case.json> <out_dir> --systems N [--seed S] [--amplitude A] [--steps K] [--device D] [--arithmetic hybrid|strict|fast]");
    const std::string out_dir = a.pos[1];
    pb200_case_t base;
    pb200_table_store_t* store = nullptr;
    check(pb200_case_load(a.pos[0].c_str(), &base, &store), "cannot read the case");
    mkdir(out_dir.c_str(), 0777);
    const size_t S = (size_t)a.systems;
    pb200_ensemble_t* e = nullptr;
    check(pb200_ensemble_create_perturbed(&base, S, (uint64_t)a.seed, a.amplitude, pb200_table_store_tables(store), pb200_table_store_count(store), a.device, &e),
          "cannot create the ensemble");
    if (a.strict) check(pb200_ensemble_set_arithmetic(e, PB200_ARITH_STRICT), "strict arithmetic");
    else if (a.arithmetic >= 0) check(pb200_ensemble_set_arithmetic(e, a.arithmetic), "arithmetic mode");
    if (base.current_time == 0.) check(pb200_ensemble_initialize_physical_values(e), "initialize_physical_values");
    const int n = base.n_particles;
    std::vector<double> e0(S), l0(S), e1(S), l1(S);
    check(pb200_ensemble_summary(e, e0.data(), l0.data()), "summary");
    uint64_t total = a.steps > 0 ? (uint64_t)a.steps : (uint64_t)std::ceil((base.time_limit - base.current_time) / base.time_step);
    FILE* hf = fopen((out_dir + "/ensemble_history.bin").c_str(), "wb");
    if (!hf) panic("cannot create " + out_dir + "/ensemble_history.bin");
    std::vector<unsigned char> records;
    uint64_t done = 0, snapshots = 0;
    auto t0 = std::chrono::steady_clock::now();
    double kernel_ms = 0.;
    const size_t capacity = pb200_ensemble_history_capacity(e);
    while (done < total) {
        // stay inside the device history buffer (drained after every call): k steps produce at most ceil(k / per) + 1 snapshots
        double per = base.historic_snapshot_period / base.time_step;
        uint64_t k = std::min<uint64_t>(total - done, (uint64_t)std::max(1.0, std::min(per >= 1. ? (double)(capacity - 1) * per : (double)capacity, 100000.0)));
        check(pb200_ensemble_step(e, k), "step");
        float ms = 0.f;
        check(pb200_ensemble_last_step_ms(e, &ms), "timing");
        kernel_ms += ms;
        done += k;
        size_t pending = pb200_ensemble_history_pending(e);
        if (pending) {
            // block layout on disk: [call][system][snapshot][body] 156-byte records; index.csv tells the reader the block sizes
            records.resize(S * pending * (size_t)n * PB200_HISTORIC_RECORD_BYTES);
            check(pb200_ensemble_history_drain(e, records.data(), records.size()), "history_drain");
            fwrite(records.data(), 1, records.size(), hf);
            snapshots += pending;
        }
        if (!a.silent) { printf("steps %llu / %llu\r", (unsigned long long)done, (unsigned long long)total); fflush(stdout); }
    }
    fclose(hf);
    double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::vector<int32_t> status(S);
    std::vector<uint32_t> warn(S);
    std::vector<uint64_t> iter(S);
    check(pb200_ensemble_status(e, status.data(), warn.data(), iter.data()), "status");
    check(pb200_ensemble_summary(e, e1.data(), l1.data()), "summary");
    std::vector<double> t(S);
    pb200_state_view_t view;
    std::memset(&view, 0, sizeof view);
    view.current_time = t.data();
    check(pb200_ensemble_download(e, &view), "download");
    FILE* sf = fopen((out_dir + "/summary.csv").c_str(), "w");
    fprintf(sf, "system,status,warnings,iteration_of_event,current_time,energy_initial,energy_final,angular_momentum_initial,angular_momentum_final\n");
    size_t ok = 0, comp = 0;
    for (size_t k = 0; k < S; k++) {
        fprintf(sf, "%zu,%d,%u,%llu,%.17g,%.17g,%.17g,%.17g,%.17g\n", k, status[k], warn[k], (unsigned long long)iter[k], t[k], e0[k], e1[k], l0[k], l1[k]);
        ok += status[k] == PB200_STATUS_OK;
        comp += status[k] == PB200_STATUS_COMPLETED;
    }
    fclose(sf);
    // the final image of every member as a recovery snapshot the reference can resume (only for small ensembles)
    if (S <= 4096) {
        for (size_t k = 0; k < S; k++) {
            pb200_case_t img;
            check(pb200_ensemble_get_case(e, k, &img), "get_case");
            img.last_recovery_snapshot_time = img.current_time;
            char name[64];
            snprintf(name, sizeof name, "/recovery_%06zu.bin", k);
            check(pb200_case_save((out_dir + name).c_str(), &img, pb200_table_store_tables(store), pb200_table_store_count(store)), "recovery");
        }
    }
    INFO("ensemble of %zu systems x %llu steps: %zu running, %zu completed, %zu stopped by a physical failure; %llu snapshot(s) per system",
         S, (unsigned long long)done, ok, comp, S - ok - comp, (unsigned long long)snapshots);
    INFO("kernel time %.3f s (%.3e system-steps/s), wall %.3f s", kernel_ms * 1e-3, (double)S * (double)done / (kernel_ms * 1e-3), wall);
    pb200_ensemble_destroy(e);
    pb200_table_store_free(store);
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 2) {
        fprintf(stderr, "%s\nusage: posidonius-b200 start <case.json> <recovery.bin> <history.bin> [-l seconds] [-s]\n"
                        "       posidonius-b200 resume <recovery.bin> <history.bin> [-l seconds] [-s] [--historic-snapshot-period d] [--recovery-snapshot-period d] [--time-limit d]\n"
                        "       posidonius-b200 ensemble <case.json> <out_dir> --systems N [--seed S] [--amplitude A] [--steps K] [--device D] [--arithmetic hybrid|strict|fast]\n",
                pb200_version());
        return 2;
    }
    std::string cmd = argv[1];
    Args a = parse(argc, argv, 2);
    if (cmd == "start") { if (a.pos.size() != 3) panic("start needs <case.json> <recovery.bin> <history.bin>"); return run_single(a, false); }
    if (cmd == "resume") { if (a.pos.size() != 2) panic("resume needs <recovery.bin> <history.bin>"); return run_single(a, true); }
    if (cmd == "ensemble") return run_ensemble(a);
    panic("unknown subcommand " + cmd);
}
