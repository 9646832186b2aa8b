// case_io.cpp — reading and writing the reference's integrator images (host side, no GPU).
//
//   pb200_case_load  <-  output::restore_snapshot / deserialize_json_snapshot / deserialize_bin_snapshot (output.rs:206-307)
//   pb200_case_save  <-  output::write_recovery_snapshot (output.rs:56-82)
//
// The image is the serde form of `WHFast` (integrator/whfast.rs:98-120) with `Universe` (particles/universe.rs:50-63),
// `Particle` (particles/particle.rs:16-52) and the effect structs, fields in declaration order. bincode 1.3.3 default
// options: little endian, fixed-width integers, usize -> u64, enum variant index u32, Vec/HashMap with a u64 length prefix,
// fixed arrays (serde_big_array) without prefix, bool one byte. One schema walk (`walk_image`) drives the bincode reader,
// the bincode writer and the JSON writer; the JSON reader is key-based because the Python case generator writes sorted keys.
// Scratch fields that the integrator recomputes before every use are written as zeros.
#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
#include <sys/stat.h>
#include "../../../include/posidonius_b200.h"
#include "json_min.hpp"

int pb200_set_error_message(int code, const std::string& msg);   // pb200_api.cu

struct pb200_table_store {
    std::vector<std::vector<double>> columns;   // 5 per table: time, radius, rg2, love, qinv
    std::vector<pb200_table_t> tables;
    void rebuild() {
        tables.clear();
        for (size_t t = 0; t * 5 < columns.size(); t++) {
            pb200_table_t T;
            const std::vector<double>* c = &columns[t * 5];
            T.n_rows = c[0].size();
            T.time = c[0].empty() ? nullptr : c[0].data();
            T.radius = c[1].empty() ? nullptr : c[1].data();
            T.radius_of_gyration_2 = c[2].empty() ? nullptr : c[2].data();
            T.love_number = c[3].empty() ? nullptr : c[3].data();
            T.inverse_tidal_q_factor = c[4].empty() ? nullptr : c[4].data();
            tables.push_back(T);
        }
    }
};

namespace {

struct Unsupported : std::runtime_error { using std::runtime_error::runtime_error; };

const char* const kCoordNames[] = {"Jacobi", "DemocraticHeliocentric", "WHDS"};
const char* const kRoleNames[] = {"CentralBody", "OrbitingBody", "Disabled"};
const char* const kTidalModelNames[] = {"ConstantTimeLag", "CreepCoplanar", "Kaula"};
const char* const kFlatModelNames[] = {"OblateSpheroid", "CreepCoplanar"};
const char* const kGrNames[] = {"Kidder1995", "Anderson1975", "Newhall1983", "Disabled"};
const char* const kWindNames[] = {"Interaction", "Disabled"};
const char* const kRefNames[] = {"MostMassiveParticle", "Particle"};
const char* const kEvoNames[] = {"GalletBolmont2017", "BolmontMathis2016", "Baraffe2015", "Leconte2011", "Baraffe1998", "LeconteChabrier2013", "NonEvolving"};

// Everything of the image that pb200_case_t does not carry and that is not recomputed: nothing. The walk reads/writes
// through these references; scratch goes through `zero`.
struct Image {
    pb200_case_t c;
    std::vector<std::vector<double>> evo[PB200_MAX_PARTICLES];  // per particle slot: 5 columns
    uint64_t evo_left_index[PB200_MAX_PARTICLES];
    uint64_t hash = 0;
    Image() {
        std::memset(&c, 0, sizeof c);
        for (int i = 0; i < PB200_MAX_PARTICLES; i++) { evo[i].assign(5, {}); evo_left_index[i] = 0; }
    }
};

// ---------------------------------------------------------------- visitors
struct BinWriter {
    std::string out;
    static const bool reading = false;
    void f64(const char*, double& v) { out.append((const char*)&v, 8); }
    void u64(const char*, uint64_t& v) { out.append((const char*)&v, 8); }
    void boolean(const char*, bool& v) { out.push_back(v ? 1 : 0); }
    void begin_struct(const char*) {}
    void end_struct() {}
    void begin_array(const char*) {}
    void end_array() {}
    void begin_vec(const char*, uint64_t& n) { u64(nullptr, n); }
    void end_vec() {}
    void begin_map(const char*, uint64_t& n) { u64(nullptr, n); }
    void map_entry(uint64_t& key, double& value) { u64(nullptr, key); f64(nullptr, value); }
    void end_map() {}
    int enum_begin(const char*, int tag, const char* const*, int, bool) { uint32_t t = (uint32_t)tag; out.append((const char*)&t, 4); return tag; }
    void enum_end(bool) {}
};

struct BinReader {
    const std::string& in;
    size_t p = 0;
    static const bool reading = true;
    explicit BinReader(const std::string& s) : in(s) {}
    void need(size_t n) { if (p + n > in.size()) throw std::runtime_error("bincode image truncated"); }
    void f64(const char*, double& v) { need(8); std::memcpy(&v, in.data() + p, 8); p += 8; }
    void u64(const char*, uint64_t& v) { need(8); std::memcpy(&v, in.data() + p, 8); p += 8; }
    void boolean(const char*, bool& v) { need(1); v = in[p] != 0; p += 1; }
    void begin_struct(const char*) {}
    void end_struct() {}
    void begin_array(const char*) {}
    void end_array() {}
    void begin_vec(const char*, uint64_t& n) { u64(nullptr, n); if (n > (1ull << 28)) throw std::runtime_error("bincode image: implausible length"); }
    void end_vec() {}
    void begin_map(const char*, uint64_t& n) { u64(nullptr, n); if (n > (1ull << 20)) throw std::runtime_error("bincode image: implausible map length"); }
    void map_entry(uint64_t& key, double& value) { u64(nullptr, key); f64(nullptr, value); }
    void end_map() {}
    int enum_begin(const char* key, int, const char* const*, int nvariants, bool) {
        need(4);
        uint32_t t;
        std::memcpy(&t, in.data() + p, 4);
        p += 4;
        if ((int)t >= nvariants) throw std::runtime_error(std::string("bincode image: bad enum tag for ") + (key ? key : "?"));
        return (int)t;
    }
    void enum_end(bool) {}
};

struct JsonWriter {
    pbjson::Writer w;
    static const bool reading = false;
    void k(const char* key) { if (key) w.key(key); }
    void f64(const char* key, double& v) { k(key); w.number(v); }
    void u64(const char* key, uint64_t& v) { k(key); w.integer(v); }
    void boolean(const char* key, bool& v) { k(key); w.boolean(v); }
    void begin_struct(const char* key) { k(key); w.begin_object(); }
    void end_struct() { w.end_object(); }
    void begin_array(const char* key) { k(key); w.begin_array(); }
    void end_array() { w.end_array(); }
    void begin_vec(const char* key, uint64_t&) { k(key); w.begin_array(); }
    void end_vec() { w.end_array(); }
    void begin_map(const char* key, uint64_t&) { k(key); w.begin_object(); }
    void map_entry(uint64_t& key, double& value) { w.key(std::to_string(key).c_str()); w.number(value); }   // serde_json: integer keys as strings
    void end_map() { w.end_object(); }
    int enum_begin(const char* key, int tag, const char* const* names, int, bool payload) {
        k(key);
        if (payload) { w.begin_object(); w.key(names[tag]); } else w.string(names[tag]);
        return tag;
    }
    void enum_end(bool payload) { if (payload) w.end_object(); }
};

// ---------------------------------------------------------------- schema walk
template <class V> void axes(V& v, const char* key, double* a) {
    v.begin_struct(key);
    v.f64("x", a[0]); v.f64("y", a[1]); v.f64("z", a[2]);
    v.end_struct();
}
template <class V> void zero_f64(V& v, const char* key) { double z = 0.; v.f64(key, z); }
template <class V> void zero_axes(V& v, const char* key) { double z[3] = {0., 0., 0.}; axes(v, key, z); }
template <class V> void zero_coordinates(V& v) {
    v.begin_struct("coordinates"); zero_axes(v, "position"); zero_axes(v, "velocity"); v.end_struct();
}
template <class V> void output_acc_dl(V& v) {
    v.begin_struct("output"); zero_axes(v, "acceleration"); zero_axes(v, "dangular_momentum_dt"); v.end_struct();
}

template <class V> void walk_particle(V& v, Image& img, int i) {
    pb200_body_t& b = img.c.bodies[i];
    const bool live = i < img.c.n_particles || V::reading;
    v.begin_struct(nullptr);
    uint64_t id = live ? (uint64_t)b.id : 0; v.u64("id", id); b.id = (int32_t)id;
    v.f64("mass", b.mass); v.f64("mass_g", b.mass_g); v.f64("radius", b.radius);
    axes(v, "inertial_position", b.inertial_position); axes(v, "inertial_velocity", b.inertial_velocity);
    axes(v, "inertial_acceleration", b.inertial_acceleration); zero_axes(v, "inertial_additional_acceleration");
    axes(v, "heliocentric_position", b.heliocentric_position); axes(v, "heliocentric_velocity", b.heliocentric_velocity);
    zero_f64(v, "heliocentric_distance"); zero_f64(v, "heliocentric_radial_velocity");
    zero_f64(v, "heliocentric_norm_velocity_vector"); zero_f64(v, "heliocentric_norm_velocity_vector_2");
    axes(v, "spin", b.spin);
    double ns2 = b.spin[0] * b.spin[0] + b.spin[1] * b.spin[1] + b.spin[2] * b.spin[2];
    v.f64("norm_spin_vector_2", ns2);
    axes(v, "angular_momentum", b.angular_momentum); zero_axes(v, "dangular_momentum_dt");
    v.f64("radius_of_gyration_2", b.radius_of_gyration_2); v.f64("moment_of_inertia", b.moment_of_inertia);
    {   // Reference (particle.rs:9-13)
        int tag = b.reference < 0 || !live ? 0 : 1;
        tag = v.enum_begin("reference", tag, kRefNames, 2, tag == 1);
        if (tag == 1) { uint64_t k = (uint64_t)(b.reference < 0 ? 0 : b.reference); v.u64(nullptr, k); b.reference = (int32_t)k; } else b.reference = -1;
        v.enum_end(tag == 1);
    }
    {   // Tides (tides/common.rs:95-100)
        v.begin_struct("tides");
        int role = live ? b.tides_role : PB200_ROLE_DISABLED;
        role = v.enum_begin("effect", role, kRoleNames, 3, role != PB200_ROLE_DISABLED);
        if (role != PB200_ROLE_DISABLED) {
            int model = v.enum_begin(nullptr, 0, kTidalModelNames, 3, true);
            if (model != 0) throw Unsupported(std::string("tidal model ") + kTidalModelNames[model] + " is outside the B200 hot path (ConstantTimeLag only)");
            v.begin_struct(nullptr);
            v.f64("dissipation_factor", b.tides_dissipation_factor); v.f64("dissipation_factor_scale", b.tides_dissipation_factor_scale);
            v.f64("love_number", b.tides_love_number);
            v.end_struct();
            v.enum_end(true);
        }
        v.enum_end(role != PB200_ROLE_DISABLED);
        b.tides_role = role;
        v.begin_struct("parameters");
        v.begin_struct("internal");
        zero_f64(v, "distance"); zero_f64(v, "radial_velocity");
        v.f64("scaled_dissipation_factor", b.tides_scaled_dissipation_factor);
        zero_f64(v, "scalar_product_of_vector_position_with_stellar_spin"); zero_f64(v, "scalar_product_of_vector_position_with_planetary_spin");
        zero_f64(v, "orthogonal_component_of_the_tidal_force_due_to_stellar_tide"); zero_f64(v, "orthogonal_component_of_the_tidal_force_due_to_planetary_tide");
        zero_f64(v, "radial_component_of_the_tidal_force"); zero_f64(v, "radial_component_of_the_tidal_force_dissipative_part_when_star_as_point_mass");
        zero_axes(v, "shape");
        v.f64("denergy_dt", b.tides_denergy_dt); v.f64("lag_angle", b.tides_lag_angle);
        v.end_struct();
        output_acc_dl(v);
        v.end_struct();
        zero_coordinates(v);
        v.end_struct();
    }
    {   // RotationalFlattening (rotational_flattening/common.rs:57-62)
        v.begin_struct("rotational_flattening");
        int role = live ? b.flattening_role : PB200_ROLE_DISABLED;
        role = v.enum_begin("effect", role, kRoleNames, 3, role != PB200_ROLE_DISABLED);
        if (role != PB200_ROLE_DISABLED) {
            int model = v.enum_begin(nullptr, 0, kFlatModelNames, 2, true);
            if (model != 0) throw Unsupported("rotational flattening model CreepCoplanar is outside the B200 hot path (OblateSpheroid only)");
            v.begin_struct(nullptr);
            v.f64("love_number", b.flattening_love_number);
            v.end_struct();
            v.enum_end(true);
        }
        v.enum_end(role != PB200_ROLE_DISABLED);
        b.flattening_role = role;
        v.begin_struct("parameters");
        v.begin_struct("internal");
        zero_f64(v, "distance");
        zero_f64(v, "scalar_product_of_vector_position_with_stellar_spin"); zero_f64(v, "scalar_product_of_vector_position_with_planetary_spin");
        zero_f64(v, "radial_component_of_the_force_induced_by_rotation");
        zero_f64(v, "factor_for_the_force_induced_by_star_rotation"); zero_f64(v, "factor_for_the_force_induced_by_planet_rotation");
        zero_f64(v, "orthogonal_component_of_the_force_induced_by_star_rotation"); zero_f64(v, "orthogonal_component_of_the_force_induced_by_planet_rotation");
        zero_axes(v, "shape");
        v.end_struct();
        output_acc_dl(v);
        v.end_struct();
        zero_coordinates(v);
        v.end_struct();
    }
    {   // GeneralRelativity (general_relativity.rs:54-59)
        v.begin_struct("general_relativity");
        int role = live ? b.general_relativity_role : PB200_ROLE_DISABLED;
        role = v.enum_begin("effect", role, kRoleNames, 3, role == PB200_ROLE_CENTRAL);
        if (role == PB200_ROLE_CENTRAL) {
            int impl = img.c.general_relativity_implementation;
            impl = v.enum_begin(nullptr, impl, kGrNames, 4, false);
            v.enum_end(false);
            if (V::reading) img.c.general_relativity_implementation = impl;   // overwritten by the universe field below
        }
        v.enum_end(role == PB200_ROLE_CENTRAL);
        b.general_relativity_role = role;
        v.begin_struct("parameters");
        v.begin_struct("internal");
        zero_f64(v, "distance"); zero_f64(v, "radial_velocity"); zero_f64(v, "norm_velocity_vector"); zero_f64(v, "norm_velocity_vector_2");
        v.f64("factor", b.general_relativity_factor);
        v.end_struct();
        output_acc_dl(v);
        v.end_struct();
        zero_coordinates(v);
        v.end_struct();
    }
    {   // Wind (wind.rs:35-39)
        v.begin_struct("wind");
        int role = live ? b.wind_role : 1;
        role = v.enum_begin("effect", role, kWindNames, 2, false);
        v.enum_end(false);
        b.wind_role = role;
        v.begin_struct("parameters");
        v.begin_struct("input"); v.f64("k_factor", b.wind_k_factor); v.f64("rotation_saturation", b.wind_rotation_saturation); v.end_struct();
        double rs2 = b.wind_rotation_saturation * b.wind_rotation_saturation;
        v.begin_struct("internal"); v.f64("rotation_saturation_2", rs2); v.end_struct();
        v.begin_struct("output"); zero_axes(v, "dangular_momentum_dt"); v.end_struct();
        v.end_struct();
        v.end_struct();
    }
    {   // Disk (disk.rs:51-56)
        v.begin_struct("disk");
        int role = live ? b.disk_role : PB200_ROLE_DISABLED;
        role = v.enum_begin("effect", role, kRoleNames, 3, role == PB200_ROLE_CENTRAL);
        if (role == PB200_ROLE_CENTRAL) {
            v.begin_struct(nullptr);
            v.f64("inner_edge_distance", b.disk_properties[0]); v.f64("outer_edge_distance", b.disk_properties[1]);
            v.f64("lifetime", b.disk_properties[2]); v.f64("alpha", b.disk_properties[3]);
            v.f64("surface_density_normalization", b.disk_properties[4]); v.f64("mean_molecular_weight", b.disk_properties[5]);
            v.end_struct();
        }
        v.enum_end(role == PB200_ROLE_CENTRAL);
        b.disk_role = role;
        v.begin_struct("parameters");
        v.begin_struct("internal");
        zero_f64(v, "distance"); zero_f64(v, "norm_velocity_vector"); zero_f64(v, "norm_velocity_vector_2"); zero_f64(v, "migration_timescale");
        v.end_struct();
        v.begin_struct("output"); zero_axes(v, "acceleration"); v.end_struct();
        v.end_struct();
        zero_coordinates(v);
        v.end_struct();
    }
    {   // EvolutionType (evolution.rs:8-17)
        int t = live ? b.evolution_type : PB200_EVO_NONEVOLVING;
        t = v.enum_begin("evolution", t, kEvoNames, 7, t != PB200_EVO_NONEVOLVING);
        if (t == PB200_EVO_LECONTECHABRIER2013) { bool f = b.evolution_parameter != 0.; v.boolean(nullptr, f); b.evolution_parameter = f ? 1. : 0.; }
        else if (t != PB200_EVO_NONEVOLVING) v.f64(nullptr, b.evolution_parameter);
        v.enum_end(t != PB200_EVO_NONEVOLVING);
        b.evolution_type = t;
    }
    v.end_struct();
}

template <class V> void walk_evolver(V& v, Image& img, int i) {
    pb200_body_t& b = img.c.bodies[i];
    v.begin_struct(nullptr);
    {
        // the evolver carries its own copy of the type (evolution.rs:20); NonEvolving for slots without a table
        bool has = !img.evo[i][0].empty();
        int t = has ? b.evolution_type : PB200_EVO_NONEVOLVING;
        double param = b.evolution_parameter;
        t = v.enum_begin("evolution", t, kEvoNames, 7, t != PB200_EVO_NONEVOLVING);
        if (t == PB200_EVO_LECONTECHABRIER2013) { bool f = param != 0.; v.boolean(nullptr, f); }
        else if (t != PB200_EVO_NONEVOLVING) v.f64(nullptr, param);
        v.enum_end(t != PB200_EVO_NONEVOLVING);
    }
    static const char* const names[5] = {"time", "radius", "radius_of_gyration_2", "love_number", "inverse_tidal_q_factor"};
    for (int c = 0; c < 5; c++) {
        std::vector<double>& col = img.evo[i][c];
        uint64_t n = col.size();
        v.begin_vec(names[c], n);
        if (V::reading) col.resize(n);
        for (uint64_t k = 0; k < n; k++) v.f64(nullptr, col[k]);
        v.end_vec();
    }
    v.u64("left_index", img.evo_left_index[i]);
    v.end_struct();
}

template <class V> void walk_image(V& v, Image& img) {
    pb200_case_t& c = img.c;
    v.begin_struct(nullptr);
    v.f64("time_step", c.time_step); v.f64("half_time_step", c.half_time_step);
    v.begin_struct("universe");
    v.f64("initial_time", c.initial_time); v.f64("time_limit", c.time_limit);
    v.begin_array("particles");
    for (int i = 0; i < PB200_MAX_PARTICLES; i++) walk_particle(v, img, i);
    v.end_array();
    uint64_t nev = PB200_MAX_PARTICLES;
    v.begin_vec("particles_evolvers", nev);
    if (nev != PB200_MAX_PARTICLES) throw std::runtime_error("image: particles_evolvers must have MAX_PARTICLES entries");
    for (int i = 0; i < PB200_MAX_PARTICLES; i++) walk_evolver(v, img, i);
    v.end_vec();
    uint64_t n = (uint64_t)c.n_particles; v.u64("n_particles", n); c.n_particles = (int32_t)n;
    v.begin_struct("consider_effects");
    bool t = c.consider_tides, f = c.consider_rotational_flattening, g = c.consider_general_relativity, d = c.consider_disk, w = c.consider_wind, e = c.consider_evolution;
    v.boolean("tides", t); v.boolean("rotational_flattening", f); v.boolean("general_relativity", g); v.boolean("disk", d); v.boolean("wind", w); v.boolean("evolution", e);
    c.consider_tides = t; c.consider_rotational_flattening = f; c.consider_general_relativity = g; c.consider_disk = d; c.consider_wind = w; c.consider_evolution = e;
    v.end_struct();
    {
        int impl = v.enum_begin("general_relativity_implementation", c.general_relativity_implementation, kGrNames, 4, false);
        v.enum_end(false);
        c.general_relativity_implementation = impl;
    }
    v.begin_struct("hosts");
    v.begin_struct("index");
    uint64_t hm = (uint64_t)c.host_most_massive, ht = (uint64_t)c.host_tides, hf = (uint64_t)c.host_rotational_flattening, hg = (uint64_t)c.host_general_relativity, hd = (uint64_t)c.host_disk;
    v.u64("most_massive", hm); v.u64("tides", ht); v.u64("rotational_flattening", hf); v.u64("general_relativity", hg); v.u64("disk", hd);
    c.host_most_massive = (int32_t)hm; c.host_tides = (int32_t)ht; c.host_rotational_flattening = (int32_t)hf; c.host_general_relativity = (int32_t)hg; c.host_disk = (int32_t)hd;
    v.end_struct();
    v.begin_struct("most_massive");
    // HostMostMassive as find_indices computes it (universe.rs:981-988)
    bool mt = c.consider_tides && hm == ht, mf = c.consider_rotational_flattening && hm == hf, mg = c.consider_general_relativity && hm == hg;
    bool md = c.consider_disk && hm == hd;
    bool all = (mt || !c.consider_tides) && (mf || !c.consider_rotational_flattening) && (mg || !c.consider_general_relativity) && (md || !c.consider_disk);
    v.boolean("all", all); v.boolean("general_relativity", mg); v.boolean("tides", mt); v.boolean("rotational_flattening", mf); v.boolean("disk", md);
    v.end_struct();
    v.end_struct();
    {   // HashMap<usize, f64> (universe.rs:61): entries in key order (upstream's order is the hasher's, i.e. arbitrary)
        const int nkeys = PB200_MAX_PARTICLES * PB200_MAX_PARTICLES;
        uint64_t nmap = 0;
        if (!V::reading) for (int k = 0; k < nkeys; k++) if (c.pair_dependent_scaled_dissipation_factor[k] == c.pair_dependent_scaled_dissipation_factor[k]) nmap++;
        v.begin_map("pair_dependent_scaled_dissipation_factor", nmap);
        if (V::reading) {
            for (int k = 0; k < nkeys; k++) c.pair_dependent_scaled_dissipation_factor[k] = NAN;
            for (uint64_t e = 0; e < nmap; e++) {
                uint64_t key = 0; double value = 0.;
                v.map_entry(key, value);
                if (key >= (uint64_t)nkeys) throw std::runtime_error("image: pair_dependent_scaled_dissipation_factor key out of range");
                c.pair_dependent_scaled_dissipation_factor[key] = value;
            }
        } else {
            for (int k = 0; k < nkeys; k++) {
                double value = c.pair_dependent_scaled_dissipation_factor[k];
                if (value != value) continue;
                uint64_t key = (uint64_t)k;
                v.map_entry(key, value);
            }
        }
        v.end_map();
    }
    v.begin_array("roche_radiuses");
    for (int i = 0; i < PB200_MAX_PARTICLES * PB200_MAX_PARTICLES; i++) v.f64(nullptr, c.roche_radiuses[i]);
    v.end_array();
    v.end_struct();
    v.f64("current_time", c.current_time);
    v.u64("current_iteration", c.current_iteration);
    v.f64("recovery_snapshot_period", c.recovery_snapshot_period); v.f64("historic_snapshot_period", c.historic_snapshot_period);
    v.f64("last_recovery_snapshot_time", c.last_recovery_snapshot_time); v.f64("last_historic_snapshot_time", c.last_historic_snapshot_time);
    v.u64("n_historic_snapshots", c.n_historic_snapshots);
    v.u64("hash", img.hash);   // SipHash of the Debug string upstream; never checked on restore: 0 is written
    v.begin_array("particles_alternative_coordinates");
    for (int i = 0; i < PB200_MAX_PARTICLES; i++) {
        v.begin_struct(nullptr);
        zero_f64(v, "mass"); zero_f64(v, "mass_g"); zero_axes(v, "position"); zero_axes(v, "velocity"); zero_axes(v, "acceleration");
        v.end_struct();
    }
    v.end_array();
    {
        int ct = v.enum_begin("alternative_coordinates_type", c.coordinates_type, kCoordNames, 3, false);
        v.enum_end(false);
        c.coordinates_type = ct;
    }
    v.u64("timestep_warning", c.timestep_warning);
    v.begin_array("inertial_velocity_errors");
    for (int i = 0; i < PB200_MAX_PARTICLES; i++) axes(v, nullptr, c.inertial_velocity_errors[i]);
    v.end_array();
    v.begin_array("particle_angular_momentum_errors");
    for (int i = 0; i < PB200_MAX_PARTICLES; i++) axes(v, nullptr, c.particle_angular_momentum_errors[i]);
    v.end_array();
    v.end_struct();
}

// ---------------------------------------------------------------- JSON reader (key based)
using pbjson::Value;

int index_of(const char* const* names, int n, const std::string& s, const char* what) {
    for (int i = 0; i < n; i++) if (s == names[i]) return i;
    throw std::runtime_error(std::string("unknown ") + what + " '" + s + "'");
}
// serde externally tagged enum: "Unit" | {"Variant": payload}
std::pair<std::string, const Value*> variant(const Value& v) {
    if (v.kind == Value::String) return {v.str, nullptr};
    if (v.kind == Value::Object && v.obj.size() == 1) return {v.obj[0].first, v.obj[0].second.get()};
    throw std::runtime_error("unrecognised enum encoding in the JSON image");
}
void read_axes(const Value& v, double* a) { a[0] = v.at("x").number(); a[1] = v.at("y").number(); a[2] = v.at("z").number(); }

void image_from_json(const Value& d, Image& img) {
    if (!d.find("alternative_coordinates_type") || !d.find("universe"))
        throw Unsupported("only the WHFast integrator image is supported (this is an IAS15 / LeapFrog or unknown image)");
    pb200_case_t& c = img.c;
    const Value& u = d.at("universe");
    c.time_step = d.at("time_step").number(); c.half_time_step = d.at("half_time_step").number();
    c.initial_time = u.at("initial_time").number(); c.time_limit = u.at("time_limit").number();
    c.current_time = d.at("current_time").number();
    c.recovery_snapshot_period = d.at("recovery_snapshot_period").number(); c.historic_snapshot_period = d.at("historic_snapshot_period").number();
    c.last_recovery_snapshot_time = d.at("last_recovery_snapshot_time").number(); c.last_historic_snapshot_time = d.at("last_historic_snapshot_time").number();
    c.current_iteration = d.at("current_iteration").u; c.n_historic_snapshots = d.at("n_historic_snapshots").u; c.timestep_warning = d.at("timestep_warning").u;
    c.coordinates_type = index_of(kCoordNames, 3, d.at("alternative_coordinates_type").str, "coordinates type");
    c.n_particles = (int32_t)u.at("n_particles").u;
    if (c.n_particles < 1 || c.n_particles > PB200_MAX_PARTICLES) throw std::runtime_error("n_particles out of range");
    const Value& ce = u.at("consider_effects");
    c.consider_tides = ce.at("tides").boolean(); c.consider_rotational_flattening = ce.at("rotational_flattening").boolean();
    c.consider_general_relativity = ce.at("general_relativity").boolean(); c.consider_disk = ce.at("disk").boolean();
    c.consider_wind = ce.at("wind").boolean(); c.consider_evolution = ce.at("evolution").boolean();
    if (c.consider_disk) throw Unsupported("disk interaction is outside the B200 hot path");
    c.general_relativity_implementation = index_of(kGrNames, 4, u.at("general_relativity_implementation").str, "GR implementation");
    const Value& hi = u.at("hosts").at("index");
    c.host_most_massive = (int32_t)hi.at("most_massive").u; c.host_tides = (int32_t)hi.at("tides").u;
    c.host_rotational_flattening = (int32_t)hi.at("rotational_flattening").u; c.host_general_relativity = (int32_t)hi.at("general_relativity").u;
    c.host_disk = (int32_t)hi.at("disk").u;
    for (int k = 0; k < PB200_MAX_PARTICLES * PB200_MAX_PARTICLES; k++) c.pair_dependent_scaled_dissipation_factor[k] = NAN;
    if (const Value* pm = u.find("pair_dependent_scaled_dissipation_factor")) {
        for (const auto& kv : pm->obj) {
            char* end = nullptr;
            unsigned long long key = std::strtoull(kv.first.c_str(), &end, 10);
            if (!end || *end || key >= (unsigned long long)(PB200_MAX_PARTICLES * PB200_MAX_PARTICLES)) throw std::runtime_error("pair_dependent_scaled_dissipation_factor key out of range");
            c.pair_dependent_scaled_dissipation_factor[key] = kv.second->number();
        }
    }
    for (int i = 0; i < c.n_particles; i++) {
        const Value& p = u.at("particles").at(i);
        pb200_body_t& b = c.bodies[i];
        b.id = (int32_t)p.at("id").u;
        b.mass = p.at("mass").number(); b.mass_g = p.at("mass_g").number(); b.radius = p.at("radius").number();
        b.radius_of_gyration_2 = p.at("radius_of_gyration_2").number(); b.moment_of_inertia = p.at("moment_of_inertia").number();
        read_axes(p.at("inertial_position"), b.inertial_position); read_axes(p.at("inertial_velocity"), b.inertial_velocity);
        read_axes(p.at("inertial_acceleration"), b.inertial_acceleration);
        read_axes(p.at("heliocentric_position"), b.heliocentric_position); read_axes(p.at("heliocentric_velocity"), b.heliocentric_velocity);
        read_axes(p.at("spin"), b.spin); read_axes(p.at("angular_momentum"), b.angular_momentum);
        {
            auto ref = variant(p.at("reference"));
            b.reference = ref.second ? (int32_t)ref.second->u : -1;
        }
        {
            auto eff = variant(p.at("tides").at("effect"));
            b.tides_role = index_of(kRoleNames, 3, eff.first, "tides effect");
            if (eff.second) {
                auto model = variant(*eff.second);
                if (model.first != "ConstantTimeLag") {
                    if (c.consider_tides) throw Unsupported("tidal model " + model.first + " is outside the B200 hot path (ConstantTimeLag only)");
                    b.tides_role = PB200_ROLE_DISABLED;
                } else {
                    b.tides_dissipation_factor = model.second->at("dissipation_factor").number();
                    b.tides_dissipation_factor_scale = model.second->at("dissipation_factor_scale").number();
                    b.tides_love_number = model.second->at("love_number").number();
                }
            }
            const Value& ti = p.at("tides").at("parameters").at("internal");
            b.tides_scaled_dissipation_factor = ti.at("scaled_dissipation_factor").number();
            b.tides_lag_angle = ti.at("lag_angle").number();
            const Value& de = ti.at("denergy_dt");
            b.tides_denergy_dt = de.kind == Value::Number ? de.num : NAN;   // serde_json writes null for NaN
        }
        {
            auto eff = variant(p.at("rotational_flattening").at("effect"));
            b.flattening_role = index_of(kRoleNames, 3, eff.first, "rotational flattening effect");
            if (eff.second) {
                auto model = variant(*eff.second);
                if (model.first != "OblateSpheroid") {
                    if (c.consider_rotational_flattening) throw Unsupported("rotational flattening model " + model.first + " is outside the B200 hot path (OblateSpheroid only)");
                    b.flattening_role = PB200_ROLE_DISABLED;
                } else b.flattening_love_number = model.second->at("love_number").number();
            }
        }
        {
            auto eff = variant(p.at("general_relativity").at("effect"));
            b.general_relativity_role = index_of(kRoleNames, 3, eff.first, "general relativity effect");
            b.general_relativity_factor = p.at("general_relativity").at("parameters").at("internal").at("factor").number();
        }
        {
            auto eff = variant(p.at("wind").at("effect"));
            b.wind_role = index_of(kWindNames, 2, eff.first, "wind effect");
            const Value& in = p.at("wind").at("parameters").at("input");
            b.wind_k_factor = in.at("k_factor").number(); b.wind_rotation_saturation = in.at("rotation_saturation").number();
        }
        {
            auto eff = variant(p.at("disk").at("effect"));
            b.disk_role = index_of(kRoleNames, 3, eff.first, "disk effect");
            if (eff.second) {
                static const char* const keys[6] = {"inner_edge_distance", "outer_edge_distance", "lifetime", "alpha", "surface_density_normalization", "mean_molecular_weight"};
                for (int k = 0; k < 6; k++) b.disk_properties[k] = eff.second->at(keys[k]).number();
            }
        }
        {
            auto evo = variant(p.at("evolution"));
            b.evolution_type = index_of(kEvoNames, 7, evo.first, "evolution type");
            b.evolution_parameter = 0.;
            if (evo.second) b.evolution_parameter = evo.second->kind == Value::Bool ? (evo.second->b ? 1. : 0.) : evo.second->number();
        }
        const Value& ev = u.at("particles_evolvers").at(i);
        static const char* const names[5] = {"time", "radius", "radius_of_gyration_2", "love_number", "inverse_tidal_q_factor"};
        for (int k = 0; k < 5; k++) {
            const Value& col = ev.at(names[k]);
            img.evo[i][k].resize(col.arr.size());
            for (size_t r = 0; r < col.arr.size(); r++) img.evo[i][k][r] = col.arr[r]->number();
        }
        img.evo_left_index[i] = ev.find("left_index") ? ev.at("left_index").u : 0;
        read_axes(d.at("inertial_velocity_errors").at(i), c.inertial_velocity_errors[i]);
        read_axes(d.at("particle_angular_momentum_errors").at(i), c.particle_angular_momentum_errors[i]);
    }
    const Value& rr = u.at("roche_radiuses");
    for (int k = 0; k < PB200_MAX_PARTICLES * PB200_MAX_PARTICLES && k < (int)rr.arr.size(); k++) c.roche_radiuses[k] = rr.arr[k]->number();
}

// image -> flat case + table store (one table per evolving body, like posidonius_b200/case.py)
void finish_case(Image& img, pb200_case_t* out, pb200_table_store_t* store) {
    pb200_case_t& c = img.c;
    for (int i = 0; i < PB200_MAX_PARTICLES; i++) {
        pb200_body_t& b = c.bodies[i];
        if (i >= c.n_particles) {
            // unused slots: Particle::new_dummy() (particle.rs:100-146)
            std::memset(&b, 0, sizeof b);
            b.tides_role = b.flattening_role = b.general_relativity_role = b.disk_role = PB200_ROLE_DISABLED;
            b.wind_role = 1;
            b.evolution_type = PB200_EVO_NONEVOLVING;
            b.evolution_table = -1;
            b.reference = -1;
            continue;
        }
        b.evolution_table = -1;
        b.evolution_left_index = (int32_t)img.evo_left_index[i];
        if (b.evolution_type != PB200_EVO_NONEVOLVING && c.consider_evolution) {
            if (img.evo[i][0].empty()) throw std::runtime_error("evolving body with an empty time table");
            b.evolution_table = (int32_t)(store->columns.size() / 5);
            for (int k = 0; k < 5; k++) store->columns.push_back(img.evo[i][k]);
        }
    }
    store->rebuild();
    *out = c;
}

bool ends_with(const std::string& s, const char* suffix) {
    size_t n = std::strlen(suffix);
    return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}

bool read_file(const char* path, std::string& out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::ostringstream ss;
    ss << f.rdbuf();
    out = ss.str();
    return true;
}

}  // namespace

extern "C" {

int pb200_case_load(const char* path, pb200_case_t* out, pb200_table_store_t** tables_out) {
    if (!path || !out || !tables_out) return pb200_set_error_message(PB200_E_INVALID, "null argument");
    std::string data;
    if (!read_file(path, data)) return pb200_set_error_message(PB200_E_INVALID, std::string("cannot open ") + path + ": " + std::strerror(errno));
    pb200_table_store_t* store = new pb200_table_store_t();
    try {
        Image img;
        if (ends_with(path, ".json")) {
            pbjson::Parser parser(data);
            pbjson::Ptr root = parser.parse();
            image_from_json(*root, img);
        } else {
            BinReader r(data);
            walk_image(r, img);
            if (r.p != data.size()) throw Unsupported("the bincode image is not a WHFast image (IAS15 / LeapFrog or a different MAX_PARTICLES build)");
            if (img.c.consider_disk) throw Unsupported("disk interaction is outside the B200 hot path");
        }
        finish_case(img, out, store);
    } catch (const Unsupported& e) {
        delete store;
        return pb200_set_error_message(PB200_E_UNSUPPORTED, e.what());
    } catch (const std::exception& e) {
        delete store;
        return pb200_set_error_message(PB200_E_INVALID, std::string(path) + ": " + e.what());
    }
    *tables_out = store;
    return PB200_OK;
}

int pb200_case_save(const char* path, const pb200_case_t* c, const pb200_table_t* tables, size_t n_tables) {
    if (!path || !c) return pb200_set_error_message(PB200_E_INVALID, "null argument");
    try {
        Image img;
        img.c = *c;
        for (int i = 0; i < c->n_particles; i++) {
            const pb200_body_t& b = c->bodies[i];
            img.evo_left_index[i] = (uint64_t)(b.evolution_left_index < 0 ? 0 : b.evolution_left_index);
            if (b.evolution_table >= 0) {
                if ((size_t)b.evolution_table >= n_tables || !tables) throw std::runtime_error("evolution table index out of range");
                const pb200_table_t& t = tables[b.evolution_table];
                const double* cols[5] = {t.time, t.radius, t.radius_of_gyration_2, t.love_number, t.inverse_tidal_q_factor};
                for (int k = 0; k < 5; k++) if (cols[k]) img.evo[i][k].assign(cols[k], cols[k] + t.n_rows);
            }
        }
        std::string bytes;
        std::string p(path);
        if (ends_with(p, ".json")) { JsonWriter w; walk_image(w, img); bytes.swap(w.w.out); }
        else { BinWriter w; walk_image(w, img); bytes.swap(w.out); }
        // keep one backup per 12 hours, like output.rs:63-69 ("[year][month][day]T[period]" + ".bin")
        struct stat st;
        if (stat(path, &st) == 0) {
            time_t now = time(nullptr);
            struct tm g;
            gmtime_r(&now, &g);
            char stamp[32];
            snprintf(stamp, sizeof stamp, "%04d%02d%02dT%s.bin", g.tm_year + 1900, g.tm_mon + 1, g.tm_mday, g.tm_hour < 12 ? "AM" : "PM");
            size_t dot = p.find_last_of('.');
            size_t slash = p.find_last_of('/');
            std::string stem = (dot != std::string::npos && (slash == std::string::npos || dot > slash)) ? p.substr(0, dot) : p;
            std::rename(path, (stem + "." + stamp).c_str());
        }
        std::ofstream f(path, std::ios::binary | std::ios::trunc);
        if (!f) throw std::runtime_error(std::string("cannot create ") + path + ": " + std::strerror(errno));
        f.write(bytes.data(), (std::streamsize)bytes.size());
        if (!f) throw std::runtime_error(std::string("write failed on ") + path);
    } catch (const Unsupported& e) {
        return pb200_set_error_message(PB200_E_UNSUPPORTED, e.what());
    } catch (const std::exception& e) {
        return pb200_set_error_message(PB200_E_INVALID, e.what());
    }
    return PB200_OK;
}

const pb200_table_t* pb200_table_store_tables(const pb200_table_store_t* s) { return s && !s->tables.empty() ? s->tables.data() : nullptr; }
size_t pb200_table_store_count(const pb200_table_store_t* s) { return s ? s->tables.size() : 0; }
void pb200_table_store_free(pb200_table_store_t* s) { delete s; }

}  // extern "C"
