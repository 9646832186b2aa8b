#!/usr/bin/env python
"""Copies the reference's own golden vectors for the WHFast hot path into tests/golden/.

Run in the build container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py

For every in-scope fixture of /root/reference/tests/data it stores
  * the Rust-serialised initial integrator `case.json` (gzip, de-duplicated by content hash), and
  * the expected inertial position / velocity / acceleration of every particle after the run
    (`particle_<i>.json`, the reference's 1e-14 regression vectors, tests/common/universe.rs:48-72)
in manifest.json. Fixtures without a case.json (test_order-*) reuse the matching test_integrator-* case
(star at index 0, which is what the stored goldens were generated with).
"""
import gzip
import hashlib
import json
import os

REF = "/root/reference/tests/data"
HERE = os.path.dirname(os.path.abspath(__file__))

FIXTURES = {
    # fixture dir: case source dir (None = own case.json)
    "test_integrator-whfast_jacobi": None,
    "test_integrator-whfast_democraticheliocentric": None,
    "test_integrator-whfast_whds": None,
    "test_tides-enabled_tides": None,
    "test_tides-disabled_tides": None,
    "test_flattening-enabled_flattening": None,
    "test_flattening-disabled_flattening": None,
    "test_general_relativity-kidder1995": None,
    "test_general_relativity-anderson1975": None,
    "test_general_relativity-newhall1983": None,
    "test_general_relativity-none": None,
    "test_evolution-brown_dwarf_leconte2011": None,
    "test_evolution-brown_dwarf_non_evolving": None,
    "test_evolution-m_dwarf_baraffe1998": None,
    "test_evolution-m_dwarf_baraffe2015": None,
    "test_evolution-m_dwarf_non_evolving": None,
    "test_star_types-brown_dwarf": None,
    "test_star_types-m_dwarf": None,
    # solar-like hosts: stellar wind (wind.rs) on every one, dynamical tides (pair-dependent dissipation factors,
    # constant_time_lag.rs:20-165) for BolmontMathis2016 / GalletBolmont2017
    "test_evolution-solar_like_baraffe1998": None,
    "test_evolution-solar_like_baraffe2015": None,
    "test_evolution-solar_like_bolmontmathis2016": None,
    "test_evolution-solar_like_galletbolmont2017": None,
    "test_evolution-solar_like_non_evolving": None,
    "test_star_types-solar_like": None,
    "test_disk-disabled_disk": None,
    "test_order-whfast_jacobi": "test_integrator-whfast_jacobi",
    "test_order-whfast_democraticheliocentric": "test_integrator-whfast_democraticheliocentric",
    "test_order-whfast_whds": "test_integrator-whfast_whds",
}
# Out-of-scope fixtures kept to test the rejection path (error, no fallback).
# (Kaula / creep fixtures ship no case.json upstream; the tests synthesise those by editing a WHFast case.)
REJECT = ["test_integrator-ias15", "test_integrator-leapfrog", "test_disk-enabled_disk"]


def strip_tables(d):
    """Rejection fixtures only need the structure; drop big arrays to keep the repo small."""
    if isinstance(d, dict):
        return {k: strip_tables(v) for k, v in d.items()}
    if isinstance(d, list):
        if len(d) > 64 and all(isinstance(x, (int, float)) for x in d):
            return d[:4]
        return [strip_tables(x) for x in d]
    return d


def main():
    os.makedirs(os.path.join(HERE, "cases"), exist_ok=True)
    manifest = {"fixtures": {}, "reject": {}}
    for name, src in FIXTURES.items():
        case_dir = os.path.join(REF, src or name)
        raw = open(os.path.join(case_dir, "case.json"), "rb").read()
        sha = hashlib.sha256(raw).hexdigest()[:16]
        out = os.path.join(HERE, "cases", sha + ".json.gz")
        if not os.path.exists(out):
            with gzip.GzipFile(out, "wb", compresslevel=9, mtime=0) as f:
                f.write(raw)
        particles = []
        i = 0
        while os.path.exists(os.path.join(REF, name, "particle_%d.json" % i)):
            particles.append(json.load(open(os.path.join(REF, name, "particle_%d.json" % i))))
            i += 1
        manifest["fixtures"][name] = {"case": "cases/%s.json.gz" % sha, "particles": particles, "tolerance_abs": 1e-14}
    for name in REJECT:
        d = strip_tables(json.load(open(os.path.join(REF, name, "case.json"))))
        raw = json.dumps(d).encode()
        sha = hashlib.sha256(raw).hexdigest()[:16]
        out = os.path.join(HERE, "cases", sha + ".json.gz")
        with gzip.GzipFile(out, "wb", compresslevel=9, mtime=0) as f:
            f.write(raw)
        manifest["reject"][name] = {"case": "cases/%s.json.gz" % sha}
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print("wrote", len(manifest["fixtures"]), "fixtures,", len(manifest["reject"]), "reject cases")


if __name__ == "__main__":
    main()
