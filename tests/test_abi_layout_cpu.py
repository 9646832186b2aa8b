"""The three statements of the C ABI structs agree: the C header (gcc: sizeof / offsetof), the ctypes mirror
(posidonius_b200/abi.py) and the #[repr(C)] mirrors of the Rust shim (integration/gpu_whfast.rs, whose compile-time
assertions are read here because no Rust toolchain exists in this image)."""
import os
import re
import subprocess

from conftest import ROOT


def _c_layout(tmp_path):
    exe = tmp_path / "abi_layout"
    subprocess.check_call(["gcc", "-std=c11", "-o", str(exe), os.path.join(ROOT, "tests", "abi_layout.c")])
    sizes, offsets = {}, {}
    for line in subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True, check=True).stdout.splitlines():
        parts = line.split()
        if len(parts) == 2:
            sizes[parts[0]] = int(parts[1])
        else:
            offsets[parts[0]] = (int(parts[1]), int(parts[2]))
    return sizes, offsets


def test_header_ctypes_and_rust_layouts_agree(tmp_path):
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_layout", os.path.join(ROOT, "integration", "gen_layout.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    # the generated files are current
    assert open(os.path.join(ROOT, "tests", "abi_layout.c")).read() == gen.c_program()
    rust = open(os.path.join(ROOT, "integration", "gpu_whfast.rs")).read()
    assert gen.rust_block() in rust
    sizes, offsets = _c_layout(tmp_path)
    rust_sizes = {m.group(1): int(m.group(2)) for m in re.finditer(r"assert!\(size_of::<(\w+)>\(\) == (\d+)\)", rust)}
    rust_offsets = {(m.group(1), m.group(2)): int(m.group(3)) for m in re.finditer(r"assert!\(offset_of!\((\w+), (\w+)\) == (\d+)\)", rust)}
    n_fields = 0
    for c_name, rust_name, size, fields in gen.layout():
        assert sizes[c_name] == size == rust_sizes[rust_name], c_name
        for name, off, fsize in fields:
            assert offsets["%s.%s" % (c_name, name)] == (off, fsize), (c_name, name)
            assert rust_offsets[(rust_name, name)] == off, (rust_name, name)
            n_fields += 1
        # the Rust struct declares exactly these fields, in this order
        body = re.search(r"pub struct %s \{(.*?)\n\}" % rust_name, rust, flags=re.S).group(1)
        declared = re.findall(r"pub (\w+):", body)
        assert declared == [name for name, _, _ in fields], rust_name
    assert n_fields == len(rust_offsets) > 80


def test_rust_shim_binds_every_entry_point_it_uses_with_the_header_signature_names():
    """Every extern "C" fn of the shim is declared by the header (no stale or invented symbol)."""
    rust = open(os.path.join(ROOT, "integration", "gpu_whfast.rs")).read()
    header = open(os.path.join(ROOT, "include", "posidonius_b200.h")).read()
    block = rust[rust.index('extern "C" {'):]
    block = block[:block.index("\n}\n")]
    names = re.findall(r"fn (pb200_\w+)\(", block)
    assert len(names) >= 25
    for n in names:
        assert re.search(r"\b%s\s*\(" % n, header), n
    # the shim implements the whole trait (src/integrator/mod.rs:16-26)
    for method in ("as_any", "get_n_historic_snapshots", "get_n_particles", "get_current_time", "set_time_limit", "set_snapshot_periods",
                   "initialize_physical_values", "iterate", "write_recovery_snapshot"):
        assert re.search(r"fn %s\(" % method, rust), method
