#!/usr/bin/env python
"""Joins an ncu report's per-SASS-instruction counters with nvdisasm line info: executed warp-instructions and
stall samples per source line / per inlined function of the step kernel.

usage: ncu_by_line.py report.ncu-rep libposidonius_b200.so 'kernel-substring' [top]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, so, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], stdout=subprocess.PIPE, text=True).stdout
    # locate the kernel's .text section
    addr2line = {}
    in_k = False
    cur = ("?", 0)
    for line in dis.splitlines():
        if line.startswith("//--------------------- .text."):
            in_k = kern in line
            continue
        if not in_k:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            addr2line[int(m.group(1), 16)] = (cur, m.group(2))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    base = None
    per_line = collections.Counter()
    samp_line = collections.Counter()
    total = tsamp = 0
    for r in rows[2:]:
        try:
            a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
            ex, sa = int(r[ie]), int(r[isamp])
        except Exception:
            continue
        if base is None:
            base = a
        key = addr2line.get(a - base, (("?", 0), ""))[0]
        per_line[key] += ex
        samp_line[key] += sa
        total += ex
        tsamp += sa
    print("total warp-instructions %d, samples %d" % (total, tsamp))
    print("%-28s %8s %8s" % ("file:line", "inst%", "stall%"))
    for key, c in per_line.most_common(top):
        print("%-28s %7.2f%% %7.2f%%" % ("%s:%d" % key, 100.0 * c / total, 100.0 * samp_line[key] / max(tsamp, 1)))
    per_file = collections.Counter()
    for key, c in per_line.items():
        per_file[key[0]] += c
    print({k: "%.1f%%" % (100.0 * v / total) for k, v in per_file.items()})


if __name__ == "__main__":
    main()
