#!/usr/bin/env python
"""Executed warp-instructions of the step kernel by CALL SITE in the kernel body (outermost frame of nvdisasm's inline
chains) and by the innermost function, split into FP64 and other instructions.

usage: ncu_by_phase.py report.ncu-rep libposidonius_b200.so 'kernel-substring' warp_steps [outer-frame file:line -> opcode histogram of that phase]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


STEP_FILE = os.environ.get("PB200_STEP_FILE", "whfast_step.cuh")   # small_step.cuh for the lane = planet kernel


def main():
    rep, so, kern = sys.argv[1:4]
    warp_steps = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
    focus = sys.argv[5] if len(sys.argv) > 5 else None
    ophist = collections.Counter()
    # One cubin per translation unit. Units compiled from the same source (kernels_tu.cu, kernels_small_tu.cu with different
    # -D flags) extract to the same file name from the linked library, so the per-unit objects next to it are searched instead.
    objdir = os.path.join(os.path.dirname(os.path.abspath(so)), "build" if os.path.basename(so) == "libposidonius_b200.so" else "build_" + os.path.basename(so))
    objs = sorted(os.path.join(objdir, f) for f in os.listdir(objdir) if f.endswith(".o")) if os.path.isdir(objdir) else []
    dis = ""
    for obj in objs + [os.path.abspath(so)]:
        tmp = tempfile.mkdtemp()
        subprocess.call(["cuobjdump", "-xelf", "all", obj], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        for f in sorted(os.listdir(tmp)):
            if not f.endswith(".cubin"):
                continue
            names = subprocess.run(["cuobjdump", "-elf", os.path.join(tmp, f)], stdout=subprocess.PIPE, text=True).stdout
            if kern in names:
                dis = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, f)], stdout=subprocess.PIPE, text=True).stdout
                break
        if dis:
            break
    addr2 = {}
    in_k = False
    chain = []
    pending = []
    for line in dis.splitlines():
        if line.startswith("//--------------------- .text."):
            in_k = kern in line
            continue
        if not in_k:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            pending.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            if pending:
                chain = pending
                pending = []
            addr2[int(m.group(1), 16)] = (chain, m.group(2))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ia, ie, isrc = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Source")
    base = None
    outer = collections.defaultdict(lambda: [0, 0, 0, 0])
    inner = collections.defaultdict(lambda: [0, 0, 0, 0])
    tot = [0, 0]
    for r in rows[2:]:
        try:
            a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
            ex = int(r[ie])
        except Exception:
            continue
        if base is None:
            base = a
        ch, _ = addr2.get(a - base, ([("?", 0)], ""))
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
        op = m.group(2) if m else "?"
        f64 = 1 if op.split(".")[0] in ("DADD", "DMUL", "DFMA", "DSETP", "MUFU") else 0
        # outermost frame that lies in whfast_step.cuh (the kernel body / midpoint / gravity)
        step_frames = [c for c in ch if c[0] == STEP_FILE]
        o = step_frames[-1] if step_frames else ch[-1]
        o2 = step_frames[0] if step_frames else ch[0]
        outer["%s:%d" % o][f64] += ex
        inner["%s:%d" % o2][f64] += ex
        smem = {"LDS": 2, "STS": 3}.get(op.split(".")[0])
        if smem:
            outer["%s:%d" % o][smem] += ex
            inner["%s:%d" % o2][smem] += ex
        tot[f64] += ex
        if focus and "%s:%d" % o == focus:
            ophist[op.split(".")[0]] += ex
    if focus:
        print("opcodes inside %s per warp-step: " % focus + ", ".join("%s %.1f" % (k, v / warp_steps) for k, v in ophist.most_common(25)))
    print("per warp-step: %.0f instructions, %.0f FP64" % ((tot[0] + tot[1]) / warp_steps, tot[1] / warp_steps))
    for name, table in (("kernel-body call site (outermost whfast_step.cuh frame)", outer), ("innermost whfast_step.cuh frame", inner)):
        print("\n%s: FP64 / other (of which LDS, STS) per warp-step" % name)
        for k, v in sorted(table.items(), key=lambda kv: -(kv[1][0] + kv[1][1]))[:40]:
            print("  %-28s %8.1f %8.1f %8.1f %8.1f" % (k, v[1] / warp_steps, v[0] / warp_steps, v[2] / warp_steps, v[3] / warp_steps))


if __name__ == "__main__":
    main()
