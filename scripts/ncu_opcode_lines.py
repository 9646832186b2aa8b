#!/usr/bin/env python
"""Executed warp-instructions of given opcodes by source line (innermost frame) of the step kernel.
usage: ncu_opcode_lines.py report.ncu-rep lib.so 'kernel-substring' warp_steps OP[,OP...] [top]"""
import collections, csv, io, os, re, subprocess, sys, tempfile


def main():
    rep, so, kern = sys.argv[1:4]
    warp_steps = float(sys.argv[4]); ops = set(sys.argv[5].split(",")); top = int(sys.argv[6]) if len(sys.argv) > 6 else 30
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], stdout=subprocess.PIPE, text=True).stdout
    addr2line, in_k, cur = {}, False, ("?", 0)
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
            addr2line[int(m.group(1), 16)] = cur
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ia, ie, isrc = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Source")
    base, cnt = None, collections.Counter()
    for r in rows[2:]:
        try:
            a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia]); ex = int(r[ie])
        except Exception:
            continue
        if base is None:
            base = a
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
        op = (m.group(2) if m else "?").split(".")[0]
        if op in ops:
            cnt["%s:%d" % addr2line.get(a - base, ("?", 0))] += ex
    print("%s: %.1f per warp-step" % (",".join(sorted(ops)), sum(cnt.values()) / warp_steps))
    for k, v in cnt.most_common(top):
        print("  %-28s %8.1f" % (k, v / warp_steps))


if __name__ == "__main__":
    main()
