#!/usr/bin/env python
"""Static SASS instruction count (and opcode histogram with -v) of every kernel in the per-unit objects of the library.
usage: sass_sizes.py [-v] [substring]"""
import collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
verbose = "-v" in sys.argv
args = [a for a in sys.argv[1:] if a != "-v"]
sub = args[0] if args else ""
objdir = os.path.join(ROOT, "posidonius_b200", os.environ.get("PB200_OBJDIR", "build"))
for o in sorted(os.listdir(objdir)):
    if not o.endswith(".o"):
        continue
    tmp = tempfile.mkdtemp()
    subprocess.call(["cuobjdump", "-xelf", "all", os.path.join(objdir, o)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        dis = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, f)], stdout=subprocess.PIPE, text=True).stdout
        name, n, ops = None, 0, collections.Counter()
        def flush():
            if name and sub in name:
                print("%-10s %6d  %s" % (o, n, name[:110]))
                if verbose:
                    print("            " + ", ".join("%s %d" % kv for kv in ops.most_common(14)))
        for line in dis.splitlines():
            m = re.match(r"//-+ \.text\.(\S+)", line)
            if m:
                flush(); name, n, ops = m.group(1), 0, collections.Counter(); continue
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m:
                n += 1; ops[m.group(2)] += 1
        flush()
