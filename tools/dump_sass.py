#!/usr/bin/env python
"""Write the SASS of selected kernels of csrc/liblbft_b200.so under profiles/ (cuobjdump -sass) plus an opcode census, so
that the instruction counts quoted in DESIGN.md are checkable.  Usage: python tools/dump_sass.py <tag> <substring> [...]
e.g.  python tools/dump_sass.py r2 'event_loop_kernelILi16ELi2ELi1ELb0ELb0ELb0ELb0ELi32' 'event_loop_kernelILi16ELi3ELi2ELb0ELb0ELb0ELb0ELi8' \
          'wide_kernelILi64ELi3ELb0ELi8ELb0ELi3' 'wide_kernelILi16ELi2ELb1ELi32ELb0ELi0'   (the kernels BASELINE configs 3, 5, 4 and 1-2 select)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("LBFT_LIB_PATH") or os.path.join(ROOT, "librabft_simulator_b200", "csrc", "liblbft_b200.so")
tag, pats = sys.argv[1], sys.argv[2:]
text = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True).stdout
funcs, cur = {}, None
for line in text.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur is not None:
        funcs[cur].append(line)
demangle = lambda n: subprocess.run(["c++filt", n], stdout=subprocess.PIPE, text=True).stdout.strip()
summary = []
for pat in pats:
    for name, lines in funcs.items():
        if pat not in name:
            continue
        pretty = demangle(name)
        short = re.sub(r"[^A-Za-z0-9]+", "_", pretty.split("(")[0].replace("void lbft::", "")).strip("_")
        ops = collections.Counter()
        n = 0
        for ln in lines:
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
            if m:
                n += 1
                ops[m.group(1).split(".")[0]] += 1
        path = os.path.join(ROOT, "profiles", "%s_sass_%s.txt" % (tag, short))
        with open(path, "w") as f:
            f.write("// %s\n// %d SASS instructions; cuobjdump -sass %s\n// (instruction encodings stripped: address, predicate, opcode and "
                    "operands only)\n" % (pretty, n, os.path.basename(LIB)))
            kept = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", ln).rstrip() for ln in lines]
            f.write("\n".join(ln for ln in kept if ln.strip()) + "\n")
        top = ", ".join("%s %d" % kv for kv in ops.most_common(14))
        mem = {k: ops[k] for k in ("LDG", "STG", "LDS", "STS", "LDL", "STL", "LDC", "ATOMG", "RED", "REDUX", "SHFL", "VOTE", "WARPSYNC", "BAR", "MUFU", "DMUL", "DADD", "DSETP", "IMAD", "LOP3")}
        summary.append("%s\n  %d instructions -> %s\n  memory / warp-collective / fp64 opcodes: %s\n  top opcodes: %s\n" % (
            pretty, n, os.path.relpath(path, ROOT), ", ".join("%s %d" % kv for kv in mem.items() if kv[1]), top))
out = os.path.join(ROOT, "profiles", "%s_sass_summary.txt" % tag)
with open(out, "w") as f:
    f.write("SASS census of the kernels the BASELINE configurations select (tools/dump_sass.py; no tensor-core or TMA opcodes are\n"
            "expected: the path has no dense contraction and no tile movement — integer state machine, a little fp64 for the ziggurat)\n\n")
    f.write("\n".join(summary))
print(open(out).read())
