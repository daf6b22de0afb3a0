#!/usr/bin/env python
"""Compare the SASS of every lbft_event_loop_kernel instantiation in two builds of the library (addresses and
encodings stripped).  Usage: python tools/sass_compare.py OLD.so NEW.so"""
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    res, name, lines = {}, None, []
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            if name:
                res[name] = lines
            name, lines = m.group(1), []
        elif re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):  # instruction lines (offsets grow past 4 hex digits)
            lines.append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", re.sub(r"^\s*/\*[0-9a-f]{4,}\*/", "", ln)).strip())
    if name:
        res[name] = lines
    return res


def key(name):  # <NMAX, QMODE, FIXED, REC, RES>; REC / RES default to 0 for builds that predate them
    m = re.search(r"kernelILi(\d+)ELi(\d)ELb([01])E(?:Lb([01])E)?(?:Lb([01])E)?", name)
    return None if not m else (int(m.group(1)), int(m.group(2)), int(m.group(3)), int(m.group(4) or 0), int(m.group(5) or 0))


old = {key(k): v for k, v in kernels(sys.argv[1]).items() if key(k)}
new = {key(k): v for k, v in kernels(sys.argv[2]).items() if key(k)}
for k in sorted(new):
    if k not in old:
        verdict = "new instantiation"
    else:
        verdict = "IDENTICAL" if old[k] == new[k] else "differs (%d -> %d instructions)" % (len(old[k]), len(new[k]))
    print("<%d,%d,%d,%d,%d> %6d instructions  %s" % (k + (len(new[k]), verdict)))
