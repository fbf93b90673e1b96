#!/usr/bin/env python
"""Static SASS statistics of the k_step variants in a cubin / shared library: instructions, fp64 instructions, registers, stack.
usage: python scripts/sass_count.py <cubin-or-so> [name-filter]"""
import re, subprocess, sys, collections
path = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else "k_step<"
sass = subprocess.run(["cuobjdump", "-sass", path], stdout=subprocess.PIPE).stdout.decode()
res = subprocess.run(["cuobjdump", "--dump-resource-usage", path], stdout=subprocess.PIPE).stdout.decode()
dem = lambda m: subprocess.run(["c++filt", m], stdout=subprocess.PIPE).stdout.decode().strip()
regs = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+) STACK:(\d+)", line)
    if m and cur:
        regs[cur] = (int(m.group(1)), int(m.group(2)))
cur = None
cnt = collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        op = m.group(1).split(".")[0]
        cnt[cur][op] += 1
        cnt[cur]["_all"] += 1
for fn in sorted(cnt, key=dem):
    name = dem(fn)
    if flt not in name:
        continue
    c = cnt[fn]
    fp64 = sum(v for k, v in c.items() if k in ("DADD", "DMUL", "DFMA", "DSETP", "MUFU", "DMNMX"))
    mem = {k: c[k] for k in ("LDG", "STG", "LDL", "STL", "LDS", "STS") if c[k]}
    r = regs.get(fn, (None, None))
    print("%-70s instr %5d  fp64 %4d (DADD %d DMUL %d DFMA %d)  regs %s stack %s  %s" % (name.replace("luma::", "")[:70], c["_all"], fp64, c["DADD"], c["DMUL"], c["DFMA"], r[0], r[1], mem))
