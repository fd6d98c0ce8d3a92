#!/usr/bin/env python
"""Static SASS opcode histogram of selected kernels of an object file (cuobjdump -sass): what the committed
profiles/*_sass_opcodes.txt files hold.  usage: sass_opcodes.py <file.o> <substring of the mangled name> ..."""
import collections, re, subprocess, sys
obj, pats = sys.argv[1], sys.argv[2:]
txt = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True, check=True).stdout
print("# SASS opcode histogram, %s (cuobjdump -sass, nvcc 12.9 -gencode arch=compute_100a,code=sm_100a -O3; static counts)" % obj)
for m in re.finditer(r"Function : (\S+)\n(.*?)(?=\n\s*Function : |\Z)", txt, re.S):
    name, body = m.group(1), m.group(2)
    if pats and not any(p in name for p in pats):
        continue
    ops = collections.Counter()
    for line in body.splitlines():
        mm = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if mm:
            ops[mm.group(1)] += 1
    print("\n== %s  (%d instructions)" % (name, sum(ops.values())))
    for op, n in ops.most_common():
        print("%6d  %s" % (n, op))
