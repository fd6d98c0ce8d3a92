#!/usr/bin/env python
"""Static check of k_brute's FFMA2 stream: distinct register-file reads per instruction after the
operand-reuse cache (an FFMA2 with 3 uncached 64-bit sources needs a third register-file cycle)."""
import collections, re, subprocess, sys, os
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = os.path.join(root, "navlab-dpe-sdr_b200", "build", "dpe_brute.o")
log = open(os.path.join(root, "navlab-dpe-sdr_b200", "build", "dpe_brute.ptxas.log")).read()
name = re.search(r"_ZN3dpe7k_brute[A-Za-z0-9_]*", log).group(0)
sass = subprocess.run(["cuobjdump", "-sass", "-fun", name, obj], capture_output=True, text=True).stdout
ins = [m.group(2).strip() for m in (re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l) for l in sass.splitlines()) if m]
ff = [i for i, t in enumerate(ins) if "FFMA2" in t]
lo, hi = ff[0], ff[-1]
prev, hist, other = {}, collections.Counter(), collections.Counter()
for t in ins[lo:hi + 1]:
    if "FFMA2" not in t:
        other[t.split()[1] if t.startswith("@") else t.split()[0]] += 1
        prev = {}
        continue
    srcs = [o.strip() for o in t.split(None, 1)[1].split(",")][1:]
    reads, nxt = set(), {}
    for slot, o in enumerate(srcs):
        m = re.match(r"-?R(\d+)(\.reuse)?(\.F32x2\.HI_LO|\.F32)?", o)
        if not m:
            continue
        r, reuse, wide = int(m.group(1)), bool(m.group(2)), m.group(3) == ".F32x2.HI_LO"
        if prev.get(slot) != r:
            reads.add(r)
            if wide:
                reads.add(r + 1)
        if reuse:
            nxt[slot] = r
    hist[max(sum(1 for r in reads if r % 2 == 0), sum(1 for r in reads if r % 2))] += 1
    prev = nxt
n = sum(hist.values())
cyc = sum(max(2, k) * v for k, v in hist.items())
print("FFMA2 in loop: %d; bank reads histogram %s; modelled pipe efficiency %.1f%%; other instrs %d %s" %
      (n, sorted(hist.items()), 100.0 * 2 * n / cyc, sum(other.values()), other.most_common(8)))
