#!/usr/bin/env python
"""profiles/<tag>_sass_summary.txt: what the shipped libmct_b200.so contains -- architectures, kernels, and per kernel the counts of the
SASS mnemonics that show TMA bulk copies (UBLKCP), mbarrier traffic (SYNCS), cluster barriers (UCGABAR / BAR...CLUSTER via
barrier.cluster), distributed shared memory (st/ld .shared::cluster -> ST/LD with cluster address via MAPA), f64 and MUFU use.
usage: python tools/sass_summary.py r02"""
import collections, os, re, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "multicameratracking_b200", "libmct_b200.so")
elf = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = collections.OrderedDict([("UBLKCP", r"\bUBLKCP"), ("UTMALDG", r"\bUTMALDG"), ("SYNCS", r"\bSYNCS"), ("UCGABAR", r"\bUCGABAR"), ("MAPA", r"\bMAPA"),
                                ("BAR", r"\bBAR\."), ("ATOMS", r"\bATOMS"), ("LDS", r"\bLDS"), ("STS", r"\bSTS"), ("DFMA/DMUL/DADD", r"\bD(FMA|MUL|ADD)"),
                                ("MUFU", r"\bMUFU"), ("SHFL", r"\bSHFL"), ("REDUX", r"\bC?REDUX")])
cur, counts, order, ninst = None, {}, [], collections.Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); order.append(cur); continue
    if cur is None or "/*" not in line:
        continue
    if re.search(r"/\*[0-9a-f]{4,}\*/", line):
        ninst[cur] += 1
        for k, p in pats.items():
            if re.search(p, line):
                counts[cur][k] += 1
def demangle(n):
    return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
out = ["cuobjdump -lelf / -sass of multicameratracking_b200/libmct_b200.so", "", "embedded ELF images:"] + ["  " + l for l in elf.splitlines()]
out += ["", "kernels: %d" % len(order), "", "%-44s %7s " % ("kernel", "instrs") + " ".join("%8s" % k[:8] for k in pats)]
for f in sorted(order, key=lambda f: -ninst[f]):
    out.append("%-44s %7d " % (demangle(f)[:44], ninst[f]) + " ".join("%8d" % counts[f][k] for k in pats))
tot = collections.Counter()
for f in order:
    tot.update(counts[f])
out.append("%-44s %7d " % ("TOTAL", sum(ninst.values())) + " ".join("%8d" % tot[k] for k in pats))
libs = subprocess.run(["ldd", lib], capture_output=True, text=True).stdout
out += ["", "ldd (no cuBLAS / cuDNN / NCCL / Triton in the product library):"] + ["  " + l.strip() for l in libs.splitlines()]
open(os.path.join(root, "profiles", "%s_sass_summary.txt" % tag), "w").write("\n".join(out) + "\n")
print("\n".join(out[:14] + out[-14:]))
