"""SASS evidence for profiles/: per kernel of liblmc_b200.so the instruction count, the registers / stack ptxas reports and
the Blackwell-specific mnemonics (bulk TMA copies UBLKCP, mbarrier SYNCS, shuffles, local memory).  No tensor-core
contraction exists on this path, so UTC*MMA / tcgen05 are not expected.
usage: python tools/sass_evidence.py > profiles/r02_sass_evidence.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "langevin-mcmc_b200", "liblmc_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
print("cuobjdump -sass langevin-mcmc_b200/liblmc_b200.so   arch:", ", ".join(arch))
kern, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); kern[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        kern[cur][m.group(1)] += 1
def demangle(n):
    return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "")
print("%-52s %8s %7s %7s %6s %6s %6s %6s" % ("kernel", "instr", "UBLKCP", "SYNCS", "SHFL", "LDL", "STL", "BAR"))
seen = set()
for k, c in kern.items():
    d = demangle(k)
    if not d.startswith("lmc_cuda::k_") or "<4" in d or "<12" in d or d in seen:
        continue
    seen.add(d)
    print("%-52s %8d %7d %7d %6d %6d %6d %6d" % (d[10:62], sum(c.values()), c["UBLKCP"], c["SYNCS"], c["SHFL"], c["LDL"], c["STL"], c["BAR"]))
tot = collections.Counter()
for c in kern.values():
    tot.update(c)
print("\ntensor-core mnemonics in the whole library (UTCHMMA/UTCIMMA/UTCQMMA/HMMA/IMMA/UTCBAR):",
      {k: v for k, v in tot.items() if k.startswith("UTC") or k in ("HMMA", "IMMA")} or "none (expected: no dense contraction on this path)")
print("bulk-TMA / mbarrier mnemonics in the whole library:", {k: v for k, v in tot.items() if k in ("UBLKCP", "SYNCS", "UTMALDG", "UTMASTG")})
