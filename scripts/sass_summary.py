"""Per-kernel SASS opcode counts of gcge_b200/lib/libgcge_b200.so (cuobjdump -sass): the instructions that show which
hardware paths a kernel uses -- UTMALDG (TMA tensor-map loads), UBLKCP (TMA bulk copies), DMMA (FP64 tensor-core
MMA), SYNCS (mbarrier operations), plus LDS/STS/LDG/STG/DADD/DMUL/DFMA totals.  No GPU needed.

    python scripts/sass_summary.py > profiles/sass_summary_r2.txt
"""
import collections, re, subprocess, sys
from pathlib import Path

lib = Path(__file__).resolve().parents[1] / "gcge_b200" / "lib" / "libgcge_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
OPS = ["UTMALDG", "UBLKCP", "DMMA", "SYNCS", "LDS", "STS", "LDG", "STG", "DADD", "DMUL", "DFMA", "SHFL", "BAR"]
counts = collections.OrderedDict(); cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("(anonymous namespace)::", "")
        cur = counts.setdefault(name, collections.Counter()); continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        for o in OPS:
            if op == o or op.startswith(o + "."):
                cur[o] += 1
tot = collections.Counter()
print(f"# {lib.name}: {len(counts)} kernels (sm_100a).  tcgen05 / UTCMMA / LDTM do not appear: there is no FP64 tcgen05 kind; the FP64 tensor path is DMMA")
print(f"{'kernel':78s} " + " ".join(f"{o:>7s}" for o in OPS))
for k, c in counts.items():
    if not any(c[o] for o in ("UTMALDG", "UBLKCP", "DMMA", "SYNCS")):
        continue
    print(f"{k[:78]:78s} " + " ".join(f"{c[o]:7d}" for o in OPS))
    tot.update(c)
print(f"{'TOTAL of the kernels listed':78s} " + " ".join(f"{tot[o]:7d}" for o in OPS))
alltot = collections.Counter()
for c in counts.values():
    alltot.update(c)
print(f"{'TOTAL of all kernels':78s} " + " ".join(f"{alltot[o]:7d}" for o in OPS))
print("UTC*MMA / LDTM / tcgen05 opcodes in the library:", len(re.findall(r"UTC\w*MMA|LDTM|UTCBAR", sass)))
