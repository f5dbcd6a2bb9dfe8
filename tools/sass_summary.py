"""Per-kernel counts of the Blackwell-specific SASS instructions in libsres_b200.so (cuobjdump -sass):
UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store, UTCBAR = tcgen05.commit,
SYNCS = mbarrier ops, SHFL = warp shuffles.      python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
lib = os.path.join(ROOT, "super-resolution-climate_b200", "lib", "libsres_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCSHIFT", "SYNCS", "SHFL", "BAR.SYNC"]
cur, counts, order, regs = None, collections.defaultdict(collections.Counter), [], {}
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void ", "")
        order.append(cur)
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["total"] += 1
        for p in pat:
            if op == p or op.startswith(p + "."):
                counts[cur][p] += 1
        if op.startswith("UTCHMMA.2CTA"):
            counts[cur]["UTCHMMA.2CTA"] += 0
print(f"# SASS instruction counts per kernel of {os.path.relpath(lib, ROOT)} (cuobjdump -sass, sm_100a)\n")
cols = ["total"] + [p for p in pat if p != "UTCHMMA.2CTA"]
print("| kernel | " + " | ".join(cols) + " |")
print("|---|" + "---:|" * len(cols))
for k in order:
    c = counts[k]
    if c["UTCHMMA"] or c["UTMALDG"] or c["LDTM"] or "kernel" in k:
        print(f"| `{k}` | " + " | ".join(str(c[p]) for p in cols) + " |")
