"""Opcode census of liblatentknn.so per kernel: the Blackwell-specific instructions that prove the hot kernels are
tcgen05 / TMEM / TMA code (B200_PROFILING.md: UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP =
cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier, UCGABAR = cluster barrier).
    python tools/sass_census.py > profiles/r02_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "latent_rag_b200", "liblatentknn.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UBLKCP", "UBLKPF", "UTCBAR", "UTCATOMSWS", "SYNCS", "UCGABAR", "ELECT",
         "FMNMX3", "LDG.E.256", "STG.E.256", "STG.E.ENL2.256", "HMMA", "IMMA", "FFMA", "DFMA"]
kern = None
counts = collections.OrderedDict()
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*", "", name)
        kern = name
        counts[kern] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and kern:
        op = m.group(1)
        counts[kern]["_total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                counts[kern][w] += 1
print(f"# SASS opcode census of {os.path.relpath(lib, ROOT)} (cuobjdump -sass, sm_100a), per kernel")
print("# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk (TMA),")
print("# UBLKPF = cp.async.bulk.prefetch, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, UCGABAR = cluster barrier")
tot = collections.Counter()
for k, c in counts.items():
    keys = [w for w in WATCH if c.get(w)]
    if not any(w in c for w in ("UTCHMMA", "LDTM", "UBLKCP", "STTM")) and "--all" not in sys.argv:
        for w in keys:
            tot[w] += c[w]
        continue
    print(f"{k}\n    instructions {c['_total']}: " + ", ".join(f"{w} {c[w]}" for w in keys))
    for w in keys:
        tot[w] += c[w]
print("# library totals: " + ", ".join(f"{w} {tot[w]}" for w in WATCH if tot.get(w)))
