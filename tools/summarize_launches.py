"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into the per-kernel share
table kept under profiles/.
    python tools/summarize_launches.py gpurun_out/launches.csv profiles/r01_launches_bench_10M.txt "<command>" """
import csv
import sys
from collections import defaultdict

src, out, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("=="))]
hdr = rows[0]
i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt, units = defaultdict(float), defaultdict(int), set()
for r in rows[1:]:
    if len(r) <= i_val:
        continue
    v = float(r[i_val].replace(",", ""))
    u = r[i_unit]
    units.add(u)
    v_ms = v / 1e6 if u == "ns" else v / 1e3 if u == "us" else v if u == "ms" else v * 1e3
    tot[r[i_name]] += v_ms
    cnt[r[i_name]] += 1
total = sum(tot.values())
lines = [f"# ncu --metrics gpu__time_duration.sum --clock-control none: {cmd}",
         "# launches of the whole bench process (index build + warm-up + timed steps + batch-1 passes); shares, not absolutes",
         f"# unit seen: {units}"]
for name, ms in sorted(tot.items(), key=lambda kv: -kv[1])[:14]:
    short = name.split("(")[0][-70:]
    lines.append(f"{ms:12.3f} ms  {100 * ms / total:6.2f}%  x{cnt[name]:<4d} {short}")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
