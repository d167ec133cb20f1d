"""Aggregate the warp-state samples of an .ncu-rep by CUDA source line (needs -lineinfo + --import-source on).
    python tools/ncu_by_line.py report.ncu-rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
sec, hdr, agg = None, None, {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        sec = r[1]; continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r; i_s = hdr.index("# Samples"); continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        line = int(r[0]); n = int(r[i_s] or 0)
    except ValueError:
        continue
    a = agg.setdefault((sec.split("/")[-1], line), [0, r[1]])
    a[0] += n
tot = sum(v[0] for v in agg.values())
print("total samples", tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{v[0]:7d} {100.0 * v[0] / max(tot, 1):5.1f}%  {k[0]}:{k[1]:<5d} {v[1].strip()[:120]}")
