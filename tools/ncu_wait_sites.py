"""Which call site do the barrier-wait samples of an .ncu-rep belong to?  SASS view: every sampled instruction is
attributed to the nearest FOLLOWING instruction whose source line is in the kernel file itself (the inlined
mbar_wait loops are followed by the caller's code).  python tools/ncu_wait_sites.py report.ncu-rep kernel_file.cu"""
import csv, io, subprocess, sys
rep, kfile = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
inst = []
for r in rows:
    if len(r) > 5 and r[0] == "Address":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    inst.append(r)
print(hdr[:8])
for r in inst[:3]: print(r[:8])
