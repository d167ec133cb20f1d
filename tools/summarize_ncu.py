"""Turn an .ncu-rep into the short text summary kept under profiles/.
    python tools/summarize_ncu.py gpurun_out/prof_b4096.ncu-rep profiles/r01_ncu_b4096.txt"""
import csv
import io
import re
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[-1]
keep = [
    r"^Kernel Name$", r"^gpu__time_duration\.sum$", r"^launch__grid_size$", r"^launch__block_size$",
    r"^launch__registers_per_thread$", r"^launch__shared_mem_per_block_dynamic$",
    r"^dram__bytes_read\.sum$", r"^dram__bytes_write\.sum$", r"^dram__bytes_read\.sum\.per_second$",
    r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^lts__t_sector_hit_rate\.pct$", r"^lts__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld\.sum$",
    r"sm__pipe_tensor_cycles_active_realtime\.avg\.pct_of_peak_sustained_elapsed$",
    r"^sm__inst_executed_pipe_tensor_subpipe_hmma\.avg\.pct_of_peak_sustained_active$",
    r"^l1tex__data_pipe_tc_wavefronts_mem_shared\.sum\.pct_of_peak_sustained_elapsed$",
    r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$", r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^sm__cycles_elapsed\.max$", r"^smsp__inst_executed\.sum$",
]
lines = [f"# ncu --set full --clock-control none summary of {rep}"]
for h, u, v in zip(hdr, units, vals):
    if any(re.search(k, h) for k in keep):
        lines.append(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
sh, sd = srows[1], srows[2:]
i_src, i_smp, i_ex = sh.index("Source"), sh.index("# Samples"), sh.index("Instructions Executed")
stall = [i for i, h in enumerate(sh) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[i_smp] or 0) for r in sd)
lines.append(f"\n# warp-state samples: {tot} total; top instructions (samples, executed, SASS, top stall reasons)")
for r in sorted(sd, key=lambda r: -int(r[i_smp] or 0))[:14]:
    st = sorted(((sh[i][6:], int(r[i] or 0)) for i in stall if int(r[i] or 0) > 0), key=lambda kv: -kv[1])[:2]
    lines.append(f"{r[i_smp]:>8} {r[i_ex]:>10}  {r[i_src][:72]:72}  {st}")
agg = {}
for r in sd:
    for i in stall:
        agg[sh[i][6:]] = agg.get(sh[i][6:], 0) + int(r[i] or 0)
lines.append("\n# stall reasons over all samples: " + ", ".join(f"{k}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
