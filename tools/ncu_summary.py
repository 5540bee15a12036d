"""Summarise an .ncu-rep: key raw metrics per kernel and the hottest source lines (development tool).
usage: python tools/ncu_summary.py report.ncu-rep [--lines N]"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
pat = re.compile(
    r"^(Kernel Name|gpu__time_duration.sum|dram__bytes_(read|write).sum|sm__cycles_elapsed.max|launch__(grid_size|block_size|registers_per_thread)|"
    r"l1tex__data_pipe_lsu_wavefronts_mem_shared(_op_(ld|st))?.sum(.pct_of_peak_sustained_elapsed)?|"
    r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared(_op_(ld|st))?.sum|smsp__inst_executed.sum|"
    r"smsp__inst_executed_op_shared_(ld|st).sum|sm__warps_active.avg.pct_of_peak_sustained_active|"
    r"sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active|smsp__issue_active.avg.pct_of_peak_sustained_active|"
    r"dram__throughput.avg.pct_of_peak_sustained_elapsed|l1tex__throughput.avg.pct_of_peak_sustained_elapsed|"
    r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio|sm__throughput.avg.pct_of_peak_sustained_elapsed)$")
for r in rows[2:]:
    for i, k in enumerate(hdr):
        if pat.match(k):
            v = r[i]
            if k.startswith("smsp__average_warps_issue_stalled") and float(v or 0) < 0.15:
                continue
            print(f"{k} = {v[:100]} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = [i for i, r in enumerate(rows) if "Source" in r and any("Warp Stall Sampling (All" in c for c in r)]
if h:
    H = rows[h[0]]
    isrc = H.index("Source")
    col = lambda name: next((i for i, c in enumerate(H) if c.startswith(name)), None)
    ist, iex = col("Warp Stall Sampling (All"), col("Instructions Executed")
    iwf = H.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in H else None
    body = [r for r in rows[h[0] + 1:] if len(r) == len(H) and r[isrc] != "Source"]
    tot = sum(float(r[ist] or 0) for r in body) or 1.0
    # position of every instruction in program order, and a coarse histogram over the kernel
    nb = 12
    per = (len(body) + nb - 1) // nb
    print("--- stall samples / executed instructions / shared wavefronts by program region ---")
    for b in range(nb):
        seg = body[b * per:(b + 1) * per]
        if not seg:
            continue
        print(f"  instr {b * per:5d}-{b * per + len(seg) - 1:5d}: stalls {100 * sum(float(r[ist] or 0) for r in seg) / tot:5.1f}%  "
              f"exec {sum(float(r[iex] or 0) for r in seg):12.0f}  wf {sum(float(r[iwf] or 0) for r in seg) if iwf is not None else 0:12.0f}")
    for i, r in enumerate(body):
        r.append(i)
    body.sort(key=lambda r: -float(r[ist] or 0))
    print(f"--- hottest SASS/source lines by stall samples (total {tot:.0f}) ---")
    for r in body[:nlines]:
        print(f"{100 * float(r[ist] or 0) / tot:5.1f}%  #{r[-1]:<5d} exec={r[iex]:>9}  wf={r[iwf] if iwf is not None else '':>9}  {r[isrc][:110]}")
