"""DRAM bytes per row of the kernels of one prepared Gram pass, from an ncu --set full report (bench.py's
`roofline.traffic`).  usage: python tools/traffic_json.py report.ncu-rep ROWS > profiles/r2_ncu_gram_cells_traffic.json"""
import csv
import io
import json
import subprocess
import sys

rep, rows = sys.argv[1], int(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
tab = list(csv.reader(io.StringIO(raw)))
hdr, units = tab[0], tab[1]


def val(r, key):
    i = hdr.index(key)
    v = float(r[i].replace(",", ""))
    u = units[i].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1,
                "msecond": 1, "nsecond": 1e-6}.get(u, 1)


kernels, total = {}, 0.0
for r in tab[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").strip()
    if name in kernels:
        continue  # one launch of each kernel of the pass
    rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    kernels[name] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "duration_ms": val(r, "gpu__time_duration.sum")}
    total += rd + wr
print(json.dumps({"source": f"ncu --set full --clock-control none, tools/gpu_job_r2.sh (kbench config 2, {rows} rows, "
                            "prepared shard), summarised by tools/traffic_json.py", "rows": rows, "kernels": kernels,
                  "note": "per PIRLS iteration of a prepared shard: the IRLS step (stats_pass_kernel instance) reads X, y "
                          "(88 B per row) and writes omega, omega z (16 B); the pair kernel reads the stored t of every "
                          "term (8 B each), omega, omega z and the stored row order (2 B per row and role) once per L2 band",
                  "dram_bytes_per_row": total / rows}, indent=1))
