"""DRAM bytes (and a few pipe metrics) of the captured CEM-iteration launch -> JSON (feeds bench.py's roofline.traffic).

usage: python scripts/extract_traffic.py <report.ncu-rep> <workload> <envs> <out.json> [<out2.json> ...]
"""
import csv, io, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tdmpc2_b200 import build

rep, wl, envs, outs = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4:]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[0]
vals = rows[-1]                       # rows[1] is the units row; the last row is the (single) captured launch


def get(name, default=None):
    try:
        return float(vals[hdr.index(name)].replace(",", ""))
    except Exception:
        return default


rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
doc = {
    "kernel": vals[hdr.index("Kernel Name")], "workload": wl, "envs": envs,
    "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
    "duration_under_ncu_ms": get("gpu__time_duration.sum") / 1e6,
    "sm__pipe_tensor_cycles_active_pct": get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    "sm__inst_executed_pipe_tensor": get("sm__inst_executed_pipe_tensor.sum"),
    "registers_per_thread": get("launch__registers_per_thread"),
    "lib_digest": build._digest(),
    "source": "ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum of ONE MODE_ITER launch",
}
for o in outs:
    with open(o, "w") as f:
        json.dump(doc, f, indent=1)
print(json.dumps(doc))
