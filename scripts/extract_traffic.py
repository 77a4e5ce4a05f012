"""DRAM bytes of the captured CEM-iteration launch -> JSON (feeds bench.py's roofline.traffic).

usage: python scripts/extract_traffic.py <report.ncu-rep> <out.json> [<out2.json> ...]
"""
import csv, io, json, subprocess, sys

rep, outs = sys.argv[1], sys.argv[2:]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[0]
vals = rows[-1]                       # rows[1] is the units row; the last row is the (single) captured launch
get = lambda name: float(vals[hdr.index(name)].replace(",", ""))
rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
name = vals[hdr.index("Kernel Name")]
dur_ns = get("gpu__time_duration.sum")
doc = {
    "kernel": f"{name} MODE_ITER, workload c2 (E=256, N=512, H=3), one launch",
    "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
    "duration_under_ncu_ms": dur_ns / 1e6,
    "source": "ncu --set full --clock-control none (scripts/gpu_final.sh), dram__bytes_read.sum + dram__bytes_write.sum",
    "algorithmic_hbm_bytes_per_launch": 90600000,
    "note": "algorithmic = fresh noise (E*(H*(N-P)+N)*A*4 B = 77.9 MB) + packed weights once (12.7 MB); everything else is "
            "the per-CTA activation scratch (148 slots, L2-resident) spilling to DRAM",
}
for o in outs:
    with open(o, "w") as f:
        json.dump(doc, f, indent=1)
print(json.dumps(doc))
