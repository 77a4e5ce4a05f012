"""Per-region warp-stall summary of an `ncu --set full --import-source on` capture of the CEM-iteration kernel.

usage: python scripts/stall_summary.py <report.ncu-rep> > profiles/r01_iter_c2_stalls.txt
Regions are located in the SASS by their signature instructions (MUFU.EX2 clusters = Mish / SimNorm / two-hot blocks,
LDTM.x32 = the row-moment pass).
"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]
col = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
ins = []
for r in rows[2:]:
    if len(r) < len(hdr) or not r[col["Address"]].startswith("0x"):
        continue
    ins.append(dict(a=int(r[col["Address"]], 16), src=r[col["Source"]], n=int(r[col["# Samples"]] or 0),
                    ie=int(r[col["Instructions Executed"]] or 0), st={s: int(r[col[s]] or 0) for s in stalls}))
base = ins[0]["a"]
total = sum(x["n"] for x in ins)
print(rows[0][1])
print(f"instructions {len(ins)}, warp samples {total}")


def summarize(seg, label):
    n = sum(x["n"] for x in seg)
    c = collections.Counter()
    for x in seg:
        for s, v in x["st"].items():
            c[s] += v
    print(f"\n== {label}: {len(seg)} instructions, {n} samples ({100 * n / total:.1f} % of all), "
          f"max warp-instructions executed {max(x['ie'] for x in seg)}")
    print("   " + ", ".join(f"{s[6:]} {100 * v / max(n, 1):.0f} %" for s, v in c.most_common(8)))
    for x in sorted(seg, key=lambda x: -x["n"])[:6]:
        top = max(x["st"].items(), key=lambda kv: kv[1])
        print(f"   +{x['a'] - base:06x} {x['n']:7d}  {top[0][6:]:16s} {x['src'][:72]}")


summarize(ins, "whole kernel")
ex = [i for i, x in enumerate(ins) if "MUFU.EX2" in x["src"]]
clusters, cur = [], [ex[0]]
for i in ex[1:]:
    if i - cur[-1] < 120:
        cur.append(i)
    else:
        clusters.append(cur); cur = [i]
clusters.append(cur)
for k, c in enumerate(clusters):
    seg = ins[max(c[0] - 60, 0):c[-1] + 80]
    if sum(x["n"] for x in seg) > 0.002 * total:
        summarize(seg, f"MUFU.EX2 block {k} at +{ins[c[0]]['a'] - base:#x} ({len(c)} ex2)")
l32 = [i for i, x in enumerate(ins) if "LDTM.x32" in x["src"]]
if l32:
    summarize(ins[max(l32[0] - 20, 0):l32[0] + 160], "row-moment pass (first LDTM.x32 block)")
