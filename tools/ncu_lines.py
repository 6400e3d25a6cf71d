"""Aggregate an ncu source page by CUDA source line: python tools/ncu_lines.py report.ncu-rep [kernel-substr] [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
sub = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur_file = cur_fn = None
hdr = None
agg = {}
seen_fn = set()
active = False
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        cur_fn = r[1]
        key = cur_fn
        active = sub in cur_fn
        continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if not active or hdr is None:
        continue
    if r[0] != "" and r[0].isdigit():
        ie = r[hdr["Instructions Executed"]]
        smp = r[hdr["# Samples"]]
        bar = r[hdr["stall_barrier"]] if "stall_barrier" in hdr else "0"
        exc = r[hdr["L1 Wavefronts Shared Excessive"]]
        try:
            k = (cur_fn[:60], cur_file, int(r[0]), r[1].strip()[:90])
            v = agg.setdefault(k, [0, 0, 0, 0])
            v[0] += int(ie); v[1] += int(smp); v[2] += int(bar); v[3] += int(exc or 0)
        except ValueError:
            pass
tot = sum(v[0] for v in agg.values()) or 1
tots = sum(v[1] for v in agg.values()) or 1
print(f"total warp-inst {tot}, samples {tots}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1 if len(sys.argv) > 4 else 0])[:top]:
    print(f"{v[0] / tot:6.1%} inst {v[1] / tots:6.1%} smp  bar {v[2]:6d} excess_smem {v[3]:9d}  {k[1]}:{k[2]}  {k[3]}")
