"""Split a kernel's executed instructions and stall samples at its barriers (the phases of a
block-cooperative kernel): python tools/ncu_segments.py report.ncu-rep kernel-substr [warps_per_cta items]"""
import collections
import csv
import re
import subprocess
import sys

rep, sub = sys.argv[1], sys.argv[2]
warps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
items = int(sys.argv[4]) if len(sys.argv) > 4 else 1
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
blocks, cur, name = [], None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        name = r[1]
        continue
    if r and r[0] == "Address":
        cur = {"name": name, "hdr": r, "data": []}
        blocks.append(cur)
        continue
    if cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
blk = [b for b in blocks if sub in b["name"]][-1]
H = {h: i for i, h in enumerate(blk["hdr"])}
data = blk["data"]
tot_i = sum(int(r[H["Instructions Executed"]]) for r in data)
tot_s = sum(int(r[H["# Samples"]]) for r in data)
print(f"{blk['name']}: {tot_i} warp-instructions ({tot_i / items:.0f} per item), {tot_s} samples")
start, acc_i, acc_s = 0, 0, 0
segs = []
for idx, r in enumerate(data):
    ie, sm = int(r[H["Instructions Executed"]]), int(r[H["# Samples"]])
    acc_i += ie
    acc_s += sm
    if "BAR.SYNC" in r[H["Source"]]:
        segs.append((start, idx, acc_i, acc_s, ie))
        start, acc_i, acc_s = idx + 1, 0, 0
segs.append((start, len(data) - 1, acc_i, acc_s, 0))
for s0, s1, ni, ns, bar in segs:
    c = collections.Counter()
    for r in data[s0:s1 + 1]:
        src = re.sub(r"^@!?U?P\d+\s+", "", r[H["Source"]].strip())
        c[src.split()[0].split(".")[0]] += int(r[H["Instructions Executed"]])
    top = ", ".join(f"{k} {100 * v / max(1, ni):.0f}%" for k, v in c.most_common(7))
    print(f"sass {s0:5d}-{s1:5d}  inst {100 * ni / tot_i:5.1f}%  samples {100 * ns / tot_s:5.1f}%  "
          f"barrier executions per item per warp {bar / items / warps:7.1f} | {top}")
