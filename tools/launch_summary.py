"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel:
python tools/launch_summary.py launches.csv "<command that was profiled>" [bench.json] > summary.md"""
import collections
import csv
import json
import re
import sys

path, cmd = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(path)) if len(r) >= 15 and r[0].isdigit()]
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r"^void ", "", r[4])
    name = re.sub(r"\(.*$", "", name).replace("(int)", "").replace("(bool)", "")
    ns = float(r[14].replace(",", ""))
    tot[name][0] += 1
    tot[name][1] += ns
total = sum(v[1] for v in tot.values())
print(f"# ncu launch list summary: `{cmd}`\n")
print(f"{len(rows)} launches, {total / 1e6:.1f} ms of kernel time (cold-cache, serialised: compare shares).  "
      f"Raw list: {path.split('/')[-1]}\n")
print("| kernel | launches | total us | share |\n|---|---|---|---|")
for k, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {n} | {ns / 1e3:.0f} | {100 * ns / total:.1f}% |")
fam = collections.defaultdict(float)
for k, (n, ns) in tot.items():
    for f in ("enhance_kernel", "stoi_stream", "align_kernel"):
        if k.startswith(f):
            fam[f] += ns
print("\nFamilies: " + ", ".join(f"{k.replace('_kernel', '')} {100 * v / total:.1f}%" for k, v in sorted(fam.items(), key=lambda kv: -kv[1])))
if len(sys.argv) > 3:
    d = json.loads(open(sys.argv[3]).read().strip().splitlines()[-1])
    f = d["roofline"]["families"]
    print(f"\nLive CUDA-event shares of the timed steps of the full bench (`{sys.argv[3].split('/')[-1]}`): "
          + ", ".join(f"{k} {100 * v['share_of_step']:.1f}%" for k, v in f.items()))
