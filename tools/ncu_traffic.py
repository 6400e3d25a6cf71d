"""Extract DRAM traffic per item for each profiled kernel from an ncu report into profiles/ncu_traffic.json.
Usage: python tools/ncu_traffic.py report.ncu-rep items_per_launch [source-note]"""
import csv
import json
import os
import subprocess
import sys

rep, items = sys.argv[1], int(sys.argv[2])
note = sys.argv[3] if len(sys.argv) > 3 else os.path.basename(rep)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}


def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
tab = json.load(open(out_path)) if os.path.exists(out_path) else {}
score = 0.0
for d in data:
    name = d[idx["Kernel Name"]].split("(")[0].replace("void ", "").strip()
    rd = to_bytes(d[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
    wr = to_bytes(d[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
    per_item = (rd + wr) / items
    tab[name] = {"dram_bytes_per_item": per_item, "source": note}
json.dump(tab, open(out_path, "w"), indent=1)
print(json.dumps(tab, indent=1))
