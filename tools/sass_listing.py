"""SASS listings of the hot kernels with an opcode histogram -> profiles/<prefix>_sass_<kernel>.txt
Usage: python tools/sass_listing.py [prefix]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
prefix = sys.argv[1] if len(sys.argv) > 1 else "r02"
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "classical_speech_enhancement_b200", "libcse_sm100a.so")
raw = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
want = {"_Z14enhance_kernelILi3ELi10ELb1ELb1ELb1EEv11EnhanceArgs": "enhance_kernel_3_10_staged_gamma",
        "_Z18select_best_kernelPK11cse_score_tPKdiP12cse_winner_t": "select_best_kernel",
        "_Z18stoi_stream_kernel9ScoreArgs": "stoi_stream_kernel",
        "_Z12align_kernelILb0EEv9ScoreArgs": "align_kernel_0"}
cur, bodies = None, collections.defaultdict(list)
for line in raw.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur in want:
        bodies[cur].append((m.group(1), m.group(2).strip()))
for sym, short in want.items():
    ins = bodies[sym]
    hist = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in ins)
    out = os.path.join(ROOT, "profiles", f"{prefix}_sass_{short}.txt")
    with open(out, "w") as f:
        f.write(f"# SASS of {sym} (sm_100a, nvcc 12.9, -O3 -lineinfo), {len(ins)} instructions\n")
        f.write("# opcode histogram: " + ", ".join(f"{k} {v}" for k, v in hist.most_common()) + "\n")
        for addr, text in ins:
            f.write(f"{addr}  {text}\n")
    print(out, len(ins))
