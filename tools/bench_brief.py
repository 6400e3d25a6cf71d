"""One-line digest of a bench.py JSON line: python tools/bench_brief.py bench.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
f = d["roofline"]["families"]
tot = sum(v["ms"] for v in f.values())
print(f"{d['n_gpus']} GPU: value {d['value']:.0f} {d['unit']}, {d['ms_per_step']:.1f} ms/step, hot kernels "
      f"{tot / d['steps']:.1f} ms/step on rank 0, e2e {d['e2e']['value']:.0f}, whole-path roofline "
      f"{d['roofline']['whole_path']['frac']:.3f}, dominant {d['roofline']['kernel']} frac {d['roofline']['frac']:.3f}, "
      f"clocks {d['clocks']}")
