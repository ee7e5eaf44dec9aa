"""One line per record of a bench.py JSON line: the headline, the other workloads, the strong-scaling records.
usage: python scripts/summarize_bench.py <file with the JSON line>   (or on stdin)"""
import json
import sys


def show(name, r):
    ms = r.get("ms_per_step", r.get("ms_per_decision"))
    frac = r.get("roofline", {}).get("frac")
    cpu = r.get("cpu_baseline", {}).get("value")
    print(f'{name:44s} {ms:9.2f} ms  {r["value"] / 1e6:9.1f} Msim/s' + (f'  frac {frac:.4f}' if frac is not None else '') +
          (f'  lanes {r["lanes_per_tree"]}' if "lanes_per_tree" in r else '') + (f'  cpu {cpu / 1e6:.2f} Msim/s' if cpu else ''))


for line in (open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin):
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    print(f'N = {d["n_gpus"]}, clocks {d.get("clocks", {}).get("sm_mhz")} MHz, e2e {d["e2e"]["value"] / 1e6:.1f} Msim/s')
    show(d["config"]["workload"] + " (headline)", d)
    for k, v in d.get("workloads", {}).items():
        show(k, v)
    for k, v in d.get("strong", {}).items():
        show(k + " [strong]", v)
