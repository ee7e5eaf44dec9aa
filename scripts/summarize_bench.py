import json, sys
for l in (open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin):
    if not l.startswith("{"):
        print(l.strip()[:200]); continue
    d = json.loads(l)
    c = d["config"]
    print(f'{c["workload"]:40s} G={c["lanes_per_tree"]:<2d} roots={c["roots_per_gpu"]:<6d} {d["ms_per_step"]:9.1f} ms  {d["value"]/1e6:8.1f} Msim/s  e2e {d["e2e"]["value"]/1e6:8.1f}  '
          f'rollsteps {d["rollout_steps_per_s"]/1e9:5.2f} G/s  frac {d["roofline"]["frac"]:.4f}  clk {d["clocks"]["sm_mhz"]}')
