"""Summarise a phase-timing log (scripts/gpu_phase_timing.sh): per-warp cycles per iteration, slowest warps, per-SM load."""
import sys
import numpy as np
rows = []
for l in open(sys.argv[1]):
    if not l.startswith("PT "):
        if "ms_per_step" in l:
            import re
            print("ms_per_step", re.search(r'"ms_per_step": ([0-9.]+)', l).group(1))
        continue
    if "level:" in l:
        print(l.strip()[l.index("level:"):]) if len(rows) < 2 else None
        l = l[:l.index("level:")]
    p = l.replace("|", " ").split()
    rows.append([float(x) for x in p[1:]])
a = np.array(rows)  # block warp smid lv st cd cr cb pd pr pb
tot = a[:, 5] + a[:, 6] + a[:, 7]
print("warps", len(a), " total cycles/iter: mean %.0f  p50 %.0f  p90 %.0f  max %.0f  (x1000 iter / 1.965 GHz = %.1f ms max)" % (tot.mean(), np.median(tot), np.percentile(tot, 90), tot.max(), tot.max() / 1965.0))
print("levels/iter mean %.2f  steps/iter mean %.2f" % (a[:, 3].mean(), a[:, 4].mean()))
print("per level/step/level cycles (mean over warps): descent %.0f rollout %.0f backup %.0f" % (a[:, 8].mean(), a[a[:, 4] > 1][:, 9].mean(), a[:, 10].mean()))
o = np.argsort(-tot)[:8]
for i in o:
    print("  slow warp: block %d warp %d sm %d  levels %.2f steps %.2f  cycles D %d R %d B %d  per-step %d %d %d" % tuple(a[i][[0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10]]))
# per block: max warp total and mean
blocks = {}
for r, t in zip(a, tot):
    blocks.setdefault(int(r[0]), []).append(t)
bm = np.array([max(v) for v in blocks.values()])
print("per-block slowest warp: min %.0f mean %.0f max %.0f" % (bm.min(), bm.mean(), bm.max()))
