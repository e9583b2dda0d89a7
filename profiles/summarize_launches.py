"""Summarise an ncu launch list (gpu__time_duration per launch) for one steady-state iteration.
usage: summarize_launches.py launches.csv [--detail]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
names = [r["Kernel Name"] for r in rows]
idx = [i for i, n in enumerate(names) if "commit_kernel" in n]
seg = rows[idx[-2] + 1: idx[-1] + 1]
tot, cnt = collections.defaultdict(float), collections.Counter()
for r in seg:
    n = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("bre::<unnamed>::", "")
    n = re.sub(r"\(int\)", "", n)
    t = float(r["Metric Value"]) / 1000.0
    tot[n] += t
    cnt[n] += 1
T = sum(tot.values())
print(f"# {path}: one iteration = {len(seg)} launches, sum of kernel durations {T:.1f} us (cold-cache, serialised)")
for n, t in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"{t:9.1f} us {100 * t / T:5.1f}%  x{cnt[n]:3d}  {n}")
if "--detail" in sys.argv:
    for i, r in enumerate(seg):
        if "igemm" in r["Kernel Name"] or "dgrad_small" in r["Kernel Name"]:
            n = re.sub(r"\(bre.*", "", r["Kernel Name"]).replace("void bre::<unnamed>::", "")
            print(i, n, r["Grid Size"], float(r["Metric Value"]) / 1000.0, "us")
