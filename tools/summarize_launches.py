"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: time by kernel name (+ grid for GEMMs)."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as fh:
    lines = [l for l in fh if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    rows.append((name, r["Grid Size"], r["Block Size"], us))
tot = sum(r[3] for r in rows)
by = defaultdict(lambda: [0, 0.0])
for n, g, b, us in rows:
    key = n if "gemm" not in n else f"{n} grid={g}"
    by[key][0] += 1
    by[key][1] += us
print(f"{len(rows)} launches, {tot / 1000:.3f} ms total")
detail = "--detail" in sys.argv
agg = defaultdict(lambda: [0, 0.0])
for k, (c, us) in by.items():
    fam = re.sub(r"<unnamed>::|gemmk::", "", k.split(" grid=")[0])
    fam = re.sub(r"<.*", "", fam).split("::")[-1]
    agg[fam][0] += c
    agg[fam][1] += us
print("-- by kernel family")
for k, (c, us) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{us / 1000:9.3f} ms {100 * us / tot:5.1f}%  x{c:<4d} {k}")
if detail:
    print("-- by kernel (GEMMs by grid)")
    for k, (c, us) in sorted(by.items(), key=lambda x: -x[1][1])[:60]:
        print(f"{us / 1000:9.3f} ms {100 * us / tot:5.1f}%  x{c:<4d} avg {us / c:8.1f} us  {k}")
