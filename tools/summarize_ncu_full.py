"""Per-kernel averages of an `ncu --set full` report: ncu -i X.ncu-rep --page raw --csv | python tools/summarize_ncu_full.py
Columns: launches, duration, DRAM read / write bytes, L2 sector traffic, tensor-pipe activity."""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(sys.stdin))
hdr, units, data = rows[0], rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread"]
want = [w for w in want if w in ci]
print("units:", {w: units[ci[w]] for w in want})
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
agg = defaultdict(lambda: [0, defaultdict(float)])
for r in data:
    name = re.sub(r"\(.*", "", r[ci["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "").replace("gemmk::", "")
    key = f"{name} grid={r[ci['Grid Size']]}"
    a = agg[key]
    a[0] += 1
    for w in want:
        try:
            a[1][w] += float(r[ci[w]].replace(",", "")) * SCALE.get(units[ci[w]], 1.0)
        except ValueError:
            pass
tot_rd = tot_wr = tot_n = 0
for key, (n, m) in sorted(agg.items(), key=lambda kv: -kv[1][1]["gpu__time_duration.sum"]):
    rd, wr = m["dram__bytes_read.sum"] / n, m["dram__bytes_write.sum"] / n
    if "gemm" in key:
        tot_rd += m["dram__bytes_read.sum"]; tot_wr += m["dram__bytes_write.sum"]; tot_n += n
    print(f"{key:70s} x{n:<3d} {m['gpu__time_duration.sum'] / n:7.1f} us  dram rd {rd / 1e6:7.2f} wr {wr / 1e6:7.2f} MB  "
          f"L2 {m['lts__t_sectors.sum'] / n * 32 / 1e6:7.1f} MB ({m['lts__throughput.avg.pct_of_peak_sustained_elapsed'] / n:4.1f}% L2 peak)  "
          f"tensor pipe {m['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'] / n:4.1f}%  regs {m['launch__registers_per_thread'] / n:.0f}")
if tot_n:
    print(f"GEMM launches: {tot_n}, mean DRAM traffic per launch {(tot_rd + tot_wr) / tot_n:.0f} bytes (read {tot_rd / tot_n:.0f} + write {tot_wr / tot_n:.0f})")
