"""Top stall locations of an ncu report's source page: ncu -i X.ncu-rep --page source --csv | python tools/ncu_hot.py [N]"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        print("==", rows[i][1][:100])
        hdr = rows[i + 1]
        ci = {h: k for k, h in enumerate(hdr)}
        j = i + 2
        data = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            r = rows[j]
            try:
                data.append((int(r[ci["Warp Stall Sampling (All Samples)"]]), j - i - 2, r))
            except Exception:
                pass
            j += 1
        tot = sum(d[0] for d in data) or 1
        reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        for cnt, idx, r in sorted(data, key=lambda x: -x[0])[:n]:
            top = sorted(((int(r[ci[h]] or 0), h[6:]) for h in reasons), reverse=True)[:2]
            print(f"{100 * cnt / tot:5.1f}%  #{idx:4d} {r[ci['Source']][:90]:90s} {top}")
        i = j
    else:
        i += 1
