"""Phase / chain summary of a step timeline written by tools/timeline.py: python tools/timeline_phases.py gpurun_out/timeline_X.json.gz"""
import collections
import gzip
import json
import re
import sys

ev = json.load(gzip.open(sys.argv[1]))["traceEvents"]
ev.sort(key=lambda e: e["ts"])
starts = [e["ts"] for e in ev if "im2col" in e["name"]]
step = [e for e in ev if starts[1] <= e["ts"] < starts[2]] if len(starts) >= 3 else ev
T0 = step[0]["ts"]
end = max(e["ts"] + e["dur"] for e in step)
print(f"step span {(end - T0) / 1e3:.3f} ms, {len(step)} kernels")


def short(n):
    n = re.sub(r"^void\s+", "", n)
    n = re.sub(r"\(anonymous namespace\)::|gemmk::", "", n).split("(")[0]
    return n.replace("gemm_bf16_tcgen05_pair", "GEMMp").replace("gemm_bf16_tcgen05", "GEMM")[:40]


by = collections.defaultdict(list)
for e in step:
    by[e["tid"]].append(e)
for s, es in sorted(by.items(), key=lambda kv: -len(kv[1])):
    busy = sum(e["dur"] for e in es)
    fam = collections.Counter()
    for e in es:
        fam[short(e["name"]).split("<")[0]] += e["dur"]
    gaps = sum(max(0.0, es[i + 1]["ts"] - es[i]["ts"] - es[i]["dur"]) for i in range(len(es) - 1))
    print(f"stream {s}: {len(es)} kernels, busy {busy / 1e3:.3f} ms, gaps {gaps / 1e3:.3f} ms, first +{(es[0]['ts'] - T0) / 1e3:.3f}, "
          f"last end +{(es[-1]['ts'] + es[-1]['dur'] - T0) / 1e3:.3f}")
    print("     " + ", ".join(f"{k} {v / 1e3:.2f}" for k, v in fam.most_common(8)))
hl = [e for e in step if "ls_ce" in e["name"]]
if hl:
    print(f"forward ends / loss at +{(hl[0]['ts'] - T0) / 1e3:.3f} ms")
for e in step:
    if "adamw" in e["name"]:
        print(f"  adamw +{(e['ts'] - T0) / 1e3:.3f} dur {e['dur']:.0f} us grid {e['args']['grid'][0]}")
