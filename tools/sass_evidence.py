"""SASS mnemonics of the built library, per kernel family: the evidence that the GEMMs run on tcgen05 (UTCHMMA) fed by TMA
(UTMALDG) with TMEM reads (LDTM) and mbarriers (SYNCS), that attention / fusion use the warp-level tensor path (HMMA) and cp.async
(LDGSTS).  CPU only:  python tools/sass_evidence.py > profiles/sass_evidence_r1.txt"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mkg_analogy_b200", "libb200mart.so")
WANT = ("UTCHMMA", "UTCBAR", "UTMALDG", "UTMAPF", "LDTM", "SYNCS", "HMMA", "LDGSTS", "LDSM", "RED", "REDG", "ATOMG", "ACQBULK", "UCGABAR", "CCTL")


def family(mangled):
    out = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
    out = re.sub(r"\(anonymous namespace\)::|gemmk::", "", out)
    m = re.match(r"(?:void )?([A-Za-z0-9_]+)(<[^(]*>)?", out)
    name = m.group(1) if m else out
    if name.startswith("gemm_bf16_tcgen05"):
        return name            # all (tile, layout, epilogue) instantiations together
    return name


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.defaultdict(collections.Counter)
    ninst = collections.Counter()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = family(m.group(1))
            ninst[cur] += 1
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            if op in WANT:
                counts[cur][op] += 1
    print(f"cuobjdump -sass {os.path.relpath(LIB)} (sm_100a): occurrences of selected mnemonics per kernel family (all instantiations summed)")
    for fam in sorted(counts, key=lambda f: -sum(counts[f].values())):
        c = counts[fam]
        print(f"{fam:34s} x{ninst[fam]:<3d} " + "  ".join(f"{k}={c[k]}" for k in WANT if c[k]))


if __name__ == "__main__":
    sys.exit(main())
