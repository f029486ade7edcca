"""Warp-stall samples of an ncu source page by opcode and by reason, barrier waits apart:
    ncu -i X.ncu-rep --page source --csv > src.csv; python tools/ncu_stalls.py src.csv <kernel name substring>"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2]
i = 0; seen = set()
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]
        hdr = rows[i+1]; ci = {h:k for k,h in enumerate(hdr)}
        j = i+2; data=[]
        while j < len(rows) and not (rows[j] and rows[j][0]=="Kernel Name"):
            data.append(rows[j]); j+=1
        key = name[:90]
        if want in name and key not in seen:
            seen.add(key)
            reasons=[h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
            tot=collections.Counter(); byop=collections.Counter(); total=0
            for r in data:
                try: c=int(r[ci["Warp Stall Sampling (All Samples)"]])
                except: continue
                src=r[ci["Source"]]
                if "TRYWAIT" in src or src.strip().startswith("@P0   BRA") or "UCGABAR" in src or "BAR.SYNC" in src: 
                    tot["(barrier waits)"]+=c; total+=c; continue
                total+=c
                op=src.replace("@P0","").replace("@P1","").replace("@!P0","").split()[0] if src.split() else "?"
                byop[op]+=c
                for h in reasons:
                    v=int(r[ci[h]] or 0)
                    if v: tot[h[6:]]+=v
            print("==",key, "total samples", total)
            print("  barrier waits:", f"{100*tot['(barrier waits)']/total:.1f}%")
            print("  by opcode:", [(o, f"{100*c/total:.1f}%") for o,c in byop.most_common(14)])
            print("  by reason:", [(o, f"{100*c/total:.1f}%") for o,c in tot.most_common(10) if o!='(barrier waits)'])
        i=j
    else: i+=1
