"""Developer tool: stall-reason breakdown (sampled) per source line, optionally restricted to files / lines.
   ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python tools/ncu_stalls.py src.csv [top] [file-substring]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
only = sys.argv[3] if len(sys.argv) > 3 else None
cur = hdr = line = None
samp, inst, text = defaultdict(int), defaultdict(int), {}
stall = defaultdict(lambda: defaultdict(int))
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ii, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        st = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or r[0] in ("Function Name", "File Name"):
        continue
    if r[0] != "":
        line = (cur, int(r[0]))
        text[line] = r[1]
    elif line is not None and len(r) > ii and r[2] not in ("", "..."):
        try:
            inst[line] += int(r[ii])
            samp[line] += int(r[isamp])
            for i, h in st:
                if i < len(r) and r[i].isdigit():
                    stall[line][h] += int(r[i])
        except ValueError:
            pass
ts = sum(samp.values())
keys = [k for k in samp if only is None or only in k[0]]
tot = defaultdict(int)
for k in keys:
    for h, v in stall[k].items():
        tot[h] += v
sel = sum(samp[k] for k in keys)
print(f"samples: {sel} of {ts} in selection; by reason: " + " ".join(f"{h[6:]}={v}" for h, v in sorted(tot.items(), key=lambda x: -x[1])[:8]))
for k in sorted(keys, key=lambda k: -samp[k])[:top]:
    t3 = sorted(stall[k].items(), key=lambda x: -x[1])[:3]
    print(f"{k[0]:>14s}:{k[1]:<4d} samp {samp[k]:5d} ({100*samp[k]/ts:4.1f}%) inst {inst[k]:8d}  "
          f"{' '.join(f'{h[6:]}={v}' for h, v in t3):45s} {text[k].strip()[:70]}")
