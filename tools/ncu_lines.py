"""Developer tool: per-source-line instruction and stall-sample totals from an ncu report captured with
--import-source on (kernels compiled with -lineinfo).
   ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python tools/ncu_lines.py src.csv [top]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr = None, None
inst = defaultdict(int)
samp = defaultdict(int)
text = {}
line = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_inst, i_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or r[0] in ("Function Name", "File Name"):
        continue
    if r[0] != "":
        line = (cur_file, int(r[0]))
        text[line] = r[1]
    elif line is not None and len(r) > i_inst and r[2] not in ("", "..."):
        try:
            inst[line] += int(r[i_inst])
            samp[line] += int(r[i_samp])
        except ValueError:
            pass
ti, ts = sum(inst.values()), sum(samp.values())
print(f"total warp instructions {ti}, samples {ts}")
for k in sorted(inst, key=lambda k: -inst[k])[:top]:
    print(f"{k[0]:>16s}:{k[1]:<4d} inst {inst[k]:>10d} ({100*inst[k]/ti:4.1f}%)  samples {samp[k]:>6d} ({100*samp[k]/max(ts,1):4.1f}%)  {text[k].strip()[:90]}")
