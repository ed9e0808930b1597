#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: per basic-block-ish runs of SASS (same execution count) with
instructions executed, average active threads and stall samples.  python tools/ncu_src.py file.csv [min_Minst]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0  # which kernel section of the export
hdrs = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hdr_i = hdrs[which]
end = hdrs[which + 1] - 1 if which + 1 < len(hdrs) else len(rows)
hdr = rows[hdr_i]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hdr_i + 1:end] if len(r) == len(hdr)]
print(f"sections {len(hdrs)}; using {which}")
tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in data)
tot_samp = sum(int(r[ix["# Samples"]]) for r in data)
print(f"total warp-inst {tot_inst/1e6:.1f} M, samples {tot_samp}")
thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
# runs of equal execution count
runs = []
cur = None
for k, r in enumerate(data):
    n = int(r[ix["Instructions Executed"]])
    if cur is None or n != cur["n"]:
        cur = {"n": n, "start": k, "len": 0, "samp": 0, "thr": 0, "ops": {}, "stalls": {}}
        runs.append(cur)
    cur["len"] += 1
    cur["samp"] += int(r[ix["# Samples"]])
    cur["thr"] += int(r[ix["Thread Instructions Executed"]])
    op = r[ix["Source"]].split()[0] if r[ix["Source"]].split() else "?"
    if op.startswith("@"):
        op = r[ix["Source"]].split()[1]
    op = op.split(".")[0]
    cur["ops"][op] = cur["ops"].get(op, 0) + 1
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            v = int(r[ix[h]] or 0)
            if v:
                cur["stalls"][h[6:]] = cur["stalls"].get(h[6:], 0) + v
for c in runs:
    winst = c["n"] * c["len"]
    if winst / 1e6 < thresh:
        continue
    ops = sorted(c["ops"].items(), key=lambda x: -x[1])[:6]
    st = sorted(c["stalls"].items(), key=lambda x: -x[1])[:4]
    print(f"@{c['start']:5d} len {c['len']:4d} exec {c['n']/1e6:7.2f}M  winst {winst/1e6:7.1f}M ({100*winst/tot_inst:4.1f}%)  thr/inst {c['thr']/max(1,winst):5.1f}  samples {100*c['samp']/tot_samp:4.1f}%  ops {ops} stalls {st}")
