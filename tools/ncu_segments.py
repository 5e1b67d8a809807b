"""Summarises an `ncu --page source --csv` export: stall samples per SASS segment (split at BAR.SYNC / EXIT), with
the opcode mix of each segment -- enough to tell the phases of a tile kernel apart without the GUI.
usage: ncu_segments.py file.csv [min_share]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ci = {n: i for i, n in enumerate(hdr)}
S, IEX = ci["# Samples"], ci["Instructions Executed"]
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
segs, cur = [], {"n": 0, "samples": 0, "inst": 0, "ops": collections.Counter(), "st": collections.Counter(), "first": None}
total = 0
for r in rows[hdr_i + 1:]:
    if len(r) <= S: continue
    src = r[ci["Source"]].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    smp = int(r[S] or 0); total += smp
    cur["n"] += 1; cur["samples"] += smp; cur["inst"] += int(r[IEX] or 0)
    cur["ops"][op.split(".")[0]] += int(r[IEX] or 0)
    for n in stalls:
        v = int(r[ci[n]] or 0)
        if v: cur["st"][n[6:]] += v
    if cur["first"] is None: cur["first"] = r[0]
    if op.startswith("BAR") or op.startswith("EXIT"):
        segs.append(cur)
        cur = {"n": 0, "samples": 0, "inst": 0, "ops": collections.Counter(), "st": collections.Counter(), "first": None}
if cur["n"]: segs.append(cur)
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
print(f"total samples {total}, {len(segs)} segments")
for i, s in enumerate(segs):
    if s["samples"] < min_share * total: continue
    ops = ", ".join(f"{k}:{v}" for k, v in s["ops"].most_common(6))
    st = ", ".join(f"{k}:{v}" for k, v in s["st"].most_common(4))
    print(f"seg {i:3d} @{s['first'][-5:]} sass={s['n']:5d} warp-inst={s['inst']:9d} samples={s['samples']:6d} ({100*s['samples']/total:5.1f}%)  [{ops}]  stalls[{st}]")
