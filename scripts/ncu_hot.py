"""Dev tool: hottest SASS instructions (by stall samples) of an `ncu --page source --csv` dump, with stall-reason totals.

    ncu -i X.ncu-rep --page source --csv > x.csv ; python scripts/ncu_hot.py x.csv [top]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
body = []
for r in rows[hdr_i + 1:]:            # first kernel instance only
    if r and r[0] in ("Kernel Name", "Address"):
        break
    if len(r) == len(hdr):
        body.append(r)
ci = {h: i for i, h in enumerate(hdr)}
S = ci["# Samples"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[S] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
agg = {h: sum(int(r[ci[h]] or 0) for r in body) for h in stalls}
print("stall totals:", ", ".join(f"{h[6:]}={v} ({100 * v / max(tot, 1):.1f}%)" for h, v in sorted(agg.items(), key=lambda x: -x[1]) if v))
exc = sum(int(r[ci["L1 Wavefronts Shared Excessive"]] or 0) for r in body)
print("excess shared wavefronts", exc)
idx = sorted(range(len(body)), key=lambda i: -int(body[i][S] or 0))[:top]
for i in sorted(idx):
    r = body[i]
    st = sorted(((int(r[ci[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
    print(f"{i:6d} {int(r[S]):7d} {100 * int(r[S]) / tot:5.1f}%  exe={r[ci['Instructions Executed']]:>9}  xs={r[ci['L1 Wavefronts Shared Excessive']]:>8}  {r[ci['Source']][:90]:90s} {st}")
