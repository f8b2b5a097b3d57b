"""Aggregate an `ncu --page source --csv` dump by instruction-address region (helper for profiles/)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
def num(x):
    try: return float(x)
    except: return 0.0
recs = []
for r in data:
    if len(r) < len(hdr): continue
    recs.append(dict(addr=int(r[col["Address"]], 16) if r[col["Address"]].startswith("0x") else int(r[col["Address"]]),
                     src=r[col["Source"]], samples=num(r[col["# Samples"]]), inst=num(r[col["Instructions Executed"]]),
                     tinst=num(r[col["Thread Instructions Executed"]]),
                     stalls={h: num(r[col[h]]) for h in hdr if h.startswith("stall_") and "Not Issued" not in h}))
tot_s = sum(x["samples"] for x in recs); tot_i = sum(x["inst"] for x in recs)
print("total samples", tot_s, "total inst", tot_i)
# find scan loop: first LDS.128 to the loop back-branch
base = recs[0]["addr"]
bounds = [int(b, 16) for b in sys.argv[2:]]  # relative offsets delimiting regions
bounds = [0] + bounds + [10**9]
for lo, hi in zip(bounds[:-1], bounds[1:]):
    sel = [x for x in recs if lo <= x["addr"] - base < hi]
    if not sel: continue
    s = sum(x["samples"] for x in sel); i = sum(x["inst"] for x in sel); ti = sum(x["tinst"] for x in sel)
    st = {}
    for x in sel:
        for k, v in x["stalls"].items(): st[k] = st.get(k, 0) + v
    top = sorted(st.items(), key=lambda kv: -kv[1])[:5]
    print(f"[{lo:#07x},{hi:#07x}) n={len(sel):5d} samples {100*s/tot_s:5.1f}%  inst {100*i/tot_i:5.1f}%  thr/inst {ti/max(i,1):5.1f}  " +
          " ".join(f"{k[6:]}={100*v/max(s,1):.0f}%" for k, v in top))
