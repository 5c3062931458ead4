#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: top instructions by executed count and by stall samples."""
import csv, sys
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot_inst = sum(f(r, "Instructions Executed") for r in body)
tot_samp = sum(f(r, "# Samples") for r in body)
print(f"total warp-instructions {tot_inst:.0f}, samples {tot_samp:.0f}, sass lines {len(body)}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(f(r, s) for r in body) for s in stalls}
print("stall totals:", ", ".join(f"{k[6:]}={v:.0f}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0))
# opcode histogram
from collections import Counter
op = Counter(); ops = Counter()
for r in body:
    s = r[ix["Source"]].strip()
    toks = s.split()
    if toks and toks[0].startswith("@"): toks = toks[1:]
    o = toks[0].split(".")[0] if toks else "?"
    op[o] += f(r, "Instructions Executed"); ops[o] += f(r, "# Samples")
print("opcode executed share:", ", ".join(f"{k}={v/tot_inst*100:.1f}%" for k, v in op.most_common(25)))
print("opcode sample share:", ", ".join(f"{k}={v/tot_samp*100:.1f}%" for k, v in ops.most_common(15)))
print("--- top by samples")
for r in sorted(body, key=lambda r: -f(r, "# Samples"))[:topn]:
    st = sorted(((f(r, s), s[6:]) for s in stalls), reverse=True)[:3]
    print(f"{f(r,'# Samples'):8.0f} {f(r,'Instructions Executed'):12.0f}  {r[ix['Source']].strip()[:70]:70s} " + " ".join(f"{n}:{v:.0f}" for v, n in st if v > 0))
