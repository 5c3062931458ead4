#!/usr/bin/env python
"""DESIGN.md = tools/DESIGN.md.in with the @@NAME@@ fields filled from the committed bench records profiles/r02_bench*_<tag>.json.
usage: python tools/fill_design.py <tag>      (e.g. v33)"""
import json, os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
def rec(name):
    path = os.path.join(root, "profiles", f"r02_{name}_{tag}.json")
    return json.loads(open(path).read().strip().splitlines()[-1])
d = rec("bench")
km, rf = d["kernel_ms"], d["roofline"]
fwdsum = km["cnn_dirty"] + km["cnn_inc_scan"] + km["cnn_forward_inc_tc"] + km["cnn_inc_merge"]
def other(name):
    try:
        r = rec(name)
        e = r.get("e2e") or {}
        s = f"{r['value'] / 1e6:.2f} M | {r['ms_per_step']:.3g} | " + (f"{e['value'] / 1e6:.2f} M" if e.get("value") else "")
        return s
    except Exception as ex:
        return f"(no record: {ex})"
cpu = d.get("cpu_baseline") or {}
f = {
    "TAG": tag, "MS": f"{d['ms_per_step']:.1f}", "VAL": f"{d['value'] / 1e6:.2f}",
    "K_PROP": f"{km['pas_propose'] + km['potts_incremental']:.2f}", "K_FWD": f"{km['cnn_forward_inc_tc']:.2f}",
    "K_MERGE": f"{km['cnn_inc_merge']:.2f}", "K_REC": f"{km['cnn_winner_sort']:.2f}", "K_BWD": f"{km['cnn_backward_tc']:.2f}",
    "K_REV": f"{km['pas_reverse_accept'] + km['cnn_grad_combine']:.2f}", "ROOF_ACH": f"{rf['achieved']:.0f}", "ROOF_FRAC": f"{rf['frac']:.2f}",
    "HBM_FRAC": f"{rf['hbm_kernels']['frac']:.2f}", "HBM_DFRAC": f"{rf['hbm_kernels'].get('frac_of_design_bytes', float('nan')):.2f}",
    "K_FWDSUM": f"{fwdsum:.1f}", "FWD_EQ": f"{rf['forward_incremental']['full_evaluation_equivalent_tflops'] / 1e3:.2f}",
    "E2E": f"{d['e2e']['value'] / 1e6:.2f}",
    "CPU": f"{cpu.get('value', float('nan')):.0f} chain-steps/s ({cpu.get('kind')}, {cpu.get('cores')} threads)",
    "UBE": other("bench_ube4b"), "PAS10": other("bench_pas10"), "PABP": other("bench_pabp128"),
}
s = open(os.path.join(root, "tools", "DESIGN.md.in")).read()
for k, v in f.items():
    s = s.replace(f"@@{k}@@", v)
left = [w for w in s.split() if w.startswith("@@")]
assert not left, left
open(os.path.join(root, "DESIGN.md"), "w").write(s)
print("DESIGN.md written from", tag, {k: f[k] for k in ("MS", "VAL", "E2E")})
