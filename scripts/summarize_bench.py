#!/usr/bin/env python
"""Markdown rows from bench lines: python scripts/summarize_bench.py profiles/r02/*.json"""
import json, sys, os
print("| file | N | workload | batch | queries/s | ms/step | e2e q/s | kernel roofline | whole step | rescans | verify (q / mism / ties) |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
    except Exception:
        continue
    if d.get("impl") == "reference" or "roofline" not in d:
        continue
    name = os.path.basename(f)
    r = d["roofline"]
    def row(batch, qps, ms, e2e, frac, unit, whole, resc, v):
        vv = f"{v['queries']} / {v['mismatches']} / {v['ties']}" if v else "-"
        print(f"| `{name}` | {d['n_gpus']} | {d['config']['workload']} | {batch} | {qps:,.0f} | {ms:.3f} | {e2e} | {frac:.3f} {unit} | {whole:.3f} | {resc} | {vv} |")
    row(d["config"]["batch"], d["value"], d["ms_per_step"], f"{d['e2e']['value']:,.0f}", r["frac"], "of " + ("HBM" if r["bound"] == "hbm" else "bf16"),
        d.get("whole_step_frac", 0), d.get("exact_rescans"), d.get("verify"))
    for s in d.get("sweep", []):
        if s["batch"] == d["config"]["batch"]:
            continue
        row(s["batch"], s["qps"], s["p50_ms"], "", s["frac"], "of " + ("HBM" if s["bound"] == "hbm" else "bf16"), s.get("whole_step_frac", 0), s.get("exact_rescans"), s.get("verify"))
    c = d.get("clustered")
    if c:
        v = c["verify"]
        print(f"| `{name}` | {d['n_gpus']} | {d['config']['workload']} **clustered** | {c['batch']} | {c['qps']:,.0f} | {c['ms_per_step']:.3f} | | | | {c['exact_rescans_per_step']:.0f} / step | {v['queries']} / {v['mismatches']} / {v['ties']} |")
