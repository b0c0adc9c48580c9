"""Pretty-print a bench.py JSON line: headline, roofline, bins.   python scripts/show_bench.py file.json"""
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.strip().startswith("{")][-1])
r = d.pop("roofline", None) or {}
bins = r.pop("bins", [])
for k, v in d.items():
    print(f"{k}: {json.dumps(v)[:400]}")
print("roofline:", json.dumps(r)[:900])
for b in bins:
    print(f"   side {b['side']} {b['team']:>11} cap {b['cap']:4d} rows {b['rows']:7d} nnz {b['nnz']:9d} ms {b['ms_per_sweep']:.4f} GB/s {b.get('GBps', 0):8.1f}")
