import json, sys
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"): continue
    d = json.loads(line)
    r = d.get("roofline") or {}
    print(f"ms/step {d['ms_per_step']:.3f} (serialized {r.get('serialized_ms_per_step', 0):.3f}) value {d['value']/1e6:.1f} Mnnz/s  sweep_frac {r.get('sweep_frac_of_peak',0):.4f} e2e_ms {(d.get('e2e') or {}).get('ms_per_step')} launches {d['gpu_launches']}")
    if len(sys.argv) > 1:
        for b in r.get("bins", []): print("   ", b)
