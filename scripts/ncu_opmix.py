"""Dynamic SASS opcode mix of one launch of an ncu report.  usage: python scripts/ncu_opmix.py report.ncu-rep [launch] [unit]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; li = int(sys.argv[2]) if len(sys.argv) > 2 else 0
unit = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
secs = []; cur = None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Address": cur = {"hdr": r, "rows": []}; secs.append(cur)
    elif cur is not None and len(r) == len(cur["hdr"]): cur["rows"].append(r)
s = secs[li]; h = s["hdr"]; ii = h.index("Instructions Executed"); isrc = h.index("Source")
mix = collections.Counter(); tot = 0
for r in s["rows"]:
    c = int(r[ii] or 0); ins = r[isrc].split()
    op = ins[1] if ins and ins[0].startswith("@") else (ins[0] if ins else "?")
    mix[op.split(".")[0]] += c; tot += c
print(f"launch {li}: {tot} warp-instructions ({tot / unit:.1f} per unit)")
for op, c in mix.most_common(28): print(f"  {op:10s} {c:12d}  {100 * c / tot:5.1f}%  {c / unit:8.1f}")
