"""Summarise an ncu report: headline metrics, stall reasons, and executed instructions per source line.
usage: python scripts/ncu_lines.py report.ncu-rep [launch_index] [top_n]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; li = int(sys.argv[2]) if len(sys.argv) > 2 else 0; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed']
stalls = [h for h in hdr if 'smsp__average_warps_issue_stalled' in h and 'per_issue_active' in h]
r = rows[2 + li]
for w in want:
    if w in idx: print(f"  {w} = {r[idx[w]]} {rows[1][idx[w]]}")
st = sorted([(float(r[idx[h]] or 0), h) for h in stalls], reverse=True)[:8]
print("  stalls/issue:", ", ".join(f"{h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')} {v:.2f}" for v, h in st))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
secs = []; cur = None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "File Path": cur = {'file': r[1], 'rows': []}; secs.append(cur)
    elif r and r[0] == "Function Name": cur['func'] = r[1]
    elif r and r[0] == "Line No": cur['hdr'] = r
    elif cur is not None and 'hdr' in cur and r: cur['rows'].append(r)
funcs = []
for s in secs:
    if not funcs or (s['func'] != funcs[-1][0]['func']) or any(s['file'] == t['file'] for t in funcs[-1]): funcs.append([s])
    else: funcs[-1].append(s)
group = funcs[li] if li < len(funcs) else funcs[0]
agg = collections.Counter(); stall = collections.Counter(); text = {}; tot = 0; stot = 0
for s in group:
    h = s['hdr']; iinst = h.index("Instructions Executed"); isamp = h.index("# Samples") if "# Samples" in h else None
    curline = None
    for r in s['rows']:
        if r[0] != '': curline = (s['file'].split('/')[-1], int(r[0])); text[curline] = r[1]
        if r[2] != '' and curline:
            try: v = int(r[iinst])
            except ValueError: v = 0
            agg[curline] += v; tot += v
            if isamp is not None:
                try: sv = int(r[isamp])
                except ValueError: sv = 0
                stall[curline] += sv; stot += sv
print(f"  total warp-instructions {tot}, samples {stot}")
print("  -- by executed instructions --")
for (f, l), v in agg.most_common(topn): print(f"  {100*v/max(tot,1):5.1f}% inst {100*stall[(f,l)]/max(stot,1):5.1f}% samp  {f}:{l}  {text[(f,l)].strip()[:95]}")
print("  -- by stall samples --")
for (f, l), v in stall.most_common(12): print(f"  {100*v/max(stot,1):5.1f}% samp  {f}:{l}  {text[(f,l)].strip()[:95]}")
