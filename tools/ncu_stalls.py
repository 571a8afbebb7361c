#!/usr/bin/env python3
"""Top stall sites of one launch in an .ncu-rep (SASS view): python tools/ncu_stalls.py rep.ncu-rep <launch-skip> [topN]"""
import csv, io, subprocess, sys
rep, skip = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
print(rows[0][:2])
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > ci['# Samples'] and r[ci['# Samples']].isdigit()]
tot = sum(int(r[ci['# Samples']]) for r in data)
print('total samples', tot, 'instructions', len(data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(int(r[ci[s]] or 0) for r in data) for s in stalls}
print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
for r in sorted(data, key=lambda r: -int(r[ci['# Samples']]))[:topn]:
    st = max(stalls, key=lambda s: int(r[ci[s]] or 0))
    print(r[ci['# Samples']].rjust(5), r[ci['Instructions Executed']].rjust(7), st.ljust(18), r[ci['Address']][-5:],
          r[ci['Source']].strip()[:110])
