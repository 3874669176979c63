#!/bin/bash
# ncu source-level capture of the level-0 multi-pick FPS launch (SASS view with per-instruction stall samples), summarised
# on the box: the 60 instructions with the most stall samples.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k "regex:fps_multi" -s 3 -c 1 -o gpurun_out/r02_prof_fps_src python scripts/one_forward.py tc serial > gpurun_out/r02_prof_fps_src.log 2>&1
ncu -i gpurun_out/r02_prof_fps_src.ncu-rep --page source --csv --print-source sass > gpurun_out/r02_fps_source.csv 2>/dev/null
python - <<'P'
import csv
rows = list(csv.reader(open('gpurun_out/r02_fps_source.csv')))
hdr = None
for i, r in enumerate(rows):
    if 'Source' in r and any('Sampling' in c for c in r):
        hdr = i; break
if hdr is None:
    print('no header found', rows[:3]); raise SystemExit
h = rows[hdr]
src = h.index('Source')
samp = [j for j, c in enumerate(h) if c.startswith('# Samples') or c == 'Warp Stall Sampling (All Samples)' or 'Samples' in c]
print('columns:', [h[j] for j in samp][:6])
sj = samp[0]
data = []
for r in rows[hdr + 1:]:
    try:
        data.append((float(r[sj].replace(',', '') or 0), r[src], r))
    except Exception:
        pass
tot = sum(d[0] for d in data)
print('total samples', tot, 'instructions', len(data))
stall_cols = [j for j, c in enumerate(h) if c.startswith('stall_')]
for v, s, r in sorted(data, key=lambda t: -t[0])[:45]:
    top = sorted(((float(r[j].replace(',', '') or 0), h[j]) for j in stall_cols), reverse=True)[:2]
    print(f'{v:8.0f} {100 * v / tot:5.1f}%  {s[:70]:70s} {top}')
P
rm -f gpurun_out/r02_prof_fps_src.ncu-rep
ls -la gpurun_out/r02_fps_source.csv
