#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python tools/sq_prof.py
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2j_sq_launches.csv python tools/sq_prof.py > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2j_sq_launches.csv')) if len(r)>10 and r[0].isdigit()]
d=collections.defaultdict(list)
for r in rows: d[r[4][:70]].append(float(r[-1].replace(',','')))
for k,v in d.items(): print(f"{len(v):4d} x {sum(v)/len(v)/1e3:9.1f} us  {k}")
PY
( time timeout 600 python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'roofline',d['roofline']['kernel'],d['roofline']['frac'],'step',d['roofline']['step']['frac'])
print('kinds',d['roofline']['kinds'])
for k,v in d['modes'].items(): print(k, v.get('MS/s'), v.get('ms'), v.get('hbm_frac'))
print('sweep',{k:(v['hbm_frac_per_gpu'],v['ms']) for k,v in d['mixed_mode_stream_sweep'].items()})
print('fallbacks',d['wbfm_tile_fallback_streams'],'rep',d['repeat_mismatches'])
PY
