#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
bash tools/final_run.sh ${TAG:-final}
python - <<'PY'
import json
d=json.load(open('gpurun_out/${TAG:-final}_bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'roofline',d['roofline']['kernel'],d['roofline']['frac'],'step',d['roofline']['step']['frac'])
print('kinds',d['roofline']['kinds'])
for k,v in d['modes'].items(): print(k, v.get('MS/s'), v.get('ms'), v.get('hbm_frac'))
print('sweep',{k:(v['hbm_frac_per_gpu'],v['ms']) for k,v in d['mixed_mode_stream_sweep'].items()})
print('e2e',d['e2e'])
print('fallbacks', d.get('wbfm_tile_fallback_streams'), d.get('wbfm_serial_rerun_streams'))
PY
