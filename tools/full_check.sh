#!/bin/bash
# tools/full_check.sh -- on the GPU box (TAG=name gpurun -- 'bash tools/full_check.sh'): every GPU test, then the bench
# line and the ncu captures of the build (tools/final_run.sh), then a digest of the line.  Afterwards, here:
# python tools/make_profiles.py $TAG
export TAG=${TAG:-final}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
bash tools/final_run.sh $TAG
python - <<'PY'
import json, os
d = json.load(open(f"gpurun_out/{os.environ['TAG']}_bench.json"))
print('value', d['value'], 'ms', d['ms_per_step'], 'roofline', d['roofline']['kernel'], d['roofline']['frac'], 'step', d['roofline']['step']['frac'])
print('kinds', d['roofline']['kinds'])
for k, v in d['modes'].items():
    print(k, v.get('MS/s'), v.get('ms'), v.get('hbm_frac'))
print('sweep', {k: (v['hbm_frac_per_gpu'], v['ms']) for k, v in d['mixed_mode_stream_sweep'].items()})
print('e2e', {k: v for k, v in d['e2e'].items() if k != 'note'})
print('fallbacks', d.get('wbfm_tile_fallback_streams'), d.get('wbfm_serial_rerun_streams'))
PY
