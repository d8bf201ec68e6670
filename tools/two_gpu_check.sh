#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/n2_n2.json 2> gpurun_out/n2_n2.err
echo "bench n2 rc=$?"; tail -3 gpurun_out/n2_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/n2_n2.json').read().strip().splitlines()[-1])
print('n_gpus',d['n_gpus'],'value',d['value'],'ms',d['ms_per_step'],'step frac',d['roofline']['step']['frac'])
for k,v in d['modes'].items(): print(k, v.get('MS/s'), v.get('ms'), v.get('hbm_frac'))
print('sweep',{k:(v['hbm_frac_per_gpu'],v['ms'],v['streams_rank0']) for k,v in d['mixed_mode_stream_sweep'].items()})
print('e2e',{k:v for k,v in d['e2e'].items() if k!='note'})
print('cpu',d['cpu_baseline'])
PY
