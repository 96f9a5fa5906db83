#!/bin/bash
# the driver's scaling step at N = 2: reference arm, then our arm, launched with torch.distributed.run; plus the replica check
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
S=$(date +%s); timeout 900 $TR --master-port 29521 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/final_ref_n2.json 2> gpurun_out/final_ref_n2.err; echo "reference arm n2 rc $? in $(( $(date +%s) - S )) s"; tail -c 400 gpurun_out/final_ref_n2.json
S=$(date +%s); timeout 900 $TR --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/final_bench_n2.json 2> gpurun_out/final_bench_n2.err; echo "bench n2 rc $? in $(( $(date +%s) - S )) s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, d['e2e']['value'], d['roofline']['frac'], d['clocks'])
x=d.get('extras',{})
if 'train' in x: print('train', {k:(v.get('ms_per_step'), v.get('value'), v.get('config',{}).get('path')) for k,v in x['train'].items()})
PY
timeout 600 $TR --master-port 29523 tools/dp_check.py auto > gpurun_out/final_dp_n2.json 2> gpurun_out/final_dp_n2.err; echo "dp_check rc $?"; tail -c 600 gpurun_out/final_dp_n2.json
