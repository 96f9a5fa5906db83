#!/bin/bash
# the driver's scaling step at N = 8 on the final build: reference arm, our arm (render + extras.train), replica check
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
S=$(date +%s); timeout 600 $TR --master-port 29531 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/final_ref_n$N.json 2> gpurun_out/final_ref_n$N.err; echo "reference arm rc $? in $(( $(date +%s) - S )) s"
S=$(date +%s); timeout 600 $TR --master-port 29532 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/final_bench_n$N.json 2> gpurun_out/final_bench_n$N.err; echo "bench rc $? in $(( $(date +%s) - S )) s"
python - <<PY
import json
d=json.loads(open('gpurun_out/final_bench_n$N.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, d['e2e']['value'], d['roofline']['frac'], d['clocks'])
x=d.get('extras',{})
if 'train' in x: print('train', {k:(v.get('ms_per_step'), v.get('value'), v.get('config',{}).get('path')) for k,v in x['train'].items()})
PY
timeout 300 $TR --master-port 29533 tools/dp_check.py auto > gpurun_out/final_dp_n$N.json 2> gpurun_out/final_dp_n$N.err; echo "dp_check rc $?"; tail -c 500 gpurun_out/final_dp_n$N.json
