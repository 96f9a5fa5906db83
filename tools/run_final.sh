#!/bin/bash
# what the driver runs at round end, on one GPU: GPU tests, smoke, the reference arm, the default bench line
mkdir -p gpurun_out
S=$(date +%s); timeout 1200 python -m pytest tests/ -x -q -m gpu > gpurun_out/final_pytest.log 2>&1; echo "pytest rc $? in $(( $(date +%s) - S )) s" | tee -a gpurun_out/final_pytest.log; tail -2 gpurun_out/final_pytest.log
S=$(date +%s); timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc $? in $(( $(date +%s) - S )) s" | tee -a gpurun_out/final_smoke.log; tail -4 gpurun_out/final_smoke.log
S=$(date +%s); timeout 900 python bench.py --impl reference > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "reference arm rc $? in $(( $(date +%s) - S )) s"; tail -c 700 gpurun_out/final_bench_ref.json
S=$(date +%s); timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc $? in $(( $(date +%s) - S )) s"; tail -c 300 gpurun_out/final_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['clocks'])
print('cpu_baseline', d.get('cpu_baseline'))
x=d.get('extras',{})
print('extras keys', list(x.keys()))
if 'train' in x: print('train', {k:(v.get('ms_per_step'), v.get('value')) for k,v in x['train'].items()})
for k in x:
    if k!='train': print(k, json.dumps(x[k])[:300])
PY
# the ncu launch list of the same command (kernel shares; numbers printed under ncu are not bench values)
if [ "$1" = "--ncu" ]; then
  S=$(date +%s); timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final_bench_under_ncu.json 2> gpurun_out/final_bench_under_ncu.err; echo "ncu launch list rc $? in $(( $(date +%s) - S )) s"; wc -l gpurun_out/final_launches.csv
fi
