#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_engine.py tests/test_gpu_trainer_loop.py tests/test_gpu_autograd.py tests/test_gpu_full_size.py tests/test_gpu_data.py -x -q > gpurun_out/t_train.log 2>&1; echo "pytest rc $?" >> gpurun_out/t_train.log
tail -3 gpurun_out/t_train.log
if grep -q "pytest rc 0" gpurun_out/t_train.log; then
  for rep in 1 2; do
    timeout 300 python bench.py --workload train --steps 60 --warmup 10 > gpurun_out/train_bulk_$rep.json 2> gpurun_out/train_bulk.err; python -c "import json;d=json.loads(open('gpurun_out/train_bulk_$rep.json').read().strip().splitlines()[-1]);print('bulk', d['ms_per_step'], d['value'])"
    NERF_B200_LIB=$PWD/tools/ab/libnerf_b200_head.so timeout 300 python bench.py --workload train --steps 60 --warmup 10 > gpurun_out/train_head_$rep.json 2> gpurun_out/train_head.err; python -c "import json;d=json.loads(open('gpurun_out/train_head_$rep.json').read().strip().splitlines()[-1]);print('head', d['ms_per_step'], d['value'])"
  done
fi
