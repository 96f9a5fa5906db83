TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29511 tools/dp_check.py p2p > gpurun_out/dp_p2p.json 2> gpurun_out/dp_p2p.err
$TR --master-port 29512 tools/dp_check.py multimem > gpurun_out/dp_mm.json 2> gpurun_out/dp_mm.err
$TR --master-port 29513 tools/dp_check.py nccl > gpurun_out/dp_nccl.json 2> gpurun_out/dp_nccl.err
$TR --master-port 29514 bench.py --gpus 2 --workload train --steps 40 --warmup 8 > gpurun_out/train_n2.json 2> gpurun_out/train_n2.err
$TR --master-port 29515 bench.py --gpus 2 --workload train --steps 40 --warmup 8 --unfused > gpurun_out/train_n2_unfused.json 2> gpurun_out/train_n2_unfused.err
