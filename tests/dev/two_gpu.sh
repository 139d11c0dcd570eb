#!/bin/bash
# 2-GPU box: the NCCL / two-device tests and the bench line at N = 2
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/two_gpu.log
: > $L
nvidia-smi -L >> $L 2>&1
timeout 600 python -m pytest tests/test_gpu_polychromatic.py tests/test_gpu_multi_device.py -x -q -m gpu >> $L 2>&1; echo "pytest rc=$?" >> $L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/two_gpu_bench.json 2>> $L; echo "bench rc=$?" >> $L
python - >> $L <<'PY'
import json
d=json.loads(open('gpurun_out/two_gpu_bench.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'n_gpus',d['n_gpus'],'e2e',round(d['e2e']['value']))
print('c4',{k:v for k,v in d['c4_polychromatic'].items() if k not in('workload','roofline')})
print('c5',{k:v for k,v in d['c5_free_space'].items() if k not in('workload','roofline')})
PY
tail -12 $L
