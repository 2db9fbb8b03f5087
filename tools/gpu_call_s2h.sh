#!/bin/bash
# 2-GPU bench line on the final kernels (single-launch env.step; the c4 record at 8192 envs per GPU)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 240 $TR bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $O/s2h_bench_2gpu.json 2> $O/s2h_bench_2gpu.err
python -c "
import json; r=json.loads([l for l in open('$O/s2h_bench_2gpu.json') if l.startswith('{')][-1]); print('n', r['n_gpus'], 'value', r['value'], 'e2e', r['e2e']['value'], {k:(v.get('value'), v.get('e2e',{}).get('value')) for k,v in r.get('configs',{}).items()})"
tail -3 $O/s2h_bench_2gpu.err
