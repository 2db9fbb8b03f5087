#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR examples/train_shac.py --env AntEnv --num-envs $((8192*N)) --max-epochs 30 --profile-epochs 15 --log-interval 10 --graph 0 --out $O/shac_ant_$((8192*N))_${N}gpu.json > $O/mg_shac_${N}gpu.log 2>&1
timeout 600 $TR bench.py --gpus $N --steps 4 --warmup 3 > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err
tail -2 $O/mg_shac_${N}gpu.log | cut -c1-300; python -c "
import json; r=json.load(open('$O/bench_${N}gpu.json')); print('n', r['n_gpus'], 'value', r['value'], 'e2e', r['e2e']['value'], 'c4', {k:(v.get('value'), v.get('e2e',{}).get('value')) for k,v in r.get('configs',{}).items()})
s=json.load(open('$O/shac_ant_$((8192*N))_${N}gpu.json')); print(s.get('epoch_split_ms'))"
