#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python tools/variant_sweep.py --envs SNUHumanoidEnv,HumanoidEnv,AntEnv --variants auto > $O/c5_time.jsonl 2> $O/c5_time.err
timeout 900 python -m pytest tests/test_gpu_atsize.py tests/test_gpu_envs.py tests/test_gpu_golden.py tests/test_gpu_edges.py -m gpu -q > $O/c5_pytest.log 2>&1
timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $O/c5_bench.json 2> $O/c5_bench.err
tail -4 $O/c5_pytest.log; cut -c1-300 $O/c5_time.jsonl; python -c "
import json; r=json.load(open('$O/c5_bench.json')); print('e2e', r['e2e']['value'], 'value', r['value'], {k:(v.get('value'), v.get('e2e',{}).get('value')) for k,v in r.get('configs',{}).items()})"
