#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_envs.py tests/test_gpu_dropin.py -m gpu -q > $O/c8_pytest.log 2>&1
timeout 400 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_envs.py -m gpu -q -k "action_map_folded and (Ant or SNU)" > $O/c8_race.log 2>&1
timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $O/c8_bench.json 2> $O/c8_bench.err
tail -3 $O/c8_pytest.log; tail -3 $O/c8_race.log; python -c "
import json; r=json.load(open('$O/c8_bench.json')); print('e2e', r['e2e']['value'], 'eager', r['e2e']['eager_env_step_loop']['value'], 'value', r['value'], r['kernel_ms'], {k:(v.get('value'), v.get('e2e',{}).get('value')) for k,v in r.get('configs',{}).items()})"
