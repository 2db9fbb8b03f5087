#!/bin/bash
# the bench line with the bf16-tape Humanoid record (config C2 as named)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 170 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $O/s2i_bench.json 2> $O/s2i_bench.err
python -c "
import json; r=json.loads([l for l in open('$O/s2i_bench.json') if l.startswith('{')][-1]); print('value', r['value'], 'e2e', r['e2e']['value']); print({k:(v.get('value'), v.get('e2e',{}).get('value'), v.get('tape_mb_per_rollout'), v.get('kernel_ms'), v.get('error')) for k,v in r.get('configs',{}).items()})"
tail -3 $O/s2i_bench.err
