#!/bin/bash
# epilogue reading the stepped state from shared memory: env tests, bench, e2e launch list
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_envs.py -m gpu -q > $O/s2d_pytest_envs.log 2>&1
timeout 500 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $O/s2d_bench.json 2> $O/s2d_bench.err
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/s2d_e2e_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-configs --ncu-range-e2e > $O/s2d_ncu_e2e.log 2>&1
tail -4 $O/s2d_pytest_envs.log; tail -c 300 $O/s2d_bench.json; echo
