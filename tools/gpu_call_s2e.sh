#!/bin/bash
# tile-cooperative transition (staged in the scratch tile): env tests, memcheck + racecheck of it, bench, e2e launch list
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_envs.py -m gpu -q > $O/s2e_pytest_envs.log 2>&1
timeout 200 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_envs.py -m gpu -q -x -k "single_launch and (AntEnv or HumanoidEnv or CartPole) and 67" > $O/s2e_race_env.log 2>&1
timeout 200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_envs.py -m gpu -q -x -k "single_launch and (SNU or Hopper or Cheetah) and 67" > $O/s2e_mem.log 2>&1
timeout 500 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $O/s2e_bench.json 2> $O/s2e_bench.err
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/s2e_e2e_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-configs --ncu-range-e2e > $O/s2e_ncu_e2e.log 2>&1
tail -4 $O/s2e_pytest_envs.log; tail -3 $O/s2e_race_env.log; tail -3 $O/s2e_mem.log; tail -c 300 $O/s2e_bench.json; echo
