#!/bin/bash
# slimmed tile transition (one body for both env families): full suite, memcheck + racecheck of the single-launch step, bench, launch lists
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q > $O/s2g_pytest.log 2>&1
timeout 200 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_envs.py -m gpu -q -x -k "single_launch and (AntEnv or CartPole) and 67" > $O/s2g_race_env.log 2>&1
timeout 200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_envs.py -m gpu -q -x -k "single_launch and (SNU or Hopper or HumanoidEnv) and 67" > $O/s2g_mem.log 2>&1
timeout 500 python bench.py --steps 4 --warmup 3 > $O/s2g_bench.json 2> $O/s2g_bench.err
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/s2g_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-configs --ncu-range > $O/s2g_ncu_launches.log 2>&1
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/s2g_e2e_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-configs --ncu-range-e2e > $O/s2g_ncu_e2e.log 2>&1
tail -3 $O/s2g_pytest.log; tail -3 $O/s2g_race_env.log; tail -3 $O/s2g_mem.log; tail -c 300 $O/s2g_bench.json; echo
