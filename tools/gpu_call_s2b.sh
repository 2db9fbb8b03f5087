#!/bin/bash
# single-launch env.step (transition as epilogue / prologue of the simulation launches) + joint-type-specialised tile kernels:
# env tests first, then timing, racecheck on the new prologue/epilogue, the full suite, the bench line
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_envs.py -m gpu -q -x > $O/s2b_pytest_envs.log 2>&1
timeout 200 python tools/variant_sweep.py --envs AntEnv,HumanoidEnv,SNUHumanoidEnv,HopperEnv,CheetahEnv,CartPoleSwingUpEnv --variants auto > $O/s2b_time.jsonl 2> $O/s2b_time.err
timeout 500 python bench.py --steps 4 --warmup 3 > $O/s2b_bench.json 2> $O/s2b_bench.err
timeout 300 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_envs.py -m gpu -q -x -k "single_launch and (AntEnv or HumanoidEnv or Hopper) and 40" > $O/s2b_race.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_envs.py -m gpu -q -x -k "single_launch and 67" > $O/s2b_mem.log 2>&1
timeout 600 python -m pytest tests -m gpu -q > $O/s2b_pytest.log 2>&1
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/s2b_e2e_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-configs --ncu-range-e2e > $O/s2b_ncu_e2e.log 2>&1
tail -4 $O/s2b_pytest_envs.log; cut -c1-300 $O/s2b_time.jsonl; tail -c 300 $O/s2b_bench.json; echo; tail -3 $O/s2b_race.log; tail -3 $O/s2b_mem.log; tail -3 $O/s2b_pytest.log
