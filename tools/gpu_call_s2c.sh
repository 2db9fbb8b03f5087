#!/bin/bash
# A/B of the merged phases and the two-column Cholesky (same box), racecheck of the new code, env tests, full suite, bench
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_envs.py -m gpu -q -x > $O/s2c_pytest_envs.log 2>&1
for tag in main nofuse cholu main nofuse; do
  if [ $tag = main ]; then L=diffrl_b200/libdfx.so; else L=build/ablib/libdfx_$tag.so; fi
  DFX_LIBRARY=$PWD/$L timeout 200 python tools/variant_sweep.py --envs AntEnv,HumanoidEnv,SNUHumanoidEnv,HopperEnv --variants auto 2>> $O/s2c_time.err | sed "s/^{/{\"lib\": \"$tag\", /" >> $O/s2c_time.jsonl
done
timeout 400 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/variant_sweep.py --no-time --envs AntEnv,HumanoidEnv,SNUHumanoidEnv,HopperEnv --variants auto > $O/s2c_race.log 2>&1
timeout 300 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_envs.py -m gpu -q -x -k "single_launch and (AntEnv or HumanoidEnv or Hopper) and 40" > $O/s2c_race_env.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_envs.py -m gpu -q -x -k "single_launch and 67" > $O/s2c_mem.log 2>&1
timeout 600 python -m pytest tests -m gpu -q > $O/s2c_pytest.log 2>&1
timeout 500 python bench.py --steps 4 --warmup 3 > $O/s2c_bench.json 2> $O/s2c_bench.err
tail -4 $O/s2c_pytest_envs.log; cut -c1-330 $O/s2c_time.jsonl; tail -c 300 $O/s2c_bench.json; echo; tail -3 $O/s2c_race.log; tail -3 $O/s2c_race_env.log; tail -3 $O/s2c_mem.log; tail -3 $O/s2c_pytest.log
