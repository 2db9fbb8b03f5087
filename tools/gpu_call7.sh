#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 400 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/variant_sweep.py --no-time --envs AntEnv,HumanoidEnv,SNUHumanoidEnv,HopperEnv --variants auto,group32 > $O/c7_race.log 2>&1
timeout 600 python tools/variant_sweep.py --envs AntEnv,HumanoidEnv,SNUHumanoidEnv --variants auto,tile8L > $O/c7_time.jsonl 2> $O/c7_time.err
timeout 600 python tools/variant_sweep.py --envs AntEnv --variants tile32,tile16 --n 8192 > $O/c7_time_8192.jsonl 2>> $O/c7_time.err
timeout 1500 python -m pytest tests -m gpu -q -x > $O/c7_pytest.log 2>&1
tail -3 $O/c7_race.log; tail -3 $O/c7_pytest.log; cat $O/c7_time.jsonl $O/c7_time_8192.jsonl | cut -c1-330
