#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python tools/variant_sweep.py --envs AntEnv,HumanoidEnv,SNUHumanoidEnv --variants auto > $O/c16_time.jsonl 2> $O/c16_time.err
timeout 500 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/variant_sweep.py --no-time --envs AntEnv,HumanoidEnv,SNUHumanoidEnv --variants auto > $O/c16_race.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > $O/c16_pytest.log 2>&1
tail -2 $O/c16_race.log; tail -3 $O/c16_pytest.log; cut -c1-330 $O/c16_time.jsonl
