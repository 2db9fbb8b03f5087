#!/bin/bash
# subtree-sum kinematics adjoint + path-sum tau adjoint + SNU on 16-environment tiles: racecheck, timing, the full GPU suite
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python tools/variant_sweep.py --envs AntEnv,HumanoidEnv,SNUHumanoidEnv --variants auto > $O/c13_time.jsonl 2> $O/c13_time.err
timeout 500 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/variant_sweep.py --no-time --envs AntEnv,HumanoidEnv,SNUHumanoidEnv,HopperEnv --variants auto,group32 > $O/c13_race.log 2>&1
timeout 900 python -m pytest tests -m gpu -q > $O/c13_pytest.log 2>&1
tail -3 $O/c13_race.log; tail -4 $O/c13_pytest.log; cut -c1-330 $O/c13_time.jsonl
