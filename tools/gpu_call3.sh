#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_atsize.py tests/test_gpu_envs.py -m gpu -q > $O/c3_pytest.log 2>&1
timeout 600 python tools/variant_sweep.py --envs AntEnv --variants tile8,tile16,tile32 --n 8192 > $O/c3_time_8192.jsonl 2> $O/c3_time.err
timeout 600 python tools/variant_sweep.py --envs AntEnv --variants tile8,tile16,tile32 --n 16384 > $O/c3_time_16384.jsonl 2>> $O/c3_time.err
timeout 600 python tools/variant_sweep.py --envs AntEnv,HumanoidEnv,SNUHumanoidEnv,HopperEnv,CheetahEnv,CartPoleSwingUpEnv --variants auto > $O/c3_time_auto.jsonl 2>> $O/c3_time.err
tail -5 $O/c3_pytest.log; cat $O/c3_time_8192.jsonl $O/c3_time_16384.jsonl | cut -c1-330
