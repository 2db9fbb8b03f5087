#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_atsize.py -m gpu -q > $O/c4_pytest.log 2>&1
timeout 900 python tools/variant_sweep.py --envs AntEnv --variants tile32,tile32P,tile32,tile32P > $O/c4_time_ant.jsonl 2> $O/c4_time.err
timeout 900 python tools/variant_sweep.py --envs HumanoidEnv,SNUHumanoidEnv --variants tile8,tile8P > $O/c4_time_big.jsonl 2>> $O/c4_time.err
NCU="ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:dfx_"
timeout 300 $NCU -o $O/prof_r2b_snu -f python tools/prof_step.py SNUHumanoidEnv 4096 > $O/c4_ncu_snu.log 2>&1
timeout 300 $NCU -o $O/prof_r2b_hum -f python tools/prof_step.py HumanoidEnv 8192 > $O/c4_ncu_hum.log 2>&1
tail -5 $O/c4_pytest.log; cat $O/c4_time_ant.jsonl $O/c4_time_big.jsonl | cut -c1-300
