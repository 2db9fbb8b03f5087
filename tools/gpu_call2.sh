#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
NCU="ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:dfx_"
timeout 300 $NCU -o $O/prof_r2_hum -f python tools/prof_step.py HumanoidEnv 8192 > $O/c2_ncu_hum.log 2>&1
timeout 300 $NCU -o $O/prof_r2_snu -f python tools/prof_step.py SNUHumanoidEnv 4096 > $O/c2_ncu_snu.log 2>&1
timeout 300 $NCU -o $O/prof_r2_ant -f python tools/prof_step.py AntEnv 4096 > $O/c2_ncu_ant.log 2>&1
timeout 400 python tools/variant_sweep.py --envs HumanoidEnv,SNUHumanoidEnv --variants tile8,tile8L > $O/c2_time.jsonl 2> $O/c2_time.err
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_dropin.py > $O/c2_pytest.log 2>&1
timeout 600 python -m pytest tests/test_gpu_dropin.py -m gpu -q > $O/c2_pytest_dropin.log 2>&1
timeout 900 python bench.py --steps 4 --warmup 3 > $O/c2_bench.json 2> $O/c2_bench.err
tail -3 $O/c2_pytest.log; tail -3 $O/c2_pytest_dropin.log; tail -c 600 $O/c2_bench.json
