#!/bin/bash
# first GPU call of round 2: parity of every kernel variant, timings, sanitizer, test-suite, reference GPU bar
cd "$(dirname "$0")/.."
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/c1_smi.txt 2>&1
timeout 600 python tools/variant_sweep.py --no-time > $O/c1_parity.jsonl 2> $O/c1_parity.err
timeout 900 python tools/variant_sweep.py --envs AntEnv,HumanoidEnv,SNUHumanoidEnv --variants auto,tile8,tile16,tile32,group32 > $O/c1_time.jsonl 2> $O/c1_time.err
timeout 300 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/variant_sweep.py --no-time --envs AntEnv,HumanoidEnv --variants auto > $O/c1_race.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python tools/variant_sweep.py --no-time --envs AntEnv,SNUHumanoidEnv --variants auto,tile8 > $O/c1_mem.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/c1_pytest.log 2>&1
timeout 600 python oracle/ref_gpu_arm.py --env AntEnv --num-envs 4096 --horizon 32 --rollouts 2 --warmup 1 > $O/c1_refgpu.json 2> $O/c1_refgpu.err
tail -3 $O/c1_pytest.log; cat $O/c1_refgpu.json | tail -1; wc -l $O/c1_parity.jsonl $O/c1_time.jsonl
