#!/bin/bash
# session-2 evidence call: full GPU test-suite (with durations), the bench line, ncu captures by phase of the three named
# articulations, ncu launch lists (kernel path + e2e), the reference arm -- ordered by importance, each step bounded
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/s2a_smi.txt 2>&1
timeout 700 python -m pytest tests -m gpu -q --durations=25 > $O/s2a_pytest.log 2>&1
timeout 200 python tools/variant_sweep.py --envs AntEnv,HumanoidEnv,SNUHumanoidEnv,HopperEnv,CheetahEnv,CartPoleSwingUpEnv --variants auto > $O/s2a_time.jsonl 2> $O/s2a_time.err
timeout 600 python bench.py --steps 4 --warmup 3 > $O/s2a_bench.json 2> $O/s2a_bench.err
NCU="ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:dfx_"
timeout 200 $NCU -o $O/prof_s2a_ant -f python tools/prof_step.py AntEnv 4096 > $O/s2a_ncu_ant.log 2>&1
timeout 200 $NCU -o $O/prof_s2a_humanoid -f python tools/prof_step.py HumanoidEnv 8192 > $O/s2a_ncu_hum.log 2>&1
timeout 200 $NCU -o $O/prof_s2a_snu -f python tools/prof_step.py SNUHumanoidEnv 4096 > $O/s2a_ncu_snu.log 2>&1
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/s2a_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-configs --ncu-range > $O/s2a_ncu_launches.log 2>&1
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/s2a_e2e_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-configs --ncu-range-e2e > $O/s2a_ncu_e2e.log 2>&1
timeout 400 python bench.py --impl reference --steps 4 --warmup 1 > $O/s2a_bench_reference.json 2> $O/s2a_bench_reference.err
tail -4 $O/s2a_pytest.log; cut -c1-300 $O/s2a_time.jsonl; tail -c 300 $O/s2a_bench.json; echo; tail -c 300 $O/s2a_bench_reference.json; ls -la $O | head -40
