#!/bin/bash
# final evidence of round 2: full GPU suite, the bench line (+ reference arm), ncu captures by phase of the three named
# articulations, ncu launch lists (kernel path + e2e), a short SHAC run of the repo trainer on the single-launch env.step
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q --durations=8 > $O/s2f_pytest.log 2>&1
timeout 600 python bench.py --steps 4 --warmup 3 > $O/s2f_bench.json 2> $O/s2f_bench.err
NCU="ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:dfx_"
timeout 200 $NCU -o $O/prof_s2f_ant -f python tools/prof_step.py AntEnv 4096 > $O/s2f_ncu_ant.log 2>&1
timeout 200 $NCU -o $O/prof_s2f_humanoid -f python tools/prof_step.py HumanoidEnv 8192 > $O/s2f_ncu_hum.log 2>&1
timeout 200 $NCU -o $O/prof_s2f_snu -f python tools/prof_step.py SNUHumanoidEnv 4096 > $O/s2f_ncu_snu.log 2>&1
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/s2f_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-configs --ncu-range > $O/s2f_ncu_launches.log 2>&1
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/s2f_e2e_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-configs --ncu-range-e2e > $O/s2f_ncu_e2e.log 2>&1
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > $O/s2f_bench_reference.json 2> $O/s2f_bench_reference.err
timeout 200 python examples/train_shac.py --env AntEnv --num-envs 64 --max-epochs 30 --log-interval 10 --out $O/s2f_shac_ant64.json > $O/s2f_shac.log 2>&1
tail -4 $O/s2f_pytest.log; tail -c 300 $O/s2f_bench.json; echo; tail -c 200 $O/s2f_bench_reference.json; echo; tail -3 $O/s2f_shac.log; ls $O | grep s2f
