#!/bin/bash
# final evidence of the round: full GPU test-suite, the bench line, ncu launch lists (kernel path + e2e), full captures of the top kernels
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > $O/f_pytest.log 2>&1
timeout 1200 python bench.py --steps 4 --warmup 3 > $O/f_bench.json 2> $O/f_bench.err
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > $O/f_bench_reference.json 2> $O/f_bench_reference.err
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-configs --ncu-range > $O/f_ncu_launches.log 2>&1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_e2e_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-configs --ncu-range-e2e > $O/f_ncu_e2e.log 2>&1
NCU="ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:dfx_"
timeout 300 $NCU -o $O/prof_r02_ant -f python tools/prof_step.py AntEnv 4096 > $O/f_ncu_ant.log 2>&1
timeout 300 $NCU -o $O/prof_r02_humanoid -f python tools/prof_step.py HumanoidEnv 8192 > $O/f_ncu_hum.log 2>&1
timeout 300 $NCU -o $O/prof_r02_snu -f python tools/prof_step.py SNUHumanoidEnv 4096 > $O/f_ncu_snu.log 2>&1
tail -3 $O/f_pytest.log; tail -c 400 $O/f_bench.json; echo; tail -c 300 $O/f_bench_reference.json
