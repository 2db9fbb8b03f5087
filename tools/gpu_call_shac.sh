#!/bin/bash
# SHAC return curves: the reference's UNCHANGED algorithms/shac.py on this repo's dflex (3 seeds) and the repo's own trainer (3 seeds)
cd "$(dirname "$0")/.."
O=gpurun_out
for s in 0 1 2; do
  timeout 1500 python tools/run_ref_shac.py --dflex ours --env ant --seed $s --out $O/ref_shac_ant_ours_seed$s.json > $O/ref_shac_ant_ours_seed$s.log 2>&1 &
done
for s in 0 1 2; do
  timeout 1500 python examples/train_shac.py --env AntEnv --num-envs 64 --seed $s --out $O/own_shac_ant_seed$s.json > $O/own_shac_ant_seed$s.log 2>&1 &
done
wait
for s in 0 1 2; do tail -1 $O/ref_shac_ant_ours_seed$s.log | cut -c1-400; tail -1 $O/own_shac_ant_seed$s.log | cut -c1-300; done
