#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python tools/variant_sweep.py --envs HumanoidEnv,SNUHumanoidEnv --variants tile8,tile16 > $O/c12_e16.jsonl 2> $O/c12.err
for nw in 8 12 16 20 24; do
  if [ $nw = 16 ]; then L=diffrl_b200/libdfx.so; else L=build/ab/libdfx_nw$nw.so; fi
  DFX_LIBRARY=$PWD/$L timeout 200 python tools/variant_sweep.py --envs AntEnv --variants tile32 2>> $O/c12.err | sed "s/^{/{\"nw\": $nw, /" >> $O/c12_nw.jsonl
done
cut -c1-260 $O/c12_e16.jsonl $O/c12_nw.jsonl
