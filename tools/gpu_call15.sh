#!/bin/bash
# ncu captures by phase after the subtree-sum kinematics adjoint
cd "$(dirname "$0")/.."
O=gpurun_out
NCU="ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:dfx_"
timeout 300 $NCU -o $O/prof_c15_ant -f python tools/prof_step.py AntEnv 4096 > $O/c15_ncu_ant.log 2>&1
timeout 300 $NCU -o $O/prof_c15_humanoid -f python tools/prof_step.py HumanoidEnv 8192 > $O/c15_ncu_hum.log 2>&1
timeout 300 $NCU -o $O/prof_c15_snu -f python tools/prof_step.py SNUHumanoidEnv 4096 > $O/c15_ncu_snu.log 2>&1
ls -la $O/prof_c15_*
