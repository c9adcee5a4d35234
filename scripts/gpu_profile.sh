#!/bin/bash
# In-situ profiles of the graphed step (CUPTI timeline), per-launch roofline table, roofline-by-ablation of the conv kernel,
# one `ncu --set full` capture of the dominant conv launch and of a wgrad launch.
# Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_profile.sh'; copy what matters from gpurun_out/ to profiles/.
mkdir -p gpurun_out
echo "== trace"; timeout 600 python scripts/trace_step.py > gpurun_out/trace_step.txt 2> gpurun_out/trace_step.err; echo "exit $?"; grep "====" gpurun_out/trace_step.txt
echo "== layers"; timeout 600 python scripts/profile_layers.py > gpurun_out/layers.txt 2> gpurun_out/layers.err; echo "exit $?"; head -4 gpurun_out/layers.txt | cut -c1-200
echo "== ablate"; timeout 600 python scripts/ablate_tc.py > gpurun_out/ablate.log 2>&1; echo "exit $?"; grep -E "^==|default" gpurun_out/ablate.log | cut -c1-100
echo "== ncu conv"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc2_kernel -s 3 -c 1 -o gpurun_out/prof_conv -f python scripts/prof_conv_tc.py > gpurun_out/ncu_conv.log 2>&1; echo "exit $?"
echo "== ncu wgrad"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 1 -c 1 -o gpurun_out/prof_wgrad -f python scripts/prof_conv_tc.py > gpurun_out/ncu_wgrad.log 2>&1; echo "exit $?"
