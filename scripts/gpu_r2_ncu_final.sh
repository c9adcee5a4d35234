#!/bin/bash
# ncu evidence of the final tree: --set full of the dominant launch (DRAM traffic) and the launch list of the bench command
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
O=gpurun_out
timeout 200 ncu --set full --clock-control none -k regex:conv_tc2 -c 3 -o /tmp/r2_dominant -f python scripts/ncu_dominant.py > $O/c29_ncu_dom.log 2>&1; echo "ncu dominant exit $?"
ncu -i /tmp/r2_dominant.ncu-rep --page raw --csv > $O/r2_ncu_raw_dominant.csv 2>/dev/null; wc -c $O/r2_ncu_raw_dominant.csv
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/r2_launches_final.csv python bench.py --steps 2 --warmup 1 --no-graphs --no-cpu-baseline --no-cudnn-baseline --quick > $O/c29_ncu_list.log 2>&1; echo "ncu list exit $?"; wc -l $O/r2_launches_final.csv
