#!/bin/bash
# round-2 GPU call 6: per-instruction stall profile of the TMA-staged conv kernel (c1 forward, fused dgrad)
mkdir -p gpurun_out
O=gpurun_out
for w in c1 bwd2; do
  for m in 1 0; do
    RAVE_TC_ETMA=$m timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 2 -c 1 -o /tmp/one_${w}_$m -f python scripts/ncu_one.py $w > $O/c6_ncu_${w}_$m.log 2>&1
    ncu -i /tmp/one_${w}_$m.ncu-rep --page source --csv > $O/c6_src_${w}_etma$m.csv 2>/dev/null
    ncu -i /tmp/one_${w}_$m.ncu-rep --page raw --csv > $O/c6_raw_${w}_etma$m.csv 2>/dev/null
  done
done
ls -la $O/c6_* | head; du -sh $O
