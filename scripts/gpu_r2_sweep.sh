#!/bin/bash
# parameter sweep of the v2 step (timed region only)
mkdir -p gpurun_out
O=gpurun_out
b() { name=$1; shift; env "$@" timeout 300 python bench.py --quick --steps 16 --warmup 4 > $O/sw_$name.json 2> $O/sw_$name.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/sw_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"], 3))
except Exception as e:
    print("$name parse:", e); print(open("gpurun_out/sw_$name.err").read()[-400:])
PY
}
b base A=1
b streams4 RAVE_DISC_STREAMS=4
b streams12 RAVE_DISC_STREAMS=12
b chain3000 RAVE_TC_CHAIN=3000
b chain9000 RAVE_TC_CHAIN=9000
b chain20000 RAVE_TC_CHAIN=20000
b estages1 RAVE_TC_ESTAGES=1
b nofuse RAVE_FUSE_UNITS=0
b base2 A=1
