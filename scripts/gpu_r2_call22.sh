#!/bin/bash
# re-entry validation of HEAD: smoke, the GPU suite the driver's way (one process), default bench line, v3 quick bench
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
O=gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pytest"; timeout 600 python -m pytest tests -x -q -m gpu > $O/c22_tests.log 2>&1; echo "rc=$?"; tail -4 $O/c22_tests.log | cut -c1-300
echo "== bench"; timeout 400 python bench.py > $O/c22_bench.json 2> $O/c22_bench.err; echo "exit $?"; tail -c 600 $O/c22_bench.json
echo "== v3"; timeout 300 python bench.py --config v3 --batch 16 --steps 6 --warmup 3 --quick > $O/c22_bench_v3.json 2> $O/c22_bench_v3.err; echo "exit $?"; tail -c 500 $O/c22_bench_v3.json
