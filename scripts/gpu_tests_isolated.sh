#!/bin/bash
# Every GPU test file (and the riskiest groups on their own) in its OWN process: a device-side trap in one kernel poisons
# only that process's CUDA context.  Usage (GPU box): bash scripts/gpu_tests_isolated.sh ; summary in gpurun_out/pytest_iso.txt
mkdir -p gpurun_out
: > gpurun_out/pytest_iso.txt
run() {   # name, pytest args...
  local name=$1; shift
  timeout 900 python -m pytest "$@" -q -m gpu -s > gpurun_out/pytest_$name.log 2>&1
  local rc=$?
  echo "== $name rc=$rc  $(grep -E 'passed|failed|error' gpurun_out/pytest_$name.log | tail -1)" | tee -a gpurun_out/pytest_iso.txt
  grep -E "^(FAILED|ERROR)" gpurun_out/pytest_$name.log | cut -c1-220 | tee -a gpurun_out/pytest_iso.txt
  grep -E "rel-L2|gradient (rel|cos)|cos " gpurun_out/pytest_$name.log | cut -c1-200 >> gpurun_out/pytest_iso.txt
}
run parity_pqmf tests/test_gpu_parity.py -k "pqmf"
run parity_rest tests/test_gpu_parity.py -k "not pqmf"
run tc_conv tests/test_gpu_tc.py -k "not fused_dilated"
run tc_unit tests/test_gpu_tc.py -k "fused_dilated"
run x3 tests/test_gpu_x3.py
run engine tests/test_gpu_engine.py
run discrete tests/test_gpu_discrete.py
for f in tests/test_gpu_*.py; do
  case $f in *parity*|*_tc.py|*x3*|*engine*|*discrete*) ;; *) run $(basename $f .py) $f ;; esac
done
