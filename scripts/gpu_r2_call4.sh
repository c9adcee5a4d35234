#!/bin/bash
# round-2 GPU call 4: epilogue ablations of the TMA-staged conv kernel, the still-failing tests in detail
mkdir -p gpurun_out
O=gpurun_out
t() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu --no-header -rf -s 2>&1 | tail -${TAILN:-60} > $O/c4_$name.log; echo "== $name: $(grep -E 'passed|failed|error' $O/c4_$name.log | tail -1)"; grep -E "^(FAILED|ERROR)" $O/c4_$name.log | cut -c1-200; }
for f in "c1 as 64" "J=2, mask" "J=4, mask" "384->4x192" "msd_192_384" "unit 96"; do timeout 300 python scripts/ablate_tc.py "$f" 2>&1 | tee -a $O/c4_ablate.txt; done
TAILN=120 t step32 tests/test_gpu_parity.py -k "training_step_matches_reference_goldens_fp32"
TAILN=120 t step16 tests/test_gpu_parity.py -k "training_step_matches_reference_goldens_bf16"
TAILN=80 t disc tests/test_gpu_discrete.py -k "training_steps_run_in_bf16_and_graphs"
du -sh gpurun_out
