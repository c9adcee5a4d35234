#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --no-header -rf -s -k "streaming or v1 or l1_feature" 2>&1 | tail -30 > $O/c18_stream.log; grep -E "passed|failed|^E  |FAILED" $O/c18_stream.log | head -20
