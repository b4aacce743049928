#!/bin/bash
# round 2, GPU call 13 (1 GPU): head_dim-72 vision tower parity; GEMM tests; banded tile walk A/B (AHA_GEMM_GROUP); prefill with the wide GEMM as default
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c13
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_gemm_gpu.py -m gpu -q -x -k "vl or gemm or llm_prefill" > $O.pytest.log 2>&1; echo "rc=$?" >> $O.pytest.log; tail -n 12 $O.pytest.log
for g in 1 8 4; do echo "== group_m $g"; AHA_GEMM_GROUP=$g timeout 300 python profiles/run_gemm.py 2>&1 | tail -n 9 | cut -c1-130; done > $O.gemm_group.txt 2>&1; cat $O.gemm_group.txt
for g in 1 8; do AHA_GEMM_GROUP=$g timeout 200 python profiles/run_prefill.py 3 2>&1 | tail -n 2; done
