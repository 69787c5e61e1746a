#!/bin/bash
# First GPU call of round 2 (run under gpurun from the repo root):
#   gpurun --timeout 900 -- 'bash scripts/round2_first_call.sh'
# 1. the whole GPU suite, the experimental kernels' gated tests included (k_banded.cu, k_order.cu: CPU-verified
#    through tests/emu only so far);
# 2. bench_banded.py: gtnb_ctc_loss at config 2 with banded = 0 (k_implicit.cu) and K = 1, 2, 4, 8;
# 3. the standard bench line.
# Everything lands in gpurun_out/r2a_*.
mkdir -p gpurun_out
(GTNB_EXPERIMENTAL=1 timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -25) > gpurun_out/r2a_tests.log
timeout 200 python scripts/bench_banded.py > gpurun_out/r2a_banded.json 2> gpurun_out/r2a_banded.err
GTNB_BENCH_DUMP=gpurun_out/r2a_dump timeout 240 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -5 gpurun_out/r2a_tests.log
grep -o '"banded_[0-9]*": {[^}]*"ms_median": [0-9.]*' gpurun_out/r2a_banded.json | head
