#!/bin/bash
mkdir -p gpurun_out; T=r2x
timeout 300 python -m pytest tests/test_frontend.py -m gpu -q -s --timeout 150 --timeout-method thread -p no:cacheprovider > gpurun_out/pytest_${T}_full.log 2>&1
grep -E "^\[frontend|passed|failed|^FAILED|^E  |Error" gpurun_out/pytest_${T}_full.log | head -40 > gpurun_out/pytest_$T.log
timeout 200 python tools/frontend_bench.py > gpurun_out/frontend_bench_$T.json 2> gpurun_out/frontend_bench_$T.err
