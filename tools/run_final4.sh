#!/bin/bash
# final validation of HEAD on one GPU: full GPU suite, smoke, CNN front-end timing
mkdir -p gpurun_out; T=r2y
timeout 700 python -m pytest tests -m gpu -q -s --timeout 150 --timeout-method thread -p no:cacheprovider > gpurun_out/pytest_${T}_full.log 2>&1
grep -E "^\[frontend|passed|failed|^FAILED|^E  |Error" gpurun_out/pytest_${T}_full.log | head -60 > gpurun_out/pytest_$T.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$T.log 2>&1
timeout 200 python tools/frontend_bench.py > gpurun_out/frontend_bench_$T.json 2> gpurun_out/frontend_bench_$T.err
