#!/bin/bash
# final validation of HEAD on one GPU: full GPU suite, smoke, default bench, cost of the differentiable solve
mkdir -p gpurun_out; T=r2zz
timeout 600 python -m pytest tests -m gpu -q --timeout 150 --timeout-method thread -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/pytest_$T.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$T.log 2>&1
timeout 120 python tools/autograd_bench.py > gpurun_out/autograd_bench_$T.jsonl 2> gpurun_out/autograd_bench_$T.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${T}_default.json 2> gpurun_out/bench_${T}_default.err
