#!/bin/bash
# last validation of HEAD: full GPU suite, then the cost of the differentiable solve
mkdir -p gpurun_out; T=r2last
timeout 400 python -m pytest tests -m gpu -q --timeout 120 --timeout-method thread -p no:cacheprovider 2>&1 | tail -4 > gpurun_out/pytest_$T.log
timeout 60 python tools/autograd_bench.py > gpurun_out/autograd_bench_$T.jsonl 2> gpurun_out/autograd_bench_$T.err
