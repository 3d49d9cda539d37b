#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_autograd.py -m gpu -q -s --timeout 200 --timeout-method thread -p no:cacheprovider > gpurun_out/pytest_r2v_full.log 2>&1
grep -E "^\[|passed|failed|^FAILED|^E  " gpurun_out/pytest_r2v_full.log > gpurun_out/pytest_r2v.log
timeout 120 python tools/autograd_bench.py > gpurun_out/autograd_bench_r2v.jsonl 2> gpurun_out/autograd_bench_r2v.err
