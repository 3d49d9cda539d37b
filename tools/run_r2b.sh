#!/bin/bash
mkdir -p gpurun_out
(cat /sys/fs/cgroup/cpu.max; nproc; lscpu | head -25; free -g | head -3) > gpurun_out/host_r2b.txt 2>&1
python -m pytest tests/test_dropin_gpu.py -q -x 2>&1 | tail -120 > gpurun_out/pytest_dropin_r2b.log
python -m pytest tests/test_gpu_baseline_sizes.py tests/test_gpu_parity.py -q 2>&1 | tail -60 > gpurun_out/pytest_r2b.log
bash tools/ab_r2b.sh
python bench.py --workload shared_b2048 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_shared1_r2b.json 2> gpurun_out/bench_shared1_r2b.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r2b.json 2> gpurun_out/bench_ref_r2b.err
