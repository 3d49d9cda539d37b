#!/bin/bash
# round-2 first GPU call: tests, default bench, reference arm, per-model bench lines, one ncu capture
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_r2a.log
python tools/cov_tolerance.py > gpurun_out/cov_tol_r2a.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r2a.json 2> gpurun_out/bench_ref_r2a.err
for wl in simple_radial_b4096 divisional_b1024 radial_b1024 pinhole_b1_320 shared_b2048; do
  python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu --no-e2e >> gpurun_out/bench_models_r2a.jsonl 2>> gpurun_out/bench_models_r2a.err
done
ncu --set full --clock-control none --import-source on -k regex:lm_pass_kernel -s 4 -c 1 -f -o gpurun_out/ncu_sr_r2a \
  python bench.py --workload simple_radial_b4096 --batch 1024 --steps 1 --warmup 1 --no-cpu --no-e2e --no-check > gpurun_out/ncu_sr_r2a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lm_pass_kernel -s 4 -c 1 -f -o gpurun_out/ncu_div_r2a \
  python bench.py --workload divisional_b1024 --batch 256 --steps 1 --warmup 1 --no-cpu --no-e2e --no-check > gpurun_out/ncu_div_r2a.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/smi_r2a.csv
lscpu | head -20 > gpurun_out/lscpu.txt
