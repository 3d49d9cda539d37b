#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s)
python -m pytest tests/test_dropin_gpu.py tests/test_gpu_baseline_sizes.py tests/test_gpu_parity.py tests/test_calibrate_batch.py -m gpu -q 2>&1 | tail -80 > gpurun_out/pytest_r2c.log
echo "pytest done $(( $(date +%s) - T0 ))s" >> gpurun_out/ab_r2c.txt
run() { # lib workload...
  lib=$1; shift
  echo "== $lib $* ($(( $(date +%s) - T0 ))s)" >> gpurun_out/ab_r2c.txt
  GCLM_LIB=$PWD/$lib python bench.py --workload $* --steps 5 --warmup 3 --no-cpu --no-e2e --no-shared 2>gpurun_out/last.err | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], 'img/s %.0f'%d['value'], 'frac %.4f'%d['roofline']['frac'], 'step_frac %.4f'%d['roofline']['step_frac'], 'launch_ms %.4f'%d['roofline']['launch_ms'], d['clocks']['sm_mhz'] if d['clocks'] else None, d['parity']['max_rel'] if d['parity'] else None, d.get('graph_replay'))" >> gpurun_out/ab_r2c.txt 2>&1
  tail -3 gpurun_out/last.err >> gpurun_out/ab_r2c.err
}
for lib in geocalib_b200/csrc/libgclm.so tools/variants/pin_b.so tools/variants/pin_c.so; do run $lib pinhole_b1024; done
for lib in geocalib_b200/csrc/libgclm.so tools/variants/d_nocols.so tools/variants/d_8x2.so tools/variants/d_8x2_nocols.so tools/variants/d_6x2.so; do
  run $lib simple_radial_b4096 --batch 2048
  run $lib divisional_b1024 --batch 512
done
run tools/variants/d_8x2_nocols.so radial_b1024
run geocalib_b200/csrc/libgclm.so pinhole_b1_320
GCLM_PERSISTENT=0 run geocalib_b200/csrc/libgclm.so pinhole_b1_320
python bench.py --workload shared_b2048 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_shared1_r2c.json 2> gpurun_out/bench_shared1_r2c.err
echo "ncu start $(( $(date +%s) - T0 ))s" >> gpurun_out/ab_r2c.txt
for wl in "sr simple_radial_b4096 1024" "div divisional_b1024 256"; do
  set -- $wl
  ncu --set full --clock-control none --import-source on -k regex:lm_pass_kernel -s 4 -c 1 -f -o /tmp/ncu_$1 \
    python bench.py --workload $2 --batch $3 --steps 1 --warmup 1 --no-cpu --no-e2e --no-check --no-shared > gpurun_out/ncu_$1_r2c.log 2>&1
  ncu -i /tmp/ncu_$1.ncu-rep --page raw --csv > gpurun_out/ncu_$1_r2c_raw.csv 2>/dev/null
  ncu -i /tmp/ncu_$1.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu_$1_r2c_source.csv 2>/dev/null
  ls -la /tmp/ncu_$1.ncu-rep >> gpurun_out/ab_r2c.txt
done
echo "done $(( $(date +%s) - T0 ))s" >> gpurun_out/ab_r2c.txt
du -sh gpurun_out >> gpurun_out/ab_r2c.txt
