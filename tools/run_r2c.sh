#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_dropin_gpu.py tests/test_gpu_baseline_sizes.py tests/test_gpu_parity.py -q 2>&1 | tail -80 > gpurun_out/pytest_r2c.log
run() { # lib workload...
  lib=$1; shift
  echo "== $lib $*" >> gpurun_out/ab_r2c.txt
  GCLM_LIB=$PWD/$lib python bench.py --workload $* --steps 5 --warmup 3 --no-cpu --no-e2e --no-shared 2>>gpurun_out/ab_r2c.err | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], 'img/s %.0f'%d['value'], 'frac %.4f'%d['roofline']['frac'], 'step_frac %.4f'%d['roofline']['step_frac'], 'launch_ms %.4f'%d['roofline']['launch_ms'], d['clocks']['sm_mhz'] if d['clocks'] else None, d['parity']['max_rel'] if d['parity'] else None, d.get('graph_replay'))" >> gpurun_out/ab_r2c.txt 2>&1
}
for lib in geocalib_b200/csrc/libgclm.so tools/variants/pin_b.so tools/variants/pin_c.so; do run $lib pinhole_b1024; done
for lib in geocalib_b200/csrc/libgclm.so tools/variants/d_nocols.so tools/variants/d_8x2.so tools/variants/d_8x2_nocols.so tools/variants/d_6x2.so tools/variants/d_4x4.so; do
  run $lib simple_radial_b4096 --batch 2048
  run $lib divisional_b1024 --batch 512
done
run tools/variants/d_8x2_nocols.so radial_b1024
run geocalib_b200/csrc/libgclm.so pinhole_b1_320
GCLM_PERSISTENT=0 run geocalib_b200/csrc/libgclm.so pinhole_b1_320
python bench.py --workload shared_b2048 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_shared1_r2c.json 2> gpurun_out/bench_shared1_r2c.err
ncu --set full --clock-control none --import-source on -k regex:lm_pass_kernel -s 4 -c 1 -f -o gpurun_out/ncu_sr_r2c \
  python bench.py --workload simple_radial_b4096 --batch 1024 --steps 1 --warmup 1 --no-cpu --no-e2e --no-check --no-shared > gpurun_out/ncu_sr_r2c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lm_pass_kernel -s 4 -c 1 -f -o gpurun_out/ncu_div_r2c \
  python bench.py --workload divisional_b1024 --batch 256 --steps 1 --warmup 1 --no-cpu --no-e2e --no-check --no-shared > gpurun_out/ncu_div_r2c.log 2>&1
