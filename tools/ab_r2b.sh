#!/bin/bash
# A/B of libgclm builds on one box: base (round-1 math) vs new default vs scalar accumulate
mkdir -p gpurun_out
tools/ubench_fp32x2 > gpurun_out/ubench_r2b.txt 2>&1
for lib in tools/variants/libgclm_base.so geocalib_b200/csrc/libgclm.so tools/variants/libgclm_sacc.so; do
  for wl in "pinhole_b1024" "simple_radial_b4096 --batch 2048" "divisional_b1024 --batch 512" "radial_b1024"; do
    echo "== $lib $wl" >> gpurun_out/ab_r2b.txt
    GCLM_LIB=$PWD/$lib python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu --no-e2e --no-shared 2>>gpurun_out/ab_r2b.err | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], 'img/s %.0f'%d['value'], 'frac %.4f'%d['roofline']['frac'], 'step_frac %.4f'%d['roofline']['step_frac'], 'launch_ms %.4f'%d['roofline']['launch_ms'], d['clocks'], d['parity'])" >> gpurun_out/ab_r2b.txt 2>&1
  done
done
