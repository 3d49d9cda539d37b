#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
for lib in tools/variants/c28784bd.so geocalib_b200/csrc/libgclm.so; do
  for wl in "pinhole_b1024" "simple_radial_b4096 --batch 2048" "divisional_b1024 --batch 512"; do
  echo -n "$lib $wl " >> gpurun_out/ab_r2u.txt
  GCLM_LIB=$PWD/$lib timeout 200 python bench.py --workload $wl --steps 8 --warmup 3 --no-cpu --no-e2e --no-shared 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('img/s %.0f'%d['value'], 'frac %.4f'%d['roofline']['frac'], 'step_frac %.4f'%d['roofline']['step_frac'], d['clocks']['sm_mhz'])" >> gpurun_out/ab_r2u.txt
  done
done; done
