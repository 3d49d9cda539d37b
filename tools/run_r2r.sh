#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2 3; do
for lib in geocalib_b200/csrc/libgclm.so tools/variants/c28784bd.so; do
  echo -n "$lib " >> gpurun_out/ab_r2r.txt
  GCLM_LIB=$PWD/$lib timeout 200 python bench.py --workload pinhole_b1024 --steps 8 --warmup 3 --no-cpu --no-e2e --no-shared 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('img/s %.0f'%d['value'], 'frac %.4f'%d['roofline']['frac'], 'step_frac %.4f'%d['roofline']['step_frac'], d['clocks']['sm_mhz'])" >> gpurun_out/ab_r2r.txt
done; done
