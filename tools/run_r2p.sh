#!/bin/bash
mkdir -p gpurun_out
run() { lib=$1; shift
  echo "== $lib $*" >> gpurun_out/ab_r2p.txt
  GCLM_LIB=$PWD/$lib timeout 200 python bench.py --workload $* --steps 5 --warmup 3 --no-cpu --no-e2e --no-shared 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], 'img/s %.0f'%d['value'], 'frac %.4f'%d['roofline']['frac'], 'launch_ms %.4f'%d['roofline']['launch_ms'], d['clocks']['sm_mhz'], d['parity']['max_rel'] if d['parity'] else None)" >> gpurun_out/ab_r2p.txt 2>&1
}
for rep in 1 2; do
for lib in geocalib_b200/csrc/libgclm.so tools/variants/d_4x4.so tools/variants/d_5x3.so; do
  run $lib simple_radial_b4096 --batch 2048
  run $lib divisional_b1024 --batch 512
done; done
