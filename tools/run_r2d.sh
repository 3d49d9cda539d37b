#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s)
python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_r2d.log
echo "pytest done $(( $(date +%s) - T0 ))s" > gpurun_out/ab_r2d.txt
run() {
  echo "== $* ($(( $(date +%s) - T0 ))s)" >> gpurun_out/ab_r2d.txt
  python bench.py --workload $* --steps 5 --warmup 3 --no-cpu --no-e2e --no-shared 2>gpurun_out/last.err | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], 'img/s %.0f'%d['value'], 'frac %.4f'%d['roofline']['frac'], 'step_frac %.4f'%d['roofline']['step_frac'], 'launch_ms %.4f'%d['roofline']['launch_ms'], d['clocks']['sm_mhz'] if d['clocks'] else None, d['parity']['max_rel'] if d['parity'] else None, d.get('graph_replay'))" >> gpurun_out/ab_r2d.txt 2>&1
  tail -3 gpurun_out/last.err >> gpurun_out/ab_r2d.err
}
run pinhole_b1024
run pinhole_b1024 --conf16
run simple_radial_b4096 --batch 2048
run simple_radial_b4096 --batch 2048 --conf16
run divisional_b1024 --batch 512
run radial_b1024
run pinhole_b1_320
run pinhole_b1_320 --rows-per-tile 8
run pinhole_b1_320 --rows-per-tile 16
GCLM_PERSISTENT=0 run pinhole_b1_320 --rows-per-tile 8
python bench.py --workload shared_b2048 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_shared1_r2d.json 2> gpurun_out/bench_shared1_r2d.err
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_r2d.json 2> gpurun_out/bench_ref_r2d.err
echo "done $(( $(date +%s) - T0 ))s" >> gpurun_out/ab_r2d.txt
