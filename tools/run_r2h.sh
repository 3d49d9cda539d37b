#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s)
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_r2h.log
for a in "0 persistent" "4 persistent"; do GCLM_LIB=$PWD/tools/variants/timers.so python tools/phase_timers.py $a; done > gpurun_out/phase_timers_r2h.txt 2>&1
run() {
  echo "== $* ($(( $(date +%s) - T0 ))s)" >> gpurun_out/ab_r2h.txt
  python bench.py --workload $* --steps 5 --warmup 3 --no-cpu --no-e2e --no-shared 2>gpurun_out/last.err | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], 'img/s %.0f'%d['value'], 'frac %.4f'%d['roofline']['frac'], 'launch_ms %.4f'%d['roofline']['launch_ms'], d['clocks']['sm_mhz'] if d['clocks'] else None, d.get('graph_replay'))" >> gpurun_out/ab_r2h.txt 2>&1
  tail -3 gpurun_out/last.err >> gpurun_out/ab_r2h.err
}
run pinhole_b1_320
run pinhole_b1_320 --rows-per-tile 4
run pinhole_b1_320 --rows-per-tile 8
GCLM_PERSISTENT=0 run pinhole_b1_320
echo "done $(( $(date +%s) - T0 ))s" >> gpurun_out/ab_r2h.txt
