#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s)
python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_r2g.log
for a in "0 persistent" "0 per_pass"; do GCLM_LIB=$PWD/tools/variants/timers.so python tools/phase_timers.py $a; done > gpurun_out/phase_timers_r2g.txt 2>&1
run() {
  echo "== $* ($(( $(date +%s) - T0 ))s)" >> gpurun_out/ab_r2g.txt
  python bench.py --workload $* --steps 5 --warmup 3 --no-cpu --no-e2e --no-shared 2>gpurun_out/last.err | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], 'img/s %.0f'%d['value'], 'frac %.4f'%d['roofline']['frac'], 'step_frac %.4f'%d['roofline']['step_frac'], 'launch_ms %.4f'%d['roofline']['launch_ms'], d['clocks']['sm_mhz'] if d['clocks'] else None, d['parity']['max_rel'] if d['parity'] else None, d.get('graph_replay'))" >> gpurun_out/ab_r2g.txt 2>&1
  tail -3 gpurun_out/last.err >> gpurun_out/ab_r2g.err
}
run pinhole_b1_320
GCLM_PERSISTENT=0 run pinhole_b1_320
run pinhole_b1024
run simple_radial_b4096 --batch 2048
run divisional_b1024 --batch 512
echo "done $(( $(date +%s) - T0 ))s" >> gpurun_out/ab_r2g.txt
