#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --timeout 120 --timeout-method thread -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/pytest_r2q.log
timeout 100 env GCLM_LIB=$PWD/tools/variants/timers.so python tools/phase_timers.py 0 persistent > gpurun_out/phase_timers_r2q.txt 2>&1
for wl in "pinhole_b1_320" "pinhole_b1024" "simple_radial_b4096 --batch 2048"; do
timeout 200 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu --no-e2e --no-shared 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], 'img/s %.0f'%d['value'], 'frac %.4f'%d['roofline']['frac'], d['clocks']['sm_mhz'], d.get('graph_replay'))" >> gpurun_out/ab_r2q.txt
done
