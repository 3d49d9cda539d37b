#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "streams or two_devices or two_gpus" --timeout 200 --timeout-method thread -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/pytest_r2o.log
for wl in "pinhole_b1024 --batch 2048" "shared_b2048"; do
  timeout 200 python bench.py --workload $wl --steps 8 --warmup 3 --no-cpu --no-e2e --no-shared 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], d['config']['batch_per_gpu'], 'img/s %.0f'%d['value'], 'frac %.4f'%d['roofline']['frac'], 'step_frac %.4f'%d['roofline']['step_frac'], d['clocks']['sm_mhz'])" >> gpurun_out/sameB_r2o.txt
done
