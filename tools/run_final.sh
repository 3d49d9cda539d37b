# one gpurun call: full GPU suite, default bench, per-workload benches (TAG names the outputs)
mkdir -p gpurun_out; T=${TAG:-r1z}
timeout 400 python -m pytest tests -x -q -m gpu --timeout=120 > gpurun_out/pytest_$T.log 2>&1; tail -2 gpurun_out/pytest_$T.log
timeout 400 python bench.py > gpurun_out/bench_${T}_default.json 2> gpurun_out/bench_${T}_default.err; tail -c 900 gpurun_out/bench_${T}_default.json
rm -f gpurun_out/bench_${T}_models.jsonl
for wl in ${WORKLOADS:-simple_radial_b4096 radial_b1024 divisional_b1024 shared_b2048 pinhole_b1_320}; do
  timeout 300 python bench.py --workload $wl --no-cpu --no-e2e 2>/dev/null | tail -1 >> gpurun_out/bench_${T}_models.jsonl
done
python -c "
import json
for l in open('gpurun_out/bench_${T}_models.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print(d['config']['workload'], round(d['value']), round(r['frac'],4), round(r['launch_ms'],4), round(r['step_ms_per_pass'],4), d['clocks']['sm_mhz'])
"
