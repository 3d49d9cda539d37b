#!/bin/bash
# final validation of HEAD on one GPU: full GPU suite, sanitizer, bench lines
mkdir -p gpurun_out; T=r2z
timeout 500 python -m pytest tests -m gpu -q --timeout 150 --timeout-method thread -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/pytest_$T.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$T.log 2>&1
rm -f gpurun_out/sanitizer_$T.txt
for tool in memcheck racecheck synccheck; do
  echo "== $tool" >> gpurun_out/sanitizer_$T.txt
  timeout 400 compute-sanitizer --tool $tool python tools/sanitize_case.py 2>&1 | grep -v "^=========$" | tail -25 >> gpurun_out/sanitizer_$T.txt
done
rm -f gpurun_out/bench_${T}_models.jsonl
for wl in simple_radial_b4096 divisional_b1024 radial_b1024 shared_b2048 pinhole_b1_320; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 >> gpurun_out/bench_${T}_models.jsonl
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${T}_default.json 2> gpurun_out/bench_${T}_default.err
timeout 100 env GCLM_LIB=$PWD/tools/variants/timers.so python tools/phase_timers.py 0 persistent > gpurun_out/phase_timers_$T.txt 2>&1
