#!/bin/bash
# round-2 evidence: launch list, ncu full captures (raw CSV summaries only), compute-sanitizer, per-workload bench lines
mkdir -p gpurun_out; T=${TAG:-r2}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$T.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/launches_$T.log 2>&1
for wl in "pin pinhole_b1024 1024" "sr simple_radial_b4096 1024" "div divisional_b1024 256" "rad radial_b1024 1024"; do
  set -- $wl
  ncu --set full --clock-control none --import-source on -k regex:lm_pass_kernel -s 4 -c 1 -f -o /tmp/ncu_$1 \
    python bench.py --workload $2 --batch $3 --steps 1 --warmup 1 --no-cpu --no-e2e --no-check --no-shared > gpurun_out/ncu_$1_$T.log 2>&1
  ncu -i /tmp/ncu_$1.ncu-rep --page raw --csv > gpurun_out/ncu_$1_${T}_raw.csv 2>/dev/null
done
for tool in memcheck racecheck synccheck; do
  echo "== $tool" >> gpurun_out/sanitizer_$T.txt
  compute-sanitizer --tool $tool python tools/sanitize_case.py 2>&1 | grep -v "^=========$" | tail -25 >> gpurun_out/sanitizer_$T.txt
done
rm -f gpurun_out/bench_${T}_models.jsonl
for wl in simple_radial_b4096 divisional_b1024 radial_b1024 shared_b2048 pinhole_b1_320; do
  python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 >> gpurun_out/bench_${T}_models.jsonl
done
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${T}_default.json 2> gpurun_out/bench_${T}_default.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_${T}_reference.json 2> gpurun_out/bench_${T}_reference.err
timeout 120 env GCLM_LIB=$PWD/tools/variants/timers.so python tools/phase_timers.py 0 persistent > gpurun_out/phase_timers_$T.txt 2>&1
python bench.py --workload pinhole_b1024 --conf16 --steps 10 --warmup 3 --no-cpu --no-e2e --no-shared 2>/dev/null | tail -1 > gpurun_out/bench_${T}_conf16.json
timeout 300 python -m pytest tests -m gpu -q --timeout 120 --timeout-method thread -p no:cacheprovider 2>&1 | tail -5 > gpurun_out/pytest_$T.log
