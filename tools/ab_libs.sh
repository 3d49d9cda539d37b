#!/bin/bash
# tools/ab_libs.sh <log> <workload> <lib...> : bench each build of libgclm on the same box, twice (ABAB)
# EXTRA="--batch 1024" adds bench.py flags, REPS="1 2 3" changes the repetitions
LOG=$1; WL=$2; shift 2
for rep in ${REPS:-1 2}; do
  for lib in "$@"; do
    echo -n "$WL $(basename $lib) rep$rep : " | tee -a $LOG
    GCLM_LIB=$PWD/$lib timeout 200 python bench.py --workload $WL --steps 4 --warmup 3 --no-e2e --no-cpu $EXTRA 2>&1 | tail -1 \
      | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])" | tee -a $LOG
  done
done
