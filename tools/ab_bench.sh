#!/bin/bash
# A/B the alternative builds of libgclm (same box, same run): tools/ab_bench.sh <log> <workloads...>
# Variants: every geocalib_b200/csrc/libgclm*.so present; GCLM_PACK=0/1 for each.
LOG=$1; shift
for wlk in "$@"; do
  for lib in geocalib_b200/csrc/libgclm*.so; do
    for pk in ${PACKS:-1}; do
      echo -n "$wlk $(basename $lib) pack=$pk : " | tee -a $LOG
      GCLM_LIB=$PWD/$lib GCLM_PACK=$pk python bench.py --workload $wlk --steps 3 --warmup 2 --no-e2e --no-cpu 2>&1 | tail -1 \
        | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['frac'],4), d['clocks']['reasons'])" | tee -a $LOG
    done
  done
done
