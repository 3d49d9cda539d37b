#!/bin/bash
# registers / spills of the hot pass-kernel instantiations of a build: tools/regs.sh [extra nvcc flags]
cd "$(dirname "$0")/.."
GCLM_OUT=${GCLM_OUT:-/tmp/regs_probe.so} bash geocalib_b200/csrc/build.sh -Xptxas -v "$@" 2>&1 | \
  awk '/Compiling entry function/ {n=$0; sub(/.*lm_pass_kernel/,"lm_pass_kernel",n); sub(/EEvNS.*/,"",n)} /Used [0-9]+ registers/ && n ~ /lm_pass_kernel/ {match($0,/Used [0-9]+ registers/); r=substr($0,RSTART,RLENGTH); print n, r; n=""} /spill/ && n ~ /ELb0ELi1ELb1/ {print "   ", $0}' | grep -E "ELb0ELi1ELb1|ELb1ELi1ELb1|ELb0ELi2ELb1|spill"
