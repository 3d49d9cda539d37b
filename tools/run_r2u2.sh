#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/frontend_profile.py > gpurun_out/frontend_profile.txt 2> gpurun_out/frontend_profile.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/frontend_ncu_launches.csv python tools/frontend_profile.py ncu > /dev/null 2> gpurun_out/frontend_ncu.err
