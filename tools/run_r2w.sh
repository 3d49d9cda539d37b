#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/autograd_probe2.py simple_divisional,simple_radial 0.2 > gpurun_out/autograd_probe2.log 2>&1
