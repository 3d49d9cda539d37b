#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -v -x --timeout 90 --timeout-method thread -p no:cacheprovider > gpurun_out/pytest_r2i.log 2>&1
echo "pytest rc $?" >> gpurun_out/pytest_r2i.log
