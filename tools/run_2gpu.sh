#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -k "two_gpus or shared_intrinsics" --timeout 240 --timeout-method thread -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/pytest_2gpu.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
nvidia-smi topo -m > gpurun_out/topo_2gpu.txt 2>&1
