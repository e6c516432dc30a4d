#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
echo "== multi-gpu tests"; timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -x 2>&1 | tail -15
echo "== bench x2"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --n-per-gpu 20000000 --strong-total 40000000 --e2e-steps 2 > gpurun_out/r2g14_bench_n2.json 2> gpurun_out/r2g14_bench_n2.err; tail -5 gpurun_out/r2g14_bench_n2.err; cat gpurun_out/r2g14_bench_n2.json | head -c 3000
