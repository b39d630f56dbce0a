#!/bin/bash
# eight GPUs: SGPR (config 4, row-sharded, strong scaling) and SVGP (config 5, data-parallel minibatches, weak scaling)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 --workload sgpr > gpurun_out/r02ab8_bench_sgpr_n8.json 2> gpurun_out/r02ab8_bench_sgpr_n8.err
for f in sgpr_n8; do head -c 300 gpurun_out/r02ab8_bench_$f.json; echo; tail -2 gpurun_out/r02ab8_bench_$f.err; done
