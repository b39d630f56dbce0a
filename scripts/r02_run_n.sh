#!/bin/bash
# eight GPUs: SGPR (config 4, row-sharded, strong scaling) and SVGP (config 5, data-parallel minibatches, weak scaling)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 5 --warmup 3 --workload sgpr > gpurun_out/r02ad8_bench_sgpr_n8.json 2> gpurun_out/r02ad8_bench_sgpr_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 10 --warmup 3 --workload svgp > gpurun_out/r02ad8_bench_svgp_n8.json 2> gpurun_out/r02ad8_bench_svgp_n8.err
for f in sgpr_n8 svgp_n8; do grep '^{' gpurun_out/r02ad8_bench_$f.json | head -c 300; echo; tail -n 2 gpurun_out/r02ad8_bench_$f.err; done
