#!/bin/bash
# two GPUs with the end-of-round build: the tests that need them, SGPR / SVGP benches
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_sgpr.py tests/test_gpu_svgp.py -q > gpurun_out/r02z_tests_2gpu.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02z_tests_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --workload sgpr > gpurun_out/r02z_bench_sgpr_n2.json 2> gpurun_out/r02z_bench_sgpr_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --workload svgp > gpurun_out/r02z_bench_svgp_n2.json 2> gpurun_out/r02z_bench_svgp_n2.err
tail -3 gpurun_out/r02z_tests_2gpu.log; for f in sgpr_n2 svgp_n2; do head -c 260 gpurun_out/r02z_bench_$f.json; echo; tail -n 2 gpurun_out/r02z_bench_$f.err; done
