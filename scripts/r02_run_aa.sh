#!/bin/bash
# balanced streamed blocks: SGPR / SVGP parity tests, 1-GPU SGPR bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sgpr.py tests/test_gpu_svgp.py tests/test_gpu_kernels_ext.py -q > gpurun_out/r02aa_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02aa_tests.log
timeout 600 python bench.py --workload sgpr --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02aa_bench_sgpr_n1.json 2> gpurun_out/r02aa_bench_sgpr_n1.err
tail -3 gpurun_out/r02aa_tests.log; head -c 300 gpurun_out/r02aa_bench_sgpr_n1.json; echo; tail -n 2 gpurun_out/r02aa_bench_sgpr_n1.err
