#!/bin/bash
# SGPR / SVGP with the fused Gram -> digit planes kernel: parity tests, bench with and without the fusion
mkdir -p gpurun_out
GPB_TEST_SKIP_CONFIG3=1 timeout 1200 python -m pytest tests/test_gpu_sgpr.py tests/test_gpu_svgp.py tests/test_gpu_fullsize.py tests/test_gpu_boundary.py -q > gpurun_out/r02i_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02i_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --workload sgpr --no-cpu-baseline > gpurun_out/r02i_bench_sgpr_fused.json 2> gpurun_out/r02i_bench_sgpr_fused.err
GPB_SGPR_FUSED=0 timeout 600 python bench.py --steps 5 --warmup 3 --workload sgpr --no-cpu-baseline > gpurun_out/r02i_bench_sgpr_unfused.json 2> gpurun_out/r02i_bench_sgpr_unfused.err
tail -3 gpurun_out/r02i_tests.log; head -c 250 gpurun_out/r02i_bench_sgpr_fused.json; echo; head -c 250 gpurun_out/r02i_bench_sgpr_unfused.json
