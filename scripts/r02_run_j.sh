#!/bin/bash
# fused Gram -> digit planes, second version (byte trick, conflict-free staging): parity + bench
mkdir -p gpurun_out
GPB_TEST_SKIP_CONFIG3=1 timeout 1200 python -m pytest tests/test_gpu_sgpr.py tests/test_gpu_svgp.py -q > gpurun_out/r02j_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02j_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --workload sgpr --no-cpu-baseline > gpurun_out/r02j_bench_sgpr_fused.json 2> gpurun_out/r02j_bench_sgpr_fused.err
tail -3 gpurun_out/r02j_tests.log; head -c 250 gpurun_out/r02j_bench_sgpr_fused.json
