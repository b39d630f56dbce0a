#!/bin/bash
# v4 integer-only write-out: parity, kernel timing, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_mll.py tests/test_gpu_primitives.py tests/test_gpu_kernels_ext.py -q > gpurun_out/r02g_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02g_tests.log
OZ_BENCH_QUICK=1 timeout 300 python scripts/ozaki_bench.py > gpurun_out/r02g_oz_quick_v4.json 2> gpurun_out/r02g_oz_quick.err
timeout 600 python bench.py --steps 3 --warmup 3 --workload exact > gpurun_out/r02g_bench_exact.json 2> gpurun_out/r02g_bench_exact.err
tail -3 gpurun_out/r02g_tests.log; head -c 300 gpurun_out/r02g_bench_exact.json
