#!/bin/bash
# integer fixed-point epilogue + RED write-out (v4), guard limit 2e6: parity tests, kernel timing, bench, conditioning sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_primitives.py tests/test_gpu_mll.py tests/test_gpu_conditioning.py tests/test_gpu_api.py -q > gpurun_out/r02d_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02d_tests.log
OZ_BENCH_QUICK=1 timeout 300 python scripts/ozaki_bench.py > gpurun_out/r02d_oz_quick_v4.json 2> gpurun_out/r02d_oz_quick.err
GPB_OZ_RED=0 OZ_BENCH_QUICK=1 timeout 300 python scripts/ozaki_bench.py > gpurun_out/r02d_oz_quick_v4_nored.json 2>> gpurun_out/r02d_oz_quick.err
GPB_OZ_KERNEL=3 OZ_BENCH_QUICK=1 timeout 300 python scripts/ozaki_bench.py > gpurun_out/r02d_oz_quick_v3.json 2>> gpurun_out/r02d_oz_quick.err
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02d_bench_exact.json 2> gpurun_out/r02d_bench_exact.err
GPB_TEST_SKIP_CONFIG3=1 timeout 900 python -m pytest tests/test_gpu_fullsize.py -q > gpurun_out/r02d_tests_fullsize.log 2>&1
timeout 900 python scripts/cond_sweep.py 8192 > gpurun_out/r02d_cond_sweep.jsonl 2> gpurun_out/r02d_cond_sweep.err
tail -5 gpurun_out/r02d_tests.log; tail -3 gpurun_out/r02d_tests_fullsize.log; cat gpurun_out/r02d_bench_exact.json | head -c 400
