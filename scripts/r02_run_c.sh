#!/bin/bash
# radix-256 planes + equal-plane term + wide (N = 256) kernel: parity tests, kernel timing per variant, bench, conditioning sweep
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_primitives.py tests/test_gpu_mll.py tests/test_gpu_conditioning.py -q > gpurun_out/r02c_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02c_tests.log
OZ_BENCH_QUICK=1 timeout 300 python scripts/ozaki_bench.py > gpurun_out/r02c_oz_quick_v4.json 2> gpurun_out/r02c_oz_quick.err
GPB_OZ_KERNEL=3 OZ_BENCH_QUICK=1 timeout 300 python scripts/ozaki_bench.py > gpurun_out/r02c_oz_quick_v3.json 2>> gpurun_out/r02c_oz_quick.err
GPB_OZ_NOLOAD=1 OZ_BENCH_QUICK=1 timeout 300 python scripts/ozaki_bench.py > gpurun_out/r02c_oz_quick_v4_noload.json 2>> gpurun_out/r02c_oz_quick.err
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02c_bench_exact.json 2> gpurun_out/r02c_bench_exact.err
timeout 900 python scripts/cond_sweep.py 8192 > gpurun_out/r02c_cond_sweep.jsonl 2> gpurun_out/r02c_cond_sweep.err
tail -5 gpurun_out/r02c_tests.log; cat gpurun_out/r02c_bench_exact.json | head -c 400
