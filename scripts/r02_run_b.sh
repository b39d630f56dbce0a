#!/bin/bash
# radix-256 digit planes + lauum diagonal blocks on the int8 pipe: parity tests, kernel timing, bench, conditioning sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_primitives.py tests/test_gpu_mll.py tests/test_gpu_conditioning.py -x -q > gpurun_out/r02b_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02b_tests.log
OZ_BENCH_QUICK=1 timeout 300 python scripts/ozaki_bench.py > gpurun_out/r02b_oz_quick.json 2> gpurun_out/r02b_oz_quick.err
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02b_bench_exact.json 2> gpurun_out/r02b_bench_exact.err
timeout 900 python scripts/cond_sweep.py 8192 > gpurun_out/r02b_cond_sweep.jsonl 2> gpurun_out/r02b_cond_sweep.err
tail -3 gpurun_out/r02b_tests.log; cat gpurun_out/r02b_bench_exact.json | head -c 600
