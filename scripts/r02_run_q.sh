#!/bin/bash
# panel x inverse-block products on the int8 pipe (GPB_OZ_PANELS): full GPU suite, then the exact bench with the switch off / on
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02q_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02q_tests.log
GPB_OZ_PANELS=0 timeout 600 python bench.py --steps 3 --warmup 3 --workload exact --no-cpu-baseline > gpurun_out/r02q_bench_panels0.json 2> gpurun_out/r02q_bench_panels0.err
GPB_OZ_PANELS=1 timeout 600 python bench.py --steps 3 --warmup 3 --workload exact --no-cpu-baseline > gpurun_out/r02q_bench_panels1.json 2> gpurun_out/r02q_bench_panels1.err
tail -15 gpurun_out/r02q_tests.log
for f in panels0 panels1; do head -c 330 gpurun_out/r02q_bench_$f.json; echo; tail -2 gpurun_out/r02q_bench_$f.err; done
