#!/bin/bash
# end-of-round records on one GPU: full suite, default bench (driver command), SVGP, config 2 size, small N
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02y_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02y_tests.log
timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/r02y_bench_default.json 2> gpurun_out/r02y_bench_default.err
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --workload svgp > gpurun_out/r02y_bench_svgp_n1.json 2> gpurun_out/r02y_bench_svgp_n1.err
GPB_BENCH_N=20000 timeout 600 python bench.py --steps 5 --warmup 3 --workload exact --no-cpu-baseline > gpurun_out/r02y_bench_exact_n20k.json 2> gpurun_out/r02y_bench_exact_n20k.err
timeout 300 python scripts/small_n_timing.py > gpurun_out/r02y_small_n.log 2>&1
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02y_bench_reference_arm.json 2> gpurun_out/r02y_bench_reference_arm.err
tail -4 gpurun_out/r02y_tests.log
for f in default svgp_n1 exact_n20k reference_arm; do head -c 300 gpurun_out/r02y_bench_$f.json; echo; tail -n 2 gpurun_out/r02y_bench_$f.err; done; cat gpurun_out/r02y_small_n.log
