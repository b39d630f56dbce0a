#!/bin/bash
# end-of-round check of the final build: full GPU suite, SVGP bench, config 3 size
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02ag_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02ag_tests.log
timeout 300 python bench.py --workload svgp --steps 10 --warmup 3 > gpurun_out/r02ag_bench_svgp.json 2> gpurun_out/r02ag_bench_svgp.err
GPB_BENCH_N=100000 timeout 600 python bench.py --steps 2 --warmup 3 --workload exact --no-cpu-baseline > gpurun_out/r02ag_bench_exact_n100k.json 2> gpurun_out/r02ag_bench_exact_n100k.err
tail -4 gpurun_out/r02ag_tests.log; for f in svgp exact_n100k; do head -c 300 gpurun_out/r02ag_bench_$f.json; echo; tail -n 2 gpurun_out/r02ag_bench_$f.err; done
