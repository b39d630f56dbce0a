#!/bin/bash
# final-form bench line (roofline_hbm block), configs 2 and 3 records, small-N timing
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/r02p_bench_default.json 2> gpurun_out/r02p_bench_default.err
GPB_BENCH_N=20000 timeout 600 python bench.py --steps 5 --warmup 3 --workload exact --no-cpu-baseline > gpurun_out/r02p_bench_exact_n20k.json 2> gpurun_out/r02p_bench_exact_n20k.err
GPB_BENCH_N=100000 timeout 900 python bench.py --steps 2 --warmup 3 --workload exact --no-cpu-baseline > gpurun_out/r02p_bench_exact_n100k.json 2> gpurun_out/r02p_bench_exact_n100k.err
timeout 300 python scripts/small_n_timing.py > gpurun_out/r02p_small_n.log 2>&1
for f in default exact_n20k exact_n100k; do head -c 330 gpurun_out/r02p_bench_$f.json; echo; tail -2 gpurun_out/r02p_bench_$f.err; done; cat gpurun_out/r02p_small_n.log
