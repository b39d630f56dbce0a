#!/bin/bash
# replicated M x M finish on the int8 pipe (GPB_SGPR_FINISH_INT8): full GPU suite, SVGP / SGPR bench with the switch off / on
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02w_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02w_tests.log
for v in 0 1; do
GPB_SGPR_FINISH_INT8=$v timeout 600 python bench.py --workload svgp --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02w_bench_svgp_finish$v.json 2> gpurun_out/r02w_bench_svgp_finish$v.err
GPB_SGPR_FINISH_INT8=$v timeout 600 python bench.py --workload sgpr --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02w_bench_sgpr_finish$v.json 2> gpurun_out/r02w_bench_sgpr_finish$v.err
done
tail -6 gpurun_out/r02w_tests.log
for f in svgp_finish0 svgp_finish1 sgpr_finish0 sgpr_finish1; do head -c 260 gpurun_out/r02w_bench_$f.json; echo; tail -n 2 gpurun_out/r02w_bench_$f.err; done
