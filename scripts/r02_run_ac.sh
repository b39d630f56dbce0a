#!/bin/bash
# Kzz chain next to the first streamed blocks (GPB_SGPR_OVERLAP_KZZ): parity tests, SVGP / SGPR bench with the switch off / on
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sgpr.py tests/test_gpu_svgp.py tests/test_gpu_kernels_ext.py tests/test_gpu_api.py -q > gpurun_out/r02ac_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02ac_tests.log
for v in 0 1; do
GPB_SGPR_OVERLAP_KZZ=$v timeout 300 python bench.py --workload svgp --steps 10 --warmup 3 > gpurun_out/r02ac_bench_svgp_overlap$v.json 2> gpurun_out/r02ac_bench_svgp_overlap$v.err
done
timeout 600 python bench.py --workload sgpr --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02ac_bench_sgpr_n1.json 2> gpurun_out/r02ac_bench_sgpr_n1.err
tail -3 gpurun_out/r02ac_tests.log; for f in svgp_overlap0 svgp_overlap1 sgpr_n1; do head -c 250 gpurun_out/r02ac_bench_$f.json; echo; tail -n 2 gpurun_out/r02ac_bench_$f.err; done
