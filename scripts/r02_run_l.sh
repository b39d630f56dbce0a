#!/bin/bash
# fused Gram -> digit planes, third version: parity + bench + launch list
mkdir -p gpurun_out
GPB_TEST_SKIP_CONFIG3=1 timeout 1200 python -m pytest tests/test_gpu_sgpr.py tests/test_gpu_svgp.py -q > gpurun_out/r02l_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02l_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --workload sgpr --no-cpu-baseline > gpurun_out/r02l_bench_sgpr_fused.json 2> gpurun_out/r02l_bench_sgpr_fused.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02l_sgpr1m_fused.csv python scripts/prof_sgpr.py 1000000 raw > gpurun_out/r02l_fused.log 2>&1
python scripts/summarize_launches.py gpurun_out/r02l_sgpr1m_fused.csv gpurun_out/r02l_sgpr1m_fused.md > /dev/null
gzip -f gpurun_out/r02l_sgpr1m_fused.csv
tail -3 gpurun_out/r02l_tests.log; head -c 250 gpurun_out/r02l_bench_sgpr_fused.json; echo; head -18 gpurun_out/r02l_sgpr1m_fused.md
