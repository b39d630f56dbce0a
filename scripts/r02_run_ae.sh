#!/bin/bash
# A/B of the Kzz overlap on the per-rank shard size of the 8-GPU SGPR run (1.25 M rows), one GPU, same box; SVGP check
mkdir -p gpurun_out
for v in 0 1; do
GPB_BENCH_SGPR_N=1250000 GPB_SGPR_OVERLAP_KZZ=$v timeout 300 python bench.py --workload sgpr --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02af_sgpr_1p25m_overlap${v}.json 2> gpurun_out/r02af_sgpr_1p25m_overlap${v}.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02af_sgpr_1p25m_overlap${v}.json")); print("overlap=$v", d["ms_per_step"], d["phases_ms_max_over_ranks"])
PY
done
timeout 300 python bench.py --workload svgp --steps 10 --warmup 3 > gpurun_out/r02af_bench_svgp.json 2> gpurun_out/r02af_bench_svgp.err; head -c 220 gpurun_out/r02af_bench_svgp.json; echo
timeout 600 python -m pytest tests/test_gpu_sgpr.py tests/test_gpu_svgp.py -q 2>&1 | tail -2
