#!/bin/bash
# fit(cuda_graph=True): parity tests; ncu launch lists of one MLL value+grad at N=1000 / 2000 (where the small-N step spends its time)
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_fit_graph.py -x -q > gpurun_out/r02g_fit_graph_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02g_fit_graph_tests.log
tail -15 gpurun_out/r02g_fit_graph_tests.log
for n in 1000 2000; do
  timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02g_mll${n}_launches.csv python scripts/prof_mll.py mll $n > gpurun_out/r02g_prof_${n}.log 2>&1
  python scripts/summarize_launches.py gpurun_out/r02g_mll${n}_launches.csv gpurun_out/r02g_mll${n}_launches.md >> gpurun_out/r02g_prof_${n}.log 2>&1
  head -30 gpurun_out/r02g_mll${n}_launches.md
done
