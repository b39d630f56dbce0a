#!/bin/bash
mkdir -p gpurun_out
timeout 58 python -m pytest tests/test_gpu_sgpr.py tests/test_gpu_svgp.py tests/test_gpu_kernels_ext.py tests/test_gpu_ozaki.py tests/test_gpu_boundary.py tests/test_gpu_conditioning.py tests/test_gpu_fit_graph.py -q -n 3 -p no:cacheprovider > gpurun_out/r02h_tests2.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02h_tests2.log
tail -8 gpurun_out/r02h_tests2.log
