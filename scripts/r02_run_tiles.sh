#!/bin/bash
mkdir -p gpurun_out
( GPB_GEMM_SMALL_TILES=0 timeout 30 python scripts/small_tiles_ab.py; GPB_GEMM_SMALL_TILES=1 timeout 30 python scripts/small_tiles_ab.py ) > gpurun_out/r02h_small_tiles_ab.log 2>&1
cat gpurun_out/r02h_small_tiles_ab.log
timeout 75 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_mll.py tests/test_gpu_api.py -x -q > gpurun_out/r02h_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02h_tests.log
tail -6 gpurun_out/r02h_tests.log
