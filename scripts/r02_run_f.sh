#!/bin/bash
# v4 write-out with staged column scales / relaxed hand-back: parity, kernel timing, bench, ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_mll.py tests/test_gpu_conditioning.py tests/test_gpu_kernels_ext.py -q -x > gpurun_out/r02f_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02f_tests.log
OZ_BENCH_QUICK=1 timeout 300 python scripts/ozaki_bench.py > gpurun_out/r02f_oz_quick_v4.json 2> gpurun_out/r02f_oz_quick.err
timeout 600 python bench.py --steps 3 --warmup 3 --workload exact > gpurun_out/r02f_bench_exact.json 2> gpurun_out/r02f_bench_exact.err
N=50000
ncu --set full --clock-control none --import-source on -k regex:ozaki_i8_kernel_w4 -s 1 -c 1 -f -o gpurun_out/r02_ozaki_w4 \
    python scripts/prof_mll.py mll $N > gpurun_out/r02f_prof_ozaki.log 2>&1
R=$((N - 2048))
python scripts/parse_ncu.py gpurun_out/r02_ozaki_w4.ncu-rep gpurun_out/r02_ozaki_w4_ncu.json \
    --algorithmic-bytes $(python -c "print($R * ($R + 1) / 2 * 16 + $R * 6 * 1024)") \
    --launch "potrf step 0, trailing update U2: lower-masked ${R}^2, K=1024, 6 planes + equal-plane term, N=$N"
tail -3 gpurun_out/r02f_tests.log; head -c 300 gpurun_out/r02f_bench_exact.json
