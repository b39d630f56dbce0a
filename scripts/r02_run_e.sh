#!/bin/bash
# composable sparse objectives + symmetrisation tests, then launch list and --set full capture of the v4 kernel at N = 50k
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels_ext.py tests/test_gpu_api.py tests/test_gpu_sgpr.py tests/test_gpu_svgp.py -q > gpurun_out/r02e_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02e_tests.log
N=50000
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02e_exact${N}_launches.csv \
    python scripts/prof_mll.py mll $N > gpurun_out/r02e_prof_launches.log 2>&1
gzip -f gpurun_out/r02e_exact${N}_launches.csv
ncu --set full --clock-control none --import-source on -k regex:ozaki_i8_kernel_w4 -s 1 -c 1 -f -o gpurun_out/r02_ozaki_w4 \
    python scripts/prof_mll.py mll $N > gpurun_out/r02e_prof_ozaki.log 2>&1
R=$((N - 2048))
python scripts/parse_ncu.py gpurun_out/r02_ozaki_w4.ncu-rep gpurun_out/r02_ozaki_w4_ncu.json \
    --algorithmic-bytes $(python -c "print($R * ($R + 1) / 2 * 16 + $R * 6 * 1024)") \
    --launch "potrf step 0, trailing update U2: lower-masked ${R}^2, K=1024, 6 planes + equal-plane term, N=$N"
tail -4 gpurun_out/r02e_tests.log; ls -la gpurun_out/r02_ozaki_w4*
