#!/bin/bash
# Round-2 ncu evidence (run under gpurun on ONE B200; outputs under gpurun_out/, summaries are copied to profiles/ by hand).
#   1. launch list of one exact-GP evaluation at N = 50,000 (default path: int8 digit planes, CTA-pair kernel)
#   2. --set full of the dominant kernel (second ozaki_i8_kernel_w4 launch = the 47,952^2 lower update of potrf step 0)
#   3. --set full of the bandwidth-class kernels the north-star names: gram_kernel, mll_bwd_kernel, gemv_{n,t}_kernel
# gpurun only brings back <= 64 MiB: the big reports are reduced to JSON on the box (scripts/parse_ncu.py) and deleted.
set -x
N=${1:-50000}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_exact${N}_launches.csv \
    python scripts/prof_mll.py mll $N > gpurun_out/r02_prof_launches.log 2>&1
gzip -f gpurun_out/r02_exact${N}_launches.csv
ncu --set full --clock-control none --import-source on -k regex:ozaki_i8_kernel_w4 -s 1 -c 1 -f -o gpurun_out/r02_ozaki_w4 \
    python scripts/prof_mll.py mll $N > gpurun_out/r02_prof_ozaki.log 2>&1
# launch #2 of the kernel = U2 of potrf step 0: lower-masked (N - 2048)^2 update, K = 1024, 6 planes (guard) -> algorithmic bytes
R=$((N - 2048))
python scripts/parse_ncu.py gpurun_out/r02_ozaki_w4.ncu-rep gpurun_out/r02_ozaki_w4_ncu.json \
    --algorithmic-bytes $(python -c "print($R * ($R + 1) / 2 * 16 + $R * 6 * 1024)") \
    --launch "potrf step 0, trailing update U2: lower-masked ${R}^2, K=1024, 6 planes + equal-plane term, N=$N"
ncu --set full --clock-control none -k regex:'gram_kernel|mll_bwd_kernel' -c 2 -f -o gpurun_out/r02_hbm_kernels \
    python scripts/prof_mll.py mll $N > gpurun_out/r02_prof_hbm.log 2>&1
python scripts/parse_ncu.py gpurun_out/r02_hbm_kernels.ncu-rep gpurun_out/r02_hbm_kernels_ncu.json
rm -f gpurun_out/r02_hbm_kernels.ncu-rep
ncu --set full --clock-control none -k regex:'gemv_' -s 6 -c 4 -f -o gpurun_out/r02_gemv \
    python scripts/prof_mll.py mll $N > gpurun_out/r02_prof_gemv.log 2>&1
python scripts/parse_ncu.py gpurun_out/r02_gemv.ncu-rep gpurun_out/r02_gemv_ncu.json
rm -f gpurun_out/r02_gemv.ncu-rep
du -sh gpurun_out
