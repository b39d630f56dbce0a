#!/bin/bash
# per-order block size (NB = 2048 from 16,384 rows): full GPU suite, block-size sweep, ncu capture of the dominant launch, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02t_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02t_tests.log
timeout 400 python scripts/nb_sweep.py 2000 5000 10000 16384 20000 30000 50000 > gpurun_out/r02t_nb_default.log 2>&1
if [ -f gpjax_b200/lib/nb4096/libgpjax_b200.so ]; then
  GPB_LIB_PATH=$PWD/gpjax_b200/lib/nb4096/libgpjax_b200.so timeout 400 python scripts/nb_sweep.py 30000 50000 > gpurun_out/r02t_nb4096.log 2>&1
fi
GPB_LIB_PATH=$PWD/gpjax_b200/lib/nb2048/libgpjax_b200.so timeout 400 python scripts/nb_sweep.py 2000 5000 16384 > gpurun_out/r02t_nb2048.log 2>&1
N=50000
NB=$(python -c "from gpjax_b200._lib import lib; print(lib().gpb_block_size_for($N))")
R=$((N - 2 * NB))
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ozaki_i8_kernel_w4 -s 3 -c 1 -f -o gpurun_out/r02t_ozaki_w4 \
    python scripts/prof_mll.py mll $N > gpurun_out/r02t_prof_ozaki.log 2>&1
python scripts/parse_ncu.py gpurun_out/r02t_ozaki_w4.ncu-rep gpurun_out/r02t_ozaki_w4_ncu.json \
    --algorithmic-bytes $(python -c "print($R * ($R + 1) / 2 * 16 + $R * 6 * $NB)") \
    --launch "potrf step 0, trailing update U2: lower-masked ${R}^2, K=$NB, 6 planes + equal-plane term, N=$N" > gpurun_out/r02t_parse.log 2>&1
rm -f gpurun_out/r02t_ozaki_w4.ncu-rep
timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/r02t_bench_default.json 2> gpurun_out/r02t_bench_default.err
tail -4 gpurun_out/r02t_tests.log; cat gpurun_out/r02t_nb_default.log gpurun_out/r02t_nb4096.log gpurun_out/r02t_nb2048.log; tail -n 5 gpurun_out/r02t_parse.log; head -c 400 gpurun_out/r02t_bench_default.json; echo; tail -n 2 gpurun_out/r02t_bench_default.err
