#!/bin/bash
# where the non-int8 time goes: launch lists at N = 20k / 50k with the 2048 block, source-level capture of potrf_leaf_kernel
mkdir -p gpurun_out
for N in 20000 50000; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02v_exact${N}_launches.csv \
    python scripts/prof_mll.py mll $N > gpurun_out/r02v_prof_launches_$N.log 2>&1
python scripts/summarize_launches.py gpurun_out/r02v_exact${N}_launches.csv gpurun_out/r02v_exact${N}_launches.md
gzip -f gpurun_out/r02v_exact${N}_launches.csv
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:potrf_leaf_kernel -s 2 -c 1 -f -o gpurun_out/r02v_leaf \
    python scripts/prof_mll.py mll 2000 > gpurun_out/r02v_prof_leaf.log 2>&1
ls -la gpurun_out/r02v_leaf.ncu-rep
head -14 gpurun_out/r02v_exact20000_launches.md; head -14 gpurun_out/r02v_exact50000_launches.md
