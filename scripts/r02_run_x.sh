#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02x_svgp_launches.csv \
    python scripts/prof_svgp.py > gpurun_out/r02x_prof_svgp.log 2>&1
python scripts/summarize_launches.py gpurun_out/r02x_svgp_launches.csv gpurun_out/r02x_svgp_launches.md
gzip -f gpurun_out/r02x_svgp_launches.csv
head -45 gpurun_out/r02x_svgp_launches.md; tail -n 3 gpurun_out/r02x_prof_svgp.log
