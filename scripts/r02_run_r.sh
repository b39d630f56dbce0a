#!/bin/bash
# int8 panel products: conditioning-sweep errors with the switch off / on (oracle fixture), launch list of the N = 50k step
mkdir -p gpurun_out
GPB_OZ_PANELS=0 timeout 300 python scripts/cond_sweep_fixture.py > gpurun_out/r02r_sweep_panels0.jsonl 2> gpurun_out/r02r_sweep0.err
GPB_OZ_PANELS=1 timeout 300 python scripts/cond_sweep_fixture.py > gpurun_out/r02r_sweep_panels1.jsonl 2> gpurun_out/r02r_sweep1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02r_exact50000_launches.csv \
    python scripts/prof_mll.py mll 50000 > gpurun_out/r02r_prof_launches.log 2>&1
python scripts/summarize_launches.py gpurun_out/r02r_exact50000_launches.csv gpurun_out/r02r_exact50000_launches.md
gzip -f gpurun_out/r02r_exact50000_launches.csv
tail -3 gpurun_out/r02r_sweep0.err gpurun_out/r02r_sweep1.err; wc -l gpurun_out/r02r_sweep_panels*.jsonl; head -12 gpurun_out/r02r_exact50000_launches.md
