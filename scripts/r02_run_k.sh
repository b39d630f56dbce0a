#!/bin/bash
# launch lists of one SGPR evaluation at N = 1M (16 blocks), fused and unfused digit extraction
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02k_sgpr1m_fused.csv python scripts/prof_sgpr.py 1000000 > gpurun_out/r02k_fused.log 2>&1
GPB_SGPR_FUSED=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02k_sgpr1m_unfused.csv python scripts/prof_sgpr.py 1000000 > gpurun_out/r02k_unfused.log 2>&1
python scripts/summarize_launches.py gpurun_out/r02k_sgpr1m_fused.csv gpurun_out/r02k_sgpr1m_fused.md > /dev/null
python scripts/summarize_launches.py gpurun_out/r02k_sgpr1m_unfused.csv gpurun_out/r02k_sgpr1m_unfused.md > /dev/null
gzip -f gpurun_out/r02k_sgpr1m_fused.csv gpurun_out/r02k_sgpr1m_unfused.csv
head -16 gpurun_out/r02k_sgpr1m_fused.md; head -18 gpurun_out/r02k_sgpr1m_unfused.md
