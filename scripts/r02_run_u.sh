#!/bin/bash
# tuning sweeps around the per-order block size: SMs left to the look-ahead chain, NB_LARGE = 4096, the switch-over order
mkdir -p gpurun_out
out=gpurun_out/r02u_sweeps.log; : > $out
for f in 8 16 24 32; do echo "## GPB_LOOKAHEAD_FREE_SMS=$f" >> $out; GPB_LOOKAHEAD_FREE_SMS=$f timeout 200 python scripts/nb_sweep.py 30000 50000 >> $out 2>&1; done
echo "## GPB_NB_LARGE=4096" >> $out; GPB_NB_LARGE=4096 timeout 200 python scripts/nb_sweep.py 50000 >> $out 2>&1
echo "## GPB_NB_LARGE=4096 GPB_LOOKAHEAD_FREE_SMS=16" >> $out; GPB_NB_LARGE=4096 GPB_LOOKAHEAD_FREE_SMS=16 timeout 200 python scripts/nb_sweep.py 50000 >> $out 2>&1
echo "## GPB_NB_LARGE_MIN_ROWS=8192 (2048 from 8192 rows)" >> $out; GPB_NB_LARGE_MIN_ROWS=8192 timeout 200 python scripts/nb_sweep.py 8192 10000 12288 16384 >> $out 2>&1
echo "## GPB_NB_LARGE_MIN_ROWS=1000000 (1024 everywhere)" >> $out; GPB_NB_LARGE_MIN_ROWS=1000000 timeout 200 python scripts/nb_sweep.py 8192 10000 12288 16384 >> $out 2>&1
cat $out
