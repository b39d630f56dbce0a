#!/bin/bash
# block size of the blocked algorithms with the int8 path: NB = 1024 (default build) vs NB = 2048 (gpjax_b200/lib/nb2048)
mkdir -p gpurun_out
timeout 400 python scripts/nb_sweep.py 10000 20000 50000 > gpurun_out/r02s_nb1024.log 2>&1
GPB_LIB_PATH=$PWD/gpjax_b200/lib/nb2048/libgpjax_b200.so timeout 400 python scripts/nb_sweep.py 10000 20000 50000 > gpurun_out/r02s_nb2048.log 2>&1
GPB_LIB_PATH=$PWD/gpjax_b200/lib/nb2048/libgpjax_b200.so timeout 300 python scripts/cond_sweep_fixture.py > gpurun_out/r02s_sweep_nb2048.jsonl 2> gpurun_out/r02s_sweep_nb2048.err
cat gpurun_out/r02s_nb1024.log gpurun_out/r02s_nb2048.log; tail -n 3 gpurun_out/r02s_sweep_nb2048.err; wc -l gpurun_out/r02s_sweep_nb2048.jsonl
