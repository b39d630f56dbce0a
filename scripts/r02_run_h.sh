#!/bin/bash
# full GPU suite + the driver's bench commands (both arms) on the final build
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/r02h_tests_full.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02h_tests_full.log
timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/r02h_bench_default.json 2> gpurun_out/r02h_bench_default.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 3 > gpurun_out/r02h_bench_reference.json 2> gpurun_out/r02h_bench_reference.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02h_smoke.log 2>&1
tail -3 gpurun_out/r02h_tests_full.log; tail -2 gpurun_out/r02h_smoke.log; head -c 300 gpurun_out/r02h_bench_default.json; echo; head -c 400 gpurun_out/r02h_bench_reference.json
