#!/bin/bash
# One GPU-box session: parity tests, smoke, bench, launch list and ncu captures of the two dominant kernels.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
if [ "$1" == "ncu" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --scan-tokens 0 > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:encode_topk_kernel -s 9 -c 1 -f -o gpurun_out/prof_encode \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --scan-tokens 0 > gpurun_out/ncu_encode.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:refine_kernel -s 2 -c 1 -f -o gpurun_out/prof_refine \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --scan-tokens 0 > gpurun_out/ncu_refine.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 2 -c 1 -f -o gpurun_out/prof_decode \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --scan-tokens 0 > gpurun_out/ncu_decode.log 2>&1
ls -la gpurun_out
fi
