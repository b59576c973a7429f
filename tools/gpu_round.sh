#!/bin/bash
# One GPU-box session (1 GPU): parity tests, smoke, bench (both arms), launch list and ncu captures of the three
# dominant kernels.  Usage: tools/gpu_round.sh <tag> [ncu]   -> files gpurun_out/<tag>_*
TAG=${1:-r02}
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > gpurun_out/${TAG}_smi.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench_ref.json
python bench.py --steps 5 --warmup 3 --alt-fp16-decode --alt-mode4 > gpurun_out/${TAG}_bench.json 2>> gpurun_out/${TAG}_bench.err; cut -c1-1500 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
if [ "$2" == "ncu" ]; then
B="python bench.py --no-cpu-baseline --scan-c3-tokens 0 --scan-c4-tokens 0 --no-c5 --no-alt-modes"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    $B --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:encode_topk_kernel -s 9 -c 1 -f -o gpurun_out/${TAG}_prof_encode \
    $B --steps 1 --warmup 1 > gpurun_out/${TAG}_ncu_encode.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:refine_kernel -s 9 -c 1 -f -o gpurun_out/${TAG}_prof_refine \
    $B --steps 1 --warmup 1 > gpurun_out/${TAG}_ncu_refine.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 9 -c 1 -f -o gpurun_out/${TAG}_prof_decode \
    $B --steps 1 --warmup 1 > gpurun_out/${TAG}_ncu_decode.log 2>&1
ls -la gpurun_out
fi
