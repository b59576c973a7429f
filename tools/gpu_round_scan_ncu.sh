#!/bin/bash
# ncu evidence for the kernels of the sharded scan's per-chunk chain, on ONE GPU through tools/scan_rank_emul.py
# (one rank of the 8-way scan with replayed exchanges).  Usage: tools/gpu_round_scan_ncu.sh <tag>
TAG=${1:-r02}
E="python tools/scan_rank_emul.py --stages 5 --prio low --chunks 4"
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_scan_launches.csv \
    $E > gpurun_out/${TAG}_scan_under_ncu.log 2>&1
for K in refine_scan_warp_kernel scan_select_bounds_kernel gathered_bounds_kernel scan_pool_g_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 5 -c 1 -f -o gpurun_out/${TAG}_prof_$K \
      $E > gpurun_out/${TAG}_ncu_$K.log 2>&1
done
ls -la gpurun_out | grep ${TAG}
