#!/bin/bash
python - <<'PY'
import ctypes
rt = ctypes.CDLL("libcudart.so") if False else None
import torch
p = torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size, "persistingL2CacheMaxSize", getattr(p, "persisting_l2_cache_max_size", None), "accessPolicyMaxWindowSize", getattr(p, "access_policy_max_window_size", None))
PY
for E in t_h0 t_h1 t_h2 t_h3 t_s8 t_9472 t_9472_persist t_9472_persist_h2 t_4736_persist_s4; do
  ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second \
      --clock-control none -k regex:encode_topk_kernel -s 1 -c 1 --csv --log-file gpurun_out/ncu_$E.csv \
      python tools/gpu_probe.py --child $E > gpurun_out/ncu_$E.log 2>&1
  grep -E "dram__|gpu__time|lts__|tensor|per_second" gpurun_out/ncu_$E.csv | awk -F'","' '{printf "%s %s %s | ", "'$E'", $(NF-2), $NF}'; echo
done
