#!/bin/bash
python - <<'PY'
import sys
sys.path.insert(0, "multimodal-sae_b200")
import torch
torch.zeros(1, device="cuda")
from saeb200 import _capi
L = _capi.lib()
for n in (b"num_sms", b"l2_bytes", b"persisting_l2_max_bytes", b"access_policy_max_window_bytes"):
    print(n.decode(), L.saeb_query(n))
PY
for E in t_p0 t_p1 t_p2 t_p3; do
  ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second \
      --clock-control none -k regex:encode_topk_kernel -s 1 -c 1 --csv --log-file gpurun_out/ncu_$E.csv \
      python tools/gpu_probe.py --child $E > gpurun_out/ncu_$E.log 2>&1
  grep -E "dram__|gpu__time|lts__|tensor|per_second" gpurun_out/ncu_$E.csv | awk -F'","' '{printf "%s %s %s | ", "'$E'", $(NF-2), $NF}'; echo
done
python tools/gpu_probe.py t_p0 t_p1 t_p2 t_p3 2>&1 | grep RESULT | cut -c1-330
