"""Turn what a GPU-box session brought back in gpurun_out/ into the small tracked files under profiles/:

    python tools/ncu_summarise.py rep   gpurun_out/X.ncu-rep  profiles/Y_summary.json   # --set full capture -> key metrics
    python tools/ncu_summarise.py share gpurun_out/X_launches.csv profiles/Y_shares.json  # launch list -> kernel shares

`rep` reads the report with `ncu -i ... --page raw --csv` (works without a GPU) and keeps the metrics B200_PROFILING.md
names (duration, DRAM bytes / throughput, tensor pipe, L2 hit rate, occupancy, registers, shared memory, stall reasons).
"""
import csv
import io
import json
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum(\.per_second)?|dram__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"sm__pipe_tensor.*cycles_active.*pct_of_peak_sustained_(elapsed|active)|sm__inst_executed_pipe_tensor.*|"
    r"sm__cycles_elapsed\.avg\.per_second|sm__cycles_active\.avg|lts__t_sector_hit_rate\.pct|"
    r"lts__t_bytes\.sum(\.per_second)?|l1tex__t_sector_hit_rate\.pct|launch__.*|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
    r"smsp__average_warps?_issue_stalled_.*_per_issue_active\.ratio|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
    r"smsp__inst_executed\.sum|sm__inst_executed_pipe_(fma|alu|lsu|xu|uniform).*pct.*|gpc__cycles_elapsed\.max)$")


def rep(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    res = []
    for r in data:
        d = {"Kernel Name": r[hdr.index("Kernel Name")]}
        for h, u, v in zip(hdr, units, r):
            if KEEP.match(h) and v != "" and not re.search(r"\.(max|min|sum)\.pct_of_peak", h):
                d[h] = [v, u]
        res.append(d)
    json.dump({"source": path.split("/")[-1] + " (ncu --set full --clock-control none --import-source on)",
               "kernels": res}, open(out, "w"), indent=1)
    for d in res:
        t = d.get("gpu__time_duration.sum", ["?", ""])
        print(d["Kernel Name"][:80], t)


def share(path, out, last=0):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = [r for r in csv.DictReader(io.StringIO("".join(lines))) if r.get("Metric Name") == "gpu__time_duration.sum"]
    if last:   # only the last `last` launches of the run (e.g. the timed pass of a tool that warms up first)
        rows = rows[-int(last):]
    per, total, n = {}, 0.0, 0
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
        name = re.sub(r"^void ", "", name)
        a = per.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
        total += us
        n += 1
    tab = sorted(((k, c, us, us / total) for k, (c, us) in per.items()), key=lambda t: -t[2])
    json.dump({"source": path.split("/")[-1] + " (ncu --metrics gpu__time_duration.sum --clock-control none: per-launch "
               "times are cold-cache and serialised -- compare SHARES, not absolutes)", "launches": n,
               "total_us": round(total, 1),
               "kernels": [{"kernel": k, "launches": c, "us": round(us, 1), "share": round(s, 4)} for k, c, us, s in tab]},
              open(out, "w"), indent=1)
    for k, c, us, s in tab[:14]:
        print(f"{s * 100:6.2f}%  {us / 1e3:9.3f} ms  x{c:<4d} {k[:90]}")


if __name__ == "__main__":
    {"rep": rep, "share": share}[sys.argv[1]](*sys.argv[2:])
