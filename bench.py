#!/usr/bin/env python
"""Benchmark of the SAE hot path (BASELINE.json metric: SAE tokens/sec at d=4096, width=131k).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step is one pass of the hot path -- fused encode + TopK, sparse decode, FVU (reference Sae.forward,
sae/sae.py:193-247) -- over one batch of 65 536 synthetic bf16 tokens (BASELINE.json configs[1]).  With N GPUs every
rank runs the same step on its own 65 536 tokens with the SAE replicated (tokens are independent: no data-path
collective, weak scaling); `value` is the whole-job tokens/s, timed on the device with CUDA events, barrier +
synchronize on both sides, max over ranks.  The JSON line also carries
  e2e          the same metric through the reference-facing `Sae` objects with HOST buffers (pinned x in, TopK + FVU
               back), copies inside the timed region;
  roofline     the fused encode kernel against the measured bf16 tensor peak (MEASURED_PEAKS.json);
  cpu_baseline the oracle's CPU forward (the reference's PyTorch ops) on a bounded sample, rank 0, N=1 only;
  scan         feature-sharded top-activation scan (north star: features split across ranks, one all-gather of
               per-shard top lists at the end) on a bounded token count.
`--impl reference` times the reference's own CPU implementation of the path (the oracle port of its PyTorch ops; a
Python reference cannot be pre-built into oracle/_ref) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "multimodal-sae_b200"))

D_IN, WIDTH, K = 4096, 131072, 64
TOKENS = 65536
FLOPS_PER_TOKEN = 2.0 * D_IN * WIDTH + 2.0 * K * D_IN  # algorithmic, independent of MMA passes
ENC_FLOPS_PER_TOKEN = 2.0 * D_IN * WIDTH
METRIC = "SAE tokens/sec (d=4096, width=131k, k=64, encode+TopK+decode)"


PRECISION = {
    3: "one fp16 tensor-core pass (activations exact after a power-of-two row scale, W_enc rounded to fp16, fp32 "
       "accumulate) + exact fp32 re-evaluation of every candidate inside the rigorous rounding bound; fp32 W_dec",
    4: "one fp16 tensor-core pass + refinement by residual correction (x . fp16 residual plane of W_enc added to the "
       "tensor-core value for every candidate inside the rounding bound; values to ~1e-6 relative); fp32 W_dec",
    2: "bf16 activations (exact) x bf16 hi+lo W_enc planes (two tensor-core passes), fp32 accumulate; fp32 W_dec",
    1: "single bf16 pass (NOT parity grade; diagnostic only)",
}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], tflops_burst=p["bf16_tflops"],
                    tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], None, set(), []
        for ts, line in self.lines:
            if ts < t0 or ts > t1 + 0.1:
                continue
            f = [s.strip() for s in line.split(",")]
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
                power.append(float(f[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


# ---------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port of the reference's PyTorch forward on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_forward_tokens_per_s(sample_tokens: int, batch: int, repeats: int, seed: int = 1234):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import sae_oracle as O

    threads = torch.get_num_threads()
    p = O.init_params(D_IN, WIDTH, K, seed)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(sample_tokens, D_IN, generator=g).to(torch.bfloat16)
    best = None
    with torch.no_grad():
        O.forward(p, x[:batch])  # warm-up
        for _ in range(repeats):
            t = time.perf_counter()
            for b0 in range(0, sample_tokens, batch):
                O.forward(p, x[b0:b0 + batch])
            dt = time.perf_counter() - t
            best = dt if best is None else min(best, dt)
    return sample_tokens / best, threads, best


def cpu_cache_chain_tokens_per_s(sample_tokens: int, batch: int, seed: int = 1234):
    """the reference's cache-path chain (features/cache.py:206-218 + :73-92: pre_acts -> topk -> zeros_like + scatter_
    -> nonzero(|x| > 1e-5)) through the oracle port, on the host cores: the CPU baseline of the scan's per-token work"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import sae_oracle as O

    p = O.init_params(D_IN, WIDTH, K, seed)
    x = torch.randn(sample_tokens, D_IN, generator=torch.Generator().manual_seed(seed + 2)).to(torch.bfloat16)
    with torch.no_grad():
        O.get_nonzeros(O.topk_masked_latents(p, x[:64].view(1, 64, D_IN)))  # warm-up
        t = time.perf_counter()
        nnz = 0
        for b0 in range(0, sample_tokens, batch):
            loc, _ = O.get_nonzeros(O.topk_masked_latents(p, x[b0:b0 + batch].view(1, -1, D_IN)))
            nnz += loc.shape[0]
        dt = time.perf_counter() - t
    return sample_tokens / dt, torch.get_num_threads(), dt, nnz


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample, batch = 1024, 512
    warm = max(args.warmup, 0)
    tok_s, threads, _ = cpu_forward_tokens_per_s(sample, batch, repeats=max(1, min(args.steps, 3)))
    line = {
        "impl": "reference", "metric": METRIC, "value": tok_s, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": warm, "ms_per_step": 1e3 * sample / tok_s, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2: d_model=4096 width=131072 k=64 SAE forward (bounded sample)",
                   "global_batch_tokens": sample, "parallelism": "cpu"},
        "cpu_baseline": {"value": tok_s, "unit": "tokens/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} tokens in {batch}-token batches per step, best of {max(1, min(args.steps, 3))}"},
        "e2e": {"value": tok_s, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    from saeb200 import _capi, dist as sdist, engine, pipeline, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _capi.lib()
    _capi.check(L.saeb_set_option(b"profile", 1), "set_option")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sae = synth.make_sae(D_IN, WIDTH, K, dev, seed=1234)
    sae.encoder_planes = args.planes
    enc = sae.packed_encoder()
    x = synth.make_activations(TOKENS, D_IN, dev, seed=1 + rank)
    acts = torch.empty((TOKENS, K), dtype=torch.float32, device=dev)
    idx = torch.empty((TOKENS, K), dtype=torch.int64, device=dev)
    sae_out = torch.empty((TOKENS, D_IN), dtype=torch.float32, device=dev)
    sq_err = torch.zeros((), dtype=torch.float64, device=dev)

    W_dec_used = sae.W_dec.data if args.decode_dtype == "fp32" else sae.W_dec.data.to(torch.float16)
    ov = None
    if args.planes in (3, 4) and not args.no_overlap:
        from saeb200.overlap import OverlappedForward

        ov = OverlappedForward(enc, W_dec_used, sae.b_dec.data, K, chunk=args.chunk)

    def step():
        sq_err.zero_()
        if ov is not None:   # GEMM of chunk c+1 on one stream, refinement + decode of chunk c on another
            ov.run(x, acts, idx, sae_out, sq_err)
        else:
            engine.encode_topk(x, enc, K, out_vals=acts, out_idx=idx)
            engine.decode(idx, acts, W_dec_used, sae.b_dec.data, x=x, sq_err=sq_err, out=sae_out)
        return (sq_err / engine.total_variance(x)).to(torch.float32)

    # ---- device-resident throughput (`value`)
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = _capi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        fvu = step()
    e1.record()
    barrier()
    t_wall1 = time.time()
    launches = _capi.launch_count() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    ms_per_step = ms_total / args.steps
    value = world * TOKENS * args.steps / (ms_total * 1e-3)
    fvu_val = float(fvu.item())

    # ---- dominant kernel, live CUDA events on the launching stream (library-side bracket of the main kernel)
    k_ms = []
    for _ in range(max(3, min(args.steps, 5))):
        engine.encode_topk(x, enc, K, out_vals=acts, out_idx=idx)
        k_ms.append(float(L.saeb_profile_last_encode_ms()))
    k_avg = sum(k_ms) / len(k_ms)
    peaks = load_peaks()
    achieved = ENC_FLOPS_PER_TOKEN * TOKENS / (k_avg * 1e-3) / 1e12
    traffic = None
    summ = os.path.join(ROOT, "profiles", "encode_kernel_traffic.json")
    if os.path.exists(summ):
        try:
            traffic = json.load(open(summ)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    wave_tokens = 9472
    n_launch = (TOKENS + wave_tokens - 1) // wave_tokens
    roofline = {"bound": "tensor", "kernel": "encode_topk_kernel (tcgen05 GEMM + fused TopK)", "achieved": achieved,
                "launches_per_step": n_launch, "tokens_per_launch": wave_tokens,
                "algorithmic_flops_per_launch": ENC_FLOPS_PER_TOKEN * wave_tokens,
                "avg_launch_ms": k_avg * wave_tokens / TOKENS,
                "peak": peaks["tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops_sustained"],
                "peak_burst": peaks["tflops_burst"], "frac_of_burst_peak": achieved / peaks["tflops_burst"],
                "traffic": traffic, "peak_source": f"{peaks['source']} bf16 sustained (cuBLAS, MEASURED_PEAKS.json)",
                "kernel_ms": k_avg, "mma_passes": 1 if enc.planes >= 3 else enc.planes,
                "path_frac": (value / world) * FLOPS_PER_TOKEN / 1e12 / peaks["tflops_sustained"],
                "hbm_frac": (value / world) * (2 * D_IN + K * 12 + K * 4 * D_IN + 4 * D_IN
                                               + 2.0 * D_IN * WIDTH * (1 if enc.planes >= 3 else enc.planes) / TOKENS)
                / 1e9 / peaks["hbm_gbs"]}

    # ---- opt-in (--alt-fp16-decode): the same step with an fp16 copy of W_dec: half the decode gather bytes, row error
    # ~2e-4 -- inside the 1e-3 bar of BASELINE.json but not the parity-grade default, so it is reported NEXT to
    # `value`, never instead of it
    alt = None
    if world == 1 and ov is not None and args.decode_dtype == "fp32" and args.alt_fp16_decode:
        try:
            from saeb200.overlap import OverlappedForward

            W16 = sae.W_dec.data.to(torch.float16)
            ov16 = OverlappedForward(enc, W16, sae.b_dec.data, K, chunk=args.chunk)
            out16 = torch.empty_like(sae_out)
            acts16, idx16, sq16 = torch.empty_like(acts), torch.empty_like(idx), torch.zeros_like(sq_err)
            for _ in range(2):
                ov16.run(x, acts16, idx16, out16, sq16)
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(args.steps):
                sq16.zero_()
                ov16.run(x, acts16, idx16, out16, sq16)
                fvu16 = (sq16 / engine.total_variance(x)).to(torch.float32)
            a1.record()
            torch.cuda.synchronize()
            ms16 = a0.elapsed_time(a1) / args.steps
            rows = slice(0, 8192)
            rel = ((out16[rows] - sae_out[rows]).norm(dim=1) / sae_out[rows].norm(dim=1)).max()
            alt = {"decode_dtype": "fp16 copy of W_dec (fp32 accumulate)", "value": TOKENS / (ms16 * 1e-3),
                   "unit": "tokens/s", "ms_per_step": ms16,
                   "path_frac": TOKENS / (ms16 * 1e-3) * FLOPS_PER_TOKEN / 1e12 / peaks["tflops_sustained"],
                   "max_rel_row_err_vs_fp32_decode": float(rel.item()), "fvu": float(fvu16.item()),
                   "topk_identical": bool(torch.equal(idx16, idx))}
            del ov16, W16, out16, acts16, idx16
        except Exception as exc:   # a diagnostic line must never take the benchmark down
            alt = {"error": repr(exc)[:300]}

    # ---- opt-in (--alt-mode4): the same step with the residual-correction refinement (packed mode 4) and, on top of
    # it, the fp16 W_dec copy; both inside the 1e-3 bar, reported next to `value`
    alt4 = None
    if world == 1 and args.alt_mode4 and args.planes == 3 and not args.no_overlap:
        try:
            from saeb200.overlap import OverlappedForward

            enc4 = sae.packed_encoder(4)
            out4 = torch.empty_like(sae_out)
            acts4, idx4, sq4 = torch.empty_like(acts), torch.empty_like(idx), torch.zeros_like(sq_err)
            alt4 = {}
            for tag, Wd in (("fp32_decode", sae.W_dec.data), ("fp16_decode", sae.W_dec.data.to(torch.float16))):
                ov4 = OverlappedForward(enc4, Wd, sae.b_dec.data, K, chunk=args.chunk)
                for _ in range(2):
                    ov4.run(x, acts4, idx4, out4, sq4)
                torch.cuda.synchronize()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                for _ in range(args.steps):
                    sq4.zero_()
                    ov4.run(x, acts4, idx4, out4, sq4)
                    fvu4 = (sq4 / engine.total_variance(x)).to(torch.float32)
                a1.record()
                torch.cuda.synchronize()
                ms4 = a0.elapsed_time(a1) / args.steps
                same = (torch.sort(idx4, 1).values == torch.sort(idx, 1).values).all(1)
                v4 = torch.gather(acts4, 1, torch.argsort(idx4, 1))[same]
                v3 = torch.gather(acts, 1, torch.argsort(idx, 1))[same]
                rows = slice(0, 8192)
                alt4[tag] = {"value": TOKENS / (ms4 * 1e-3), "unit": "tokens/s", "ms_per_step": ms4,
                             "path_frac": TOKENS / (ms4 * 1e-3) * FLOPS_PER_TOKEN / 1e12 / peaks["tflops_sustained"],
                             "rows_with_different_topk_set_vs_mode3": int((~same).sum().item()),
                             "max_rel_value_diff_vs_mode3": float(((v4 - v3).abs() / v3.abs()).max().item()),
                             "max_rel_row_err_vs_mode3_fp32_decode":
                                 float(((out4[rows] - sae_out[rows]).norm(dim=1) / sae_out[rows].norm(dim=1)).max().item()),
                             "fvu": float(fvu4.item())}
                del ov4
        except Exception as exc:
            alt4 = {"error": repr(exc)[:300]}

    # ---- end to end through the reference-facing objects with host buffers (`e2e`)
    x_host = synth.make_activations(TOKENS, D_IN, dev, seed=1 + rank, pinned_host=True)
    acts_host = torch.empty((TOKENS, K), dtype=torch.float32, pin_memory=True)
    idx_host = torch.empty((TOKENS, K), dtype=torch.int64, pin_memory=True)
    hf = pipeline.HostForward(sae, TOKENS, chunk=args.chunk)
    for _ in range(max(1, min(args.warmup, 2))):
        hf.run(x_host, acts_host, idx_host)
    barrier()
    e0.record()
    for _ in range(args.steps):
        fvu_host = hf.run(x_host, acts_host, idx_host)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e = {"value": world * TOKENS * args.steps / (e2e_ms * 1e-3), "unit": "tokens/s",
           "h2d_bytes_per_step": hf.h2d_bytes * world, "d2h_bytes_per_step": hf.d2h_bytes * world,
           "api": "sae_auto_interp.sae.Sae + saeb200.pipeline.HostForward (pinned x in; TopK acts/indices + FVU out)",
           "fvu": float(fvu_host.item())}

    # ---- feature-sharded top-activation scan (bounded token count; one all-gather of top lists at the end)
    scan = None
    if args.scan_tokens > 0:
        ctx_len, n_top = 64, args.scan_top
        lo, hi = sdist.shard_range(WIDTH, world, rank)
        ops = sdist.EngineOps(sae.encoder.weight.data[lo:hi], sae.encoder.bias.data[lo:hi], sae.b_dec.data, lo, hi,
                              n_top, ctx_len, dev, planes=args.planes)
        if args.scan_exchange:
            ops.exchange = args.scan_exchange
        chunk = ops.chunk_tokens(world)  # four single-wave GEMM launches per exchange round
        xs = synth.make_activations(args.scan_tokens, D_IN, dev, seed=99)  # same tokens on every rank

        def chunks():
            for t0 in range(0, args.scan_tokens, chunk):
                yield xs[t0:t0 + chunk]

        def warm_chunks():
            for t0 in range(0, min(args.scan_tokens, 4 * chunk), chunk):
                yield xs[t0:t0 + chunk]

        sdist.sharded_scan(warm_chunks(), ops, K, ctx_len, WIDTH)  # warm-up (NCCL channels, scratch, lists)
        ops.scan = engine.TopActivationScan(lo, hi, n_top, ctx_len, dev)
        barrier()
        e0.record()
        phases = {}
        res = sdist.sharded_scan(chunks(), ops, K, ctx_len, WIDTH, phase_times=phases if args.scan_phases else None)
        e1.record()
        barrier()
        sms = max_over_ranks(e0.elapsed_time(e1))
        scan = {"tokens": args.scan_tokens, "features": WIDTH, "n_top": n_top, "ctx_len": ctx_len, "ms": sms,
                "tokens_per_s": args.scan_tokens / (sms * 1e-3), "sharding": f"features/{world}", "exact_topk_mask": True,
                "chunk_tokens": chunk,
                "exchange": None if world == 1 else
                            {"nccl": "2 NCCL all-gathers per chunk", "push": "2 saeb_push_gather kernels per chunk (" +
                             (ops._push.transport if ops._push is not None else "not used") + ")"}[ops.exchange],
                "exchange1_columns": sdist.bounds_width(K, min(K, hi - lo), world) if world > 1 else None,
                "filled_features": int((res.top_win[:, 0] >= 0).sum().item()),
                "schedule": "sequential (phase timing)" if args.scan_phases else
                            {"lookahead": "one-chunk lookahead, both all-gathers asynchronous behind the next "
                                          "chunk's GEMM",
                             "streams": "two streams: GEMM of chunk c+1 overlaps exchange/refine/list update of "
                                        "chunk c"}[sdist.scan_schedule(world)],
                "phase_ms_rank0": {k_: round(v_, 2) for k_, v_ in phases.items()} or None}

    # ---- opt-in cross-check (--scan-crosscheck, N > 1): the token-parallel form of the same scan (full SAE on every
    # rank, tokens split, no per-chunk exchange, one all-gather + merge at the end) must give the same lists
    if scan is not None and world > 1 and args.scan_crosscheck:
        ops_tp = sdist.EngineOps(sae.encoder.weight.data, sae.encoder.bias.data, sae.b_dec.data, 0, WIDTH, n_top,
                                 ctx_len, dev, planes=args.planes)
        w_lo, w_hi = sdist.token_slice(args.scan_tokens // ctx_len, world, rank)
        x_tp = xs[w_lo * ctx_len:w_hi * ctx_len]
        chunk_tp = ops_tp.chunk_tokens(1)

        def chunks_tp():
            for t0 in range(0, x_tp.shape[0], chunk_tp):
                yield x_tp[t0:t0 + chunk_tp]

        barrier()
        e0.record()
        res_tp = sdist.token_parallel_scan(chunks_tp(), ops_tp, K, ctx_len, WIDTH, w_lo, n_top=n_top)
        e1.record()
        barrier()
        tp_ms = max_over_ranks(e0.elapsed_time(e1))
        same_w = bool(torch.equal(res_tp.top_win, res.top_win))
        rel = ((res_tp.top_vals - res.top_vals).abs() / res.top_vals.abs().clamp_min(1e-30)).max()
        scan["crosscheck_token_parallel"] = {"ms": tp_ms, "tokens_per_s": args.scan_tokens / (tp_ms * 1e-3),
                                             "same_windows": same_w, "max_rel_score_diff": float(rel.item())}
        del ops_tp

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        tok_s, threads, secs = cpu_forward_tokens_per_s(2048, 512, repeats=2)
        cpu = {"value": tok_s, "unit": "tokens/s", "cores": threads, "kind": "port",
               "sample": f"oracle port of reference Sae.forward (PyTorch CPU fp32), 2048 tokens in 512-token batches, "
                         f"best of 2 ({secs:.1f} s)"}

    if cpu is not None and scan is not None:
        tok_s, threads, secs, nnz = cpu_cache_chain_tokens_per_s(1024, 512)
        scan["cpu_baseline"] = {"value": tok_s, "unit": "tokens/s", "cores": threads, "kind": "port",
                                "sample": f"oracle port of the reference cache chain (pre_acts -> topk -> scatter -> "
                                          f"nonzero, features/cache.py:206-218,73-92), 1024 tokens in 512-token "
                                          f"batches ({secs:.1f} s, {nnz} cached activations); the per-feature window "
                                          f"ranking the reference then runs feature by feature is not included"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "C2: d_model=4096 width=131072 k=64 SAE forward over 65536 bf16 tokens per GPU "
                                   "(encode+TopK+decode+FVU), inputs resident in HBM",
                       "global_batch_tokens": world * TOKENS, "parallelism": f"token-parallel x{world}, SAE replicated",
                       "precision": PRECISION[args.planes].replace("fp32 W_dec", f"{args.decode_dtype} W_dec"),
                       "l2": "inputs (x 512 MiB, weights 6 GiB) larger than the 126 MB L2"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "scan": scan, "fvu": fvu_val, "alt_fp16_decode": alt, "alt_mode4": alt4,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scan-tokens", type=int, default=1048576,
                    help="tokens of the feature-sharded top-activation scan (BASELINE C3: 1048576, C4: 4194304)")
    ap.add_argument("--scan-top", type=int, default=20, help="examples kept per feature (C3: 5, C4: 20)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scan-exchange", default=None, choices=["nccl", "push"],
                    help="per-chunk exchanges of the sharded scan: NCCL all-gathers (default) or the library's own "
                         "peer-memory all-gather (saeb_push_gather)")
    ap.add_argument("--scan-crosscheck", action="store_true",
                    help="N > 1: also run the token-parallel form of the scan and compare its lists with the "
                         "feature-sharded ones")
    ap.add_argument("--scan-phases", action="store_true", help="per-phase CUDA-event timing of the scan (diagnostic)")
    ap.add_argument("--alt-mode4", action="store_true",
                    help="also measure packed mode 4 (residual-correction refinement), with fp32 and fp16 W_dec")
    ap.add_argument("--alt-fp16-decode", action="store_true",
                    help="also measure the step with an fp16 copy of W_dec (reported as alt_fp16_decode)")
    ap.add_argument("--no-overlap", action="store_true", help="run the phases of a step back to back on one stream")
    ap.add_argument("--chunk", type=int, default=9472, help="tokens per pipeline chunk (multiple of 9472 = one wave of the GEMM grid)")
    ap.add_argument("--decode-dtype", default="fp32", choices=["fp32", "fp16"],
                    help="W_dec copy the decode gathers from: fp32 (parity default) or fp16 (half the bytes, ~2e-4 row error)")
    ap.add_argument("--planes", type=int, default=3, choices=[1, 2, 3, 4],
                    help="encoder mode: 3 = fp16 pass + exact refinement (default), 4 = fp16 pass + residual-correction "
                         "refinement, 2 = bf16 hi+lo, 1 = bf16 (diagnostic)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
