#!/usr/bin/env python
"""Benchmark of the SAE hot path (BASELINE.json metric: SAE tokens/sec at d=4096, width=131k).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step is one pass of the hot path -- fused encode + TopK, sparse decode, FVU (reference Sae.forward,
sae/sae.py:193-247) -- over one batch of 65 536 synthetic bf16 tokens (BASELINE.json configs[1], "C2").  With N GPUs
every rank runs the same step on its own 65 536 tokens with the SAE replicated (tokens are independent: no data-path
collective, weak scaling -- the trivial axis); `value` is the whole-job tokens/s, timed on the device with CUDA
events, barrier + synchronize on both sides, max over ranks.  The JSON line also carries
  e2e           the same metric through the reference-facing `Sae` objects with HOST buffers (pinned x in, TopK + FVU
                back), copies inside the timed region; e2e_full additionally copies the reconstruction back;
  roofline      the fused encode kernel against the measured bf16 tensor peak (MEASURED_PEAKS.json), plus the two
                HBM-bound gather kernels (refinement, decode) timed alone in the same run;
  parity_sample fp64 check of a few hundred rows of the timed batch (TopK index sets, values, reconstruction);
  c5            BASELINE configs[4]: steering hook over 32 768 fp16 tokens (clamp latent 12345 to 50, fp16 out);
  scan_c3/c4    BASELINE configs[2]/[3]: feature-sharded top-activation scan (features split across ranks, per-chunk
                threshold exchange, ONE all-gather of the per-shard top lists at the end): 1 M tokens top-5 and
                4 M tokens top-20.  The north-star multi-GPU curve is `scan_c4.tokens_per_s` over N;
  cpu_baseline  the oracle's CPU forward (the reference's PyTorch ops) on a bounded sample, rank 0, N=1 only.
`--impl reference` times the reference's own CPU implementation of the path (the oracle port of its PyTorch ops; a
Python reference cannot be pre-built into oracle/_ref) on all host cores, each step a bounded sample of the workload.
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
REF_SAMPLE_TOKENS, REF_BATCH = 1024, 512   # one step of the reference arm (CPU): two 512-token reference batches
WAVE_TOKENS = 9472

PRECISION = {
    3: "one fp16 tensor-core pass (activations exact after a power-of-two row scale, W_enc rounded to fp16, fp32 "
       "accumulate) + exact fp32 re-evaluation of every candidate inside the rigorous rounding bound; fp32 W_dec",
    "3b": "one fp16 tensor-core pass (activations exact after a power-of-two row scale, W_enc rounded to fp16, fp32 "
          "accumulate); TopK index set decided rigorously: every candidate whose rounding-error interval straddles "
          "the k-th boundary is re-evaluated exactly in fp32, members that are in the set whatever their exact value "
          "keep the tensor-core value (rigorous bound eps_j, ~5e-5 relative measured: see parity_sample; north-star "
          "bar 1e-3, indices bit-exact); fp32 W_dec",
    4: "one fp16 tensor-core pass + refinement by residual correction (x . fp16 residual plane of W_enc added to the "
       "tensor-core value for every candidate inside the rounding bound; values to ~1e-6 relative); fp32 W_dec",
    2: "bf16 activations (exact) x bf16 hi+lo W_enc planes (two tensor-core passes), fp32 accumulate; fp32 W_dec",
    1: "single bf16 pass (NOT parity grade; diagnostic only)",
}


def workload_config(world: int, planes=3, decode_dtype: str = "fp32"):
    """`config` of the JSON line -- the same dict on the GPU arm and on the reference arm"""
    return {"workload": "C2: d_model=4096 width=131072 k=64 SAE forward over 65536 bf16 tokens per GPU "
                        "(encode+TopK+decode+FVU), inputs resident in HBM",
            "global_batch_tokens": world * TOKENS, "parallelism": f"token-parallel x{world}, SAE replicated",
            "precision": PRECISION[planes].replace("fp32 W_dec", f"{decode_dtype} W_dec"),
            "l2": "inputs (x 512 MiB, weights 6 GiB) larger than the 126 MB L2",
            "reference_arm_sample": f"the CPU reference arm (--impl reference) times {REF_SAMPLE_TOKENS} tokens of this "
                                    f"workload per step ({REF_BATCH}-token batches, as the reference's dense [T, N] "
                                    f"latents require), same weights / token distribution"}


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
def _use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use every host core"""
    import torch

    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def cpu_forward_setup(seed: int = 1234):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import sae_oracle as O

    p = O.init_params(D_IN, WIDTH, K, seed)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(REF_SAMPLE_TOKENS, D_IN, generator=g).to(torch.bfloat16)
    return O, p, x


def cpu_forward_step(O, p, x, batch: int = REF_BATCH):
    """one bounded sample: the reference's Sae.forward over x in `batch`-token batches; returns seconds"""
    import torch

    t = time.perf_counter()
    with torch.no_grad():
        for b0 in range(0, x.shape[0], batch):
            O.forward(p, x[b0:b0 + batch])
    return time.perf_counter() - t


def cpu_forward_tokens_per_s(repeats: int):
    threads = _use_all_host_threads()
    O, p, x = cpu_forward_setup()
    cpu_forward_step(O, p, x[:REF_BATCH])  # warm-up
    secs = [cpu_forward_step(O, p, x) for _ in range(repeats)]
    return x.shape[0] * len(secs) / sum(secs), threads, sum(secs)


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
    """CPU arm: W warm-up steps, then exactly K timed steps, each one the bounded sample named in config"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = _use_all_host_threads()
    O, p, x = cpu_forward_setup()
    for _ in range(max(args.warmup, 0)):
        cpu_forward_step(O, p, x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_forward_step(O, p, x)
    total = time.perf_counter() - t0
    tok_s = REF_SAMPLE_TOKENS * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": tok_s, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(args.warmup, 0), "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(max(args.gpus, 1)),
        "cpu_baseline": {"value": tok_s, "unit": "tokens/s", "cores": threads, "kind": "port",
                         "sample": f"oracle port of the reference Sae.forward (PyTorch CPU fp32; its decode is the "
                                   f"reference's eager_decode, sae/utils.py:108-111, a second dense GEMM -- what the "
                                   f"reference itself runs on a CPU): {REF_SAMPLE_TOKENS} tokens in {REF_BATCH}-token "
                                   f"batches per step, {args.steps} timed steps after {max(args.warmup, 0)} warm-up steps, "
                                   f"one process on all host threads whatever N is"},
        "e2e": {"value": tok_s, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# GPU arm helpers
# ---------------------------------------------------------------------------------------------
def fp64_parity_sample(torch, sae, x, acts, idx, sae_out, rows, clamp=None, out_tol=1e-3):
    """fp64 evaluation of `rows` rows of the timed batch on the device (torch, outside every timed region): TopK index
    sets must be identical except on rows whose fp64 k / k+1 gap is below 2e-5 relative (near ties: two fp32
    implementations may legitimately differ there), values / reconstruction within 1e-3 relative."""
    W, b, bd, Wd = sae.encoder.weight.data, sae.encoder.bias.data, sae.b_dec.data, sae.W_dec.data
    r = torch.as_tensor(rows, device=x.device)
    xs = x[r].double() - bd.double()
    pre = torch.empty((len(rows), W.shape[0]), dtype=torch.float64, device=x.device)
    for n0 in range(0, W.shape[0], 16384):
        pre[:, n0:n0 + 16384] = xs @ W[n0:n0 + 16384].double().T + b[n0:n0 + 16384].double()
    pre.clamp_(min=0)
    if clamp is not None:
        pre[:, clamp[0]] = clamp[1]
    k = acts.shape[-1]
    rv, ri = pre.topk(k + 1, dim=-1)
    gap = (rv[:, k - 1] - rv[:, k]) / rv[:, k - 1].clamp_min(1e-30)
    ri_k = ri[:, :k]
    got_i = idx[r]
    same = (torch.sort(got_i, 1).values == torch.sort(ri_k, 1).values).all(1)
    near_tie = gap < 2e-5
    ref_at_got = torch.gather(pre, 1, got_i)   # fp64 values of the features the kernel selected
    rel_val = ((acts[r].double() - ref_at_got).abs() / ref_at_got.abs().clamp_min(1e-6)).max()
    out = {"rows": len(rows), "set_mismatch": int((~same).sum().item()),
           "set_mismatch_not_near_tie": int((~same & ~near_tie).sum().item()), "near_tie_rows": int(near_tie.sum().item()),
           "max_rel_val": float(rel_val.item()), "reference": "fp64 evaluation of the reference formula on the device"}
    if sae_out is not None:
        ref_out = torch.zeros((len(rows), W.shape[1]), dtype=torch.float64, device=x.device)
        for j in range(k):
            ref_out += rv[:, j:j + 1] * Wd[ri[:, j]].double()
        ref_out += bd.double()
        err = (sae_out[r].double() - ref_out).norm(dim=1) / ref_out.norm(dim=1)
        out["max_rel_row_err_out"] = float(err[same].max().item()) if bool(same.any()) else None
    out["pass"] = bool(out["set_mismatch_not_near_tie"] == 0 and out["max_rel_val"] < 1e-3
                       and (out.get("max_rel_row_err_out") or 0.0) < out_tol)
    return out


def time_events(torch, fn, iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run_scan(torch, sdist, engine, synth, sae, args, rank, world, dev, barrier, max_over_ranks, tokens, n_top, tag,
             peaks, crosscheck_tokens=0):
    """feature-sharded top-activation scan over `tokens` tokens (same tokens on every rank), top `n_top` windows of
    64 tokens per feature"""
    ctx_len = 64
    lo, hi = sdist.shard_range(WIDTH, world, rank)
    ops = sdist.EngineOps(sae.encoder.weight.data[lo:hi], sae.encoder.bias.data[lo:hi], sae.b_dec.data, lo, hi,
                          n_top, ctx_len, dev, planes=args.planes)
    if args.scan_exchange:
        ops.exchange = args.scan_exchange
    chunk = ops.chunk_tokens(world)
    xs = synth.make_activations(tokens, D_IN, dev, seed=99)  # same tokens on every rank

    def chunks(n=tokens):
        for t0 in range(0, n, chunk):
            yield xs[t0:min(n, t0 + chunk)]

    sdist.sharded_scan(chunks(min(tokens, 4 * chunk)), ops, K, ctx_len, WIDTH)  # warm-up (NCCL, scratch, lists)
    ops.scan = engine.TopActivationScan(lo, hi, n_top, ctx_len, dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    phases = {}
    res = sdist.sharded_scan(chunks(), ops, K, ctx_len, WIDTH, phase_times=phases if args.scan_phases else None)
    e1.record()
    barrier()
    sms = max_over_ranks(e0.elapsed_time(e1))
    tps = tokens / (sms * 1e-3)
    m1 = sdist.bounds_width(K, min(K, hi - lo), world)
    scan = {"workload": tag, "tokens": tokens, "features": WIDTH, "n_top": n_top, "ctx_len": ctx_len, "ms": sms,
            "tokens_per_s": tps, "sharding": f"features/{world}", "exact_topk_mask": True, "chunk_tokens": chunk,
            # every GPU runs 1/world of the encoder GEMM for every token
            "tensor_frac_per_gpu": tps * ENC_FLOPS_PER_TOKEN / world / 1e12 / peaks["tflops_sustained"],
            "exchange": None if world == 1 else getattr(ops, "exchange", None),
            "exchange1_columns": 2 * m1 if world > 1 else None,   # lower + upper bound lists
            "exchange2_columns": K if world > 1 else None,         # member values
            "nvlink_bytes_per_chunk_per_rank_received": None if world == 1 else chunk * 4 * (2 * m1 + K) * (world - 1),
            "refine_mode": "scan (value_mode 2): only TopK members that can still enter their feature's list are "
                           "gathered, exact values" if getattr(ops, "scan_value_mode", 0) == 2 else "every member exact",
            "final_allgather_bytes_per_rank": (hi - lo) * n_top * 12,
            "filled_features": int((res.top_win[:, 0] >= 0).sum().item()),
            "flagged_rows": int(ops.status.item()),
            "schedule": "sequential (phase timing)" if args.scan_phases else sdist.scan_schedule(world),
            "phase_ms_rank0": {k_: round(v_, 2) for k_, v_ in phases.items()} or None}
    if world > 1 and not args.no_scan_single:
        # the SAME scan on one GPU in the same run (every rank holds the full SAE for the forward bench and runs it on
        # its own GPU, max over ranks): the strong-scaling ratio can be formed from this record alone
        ops1 = sdist.EngineOps(sae.encoder.weight.data, sae.encoder.bias.data, sae.b_dec.data, 0, WIDTH, n_top,
                               ctx_len, dev, planes=args.planes)
        chunk1 = ops1.chunk_tokens(1)

        def chunks1(n=tokens):
            for t0 in range(0, n, chunk1):
                yield xs[t0:min(n, t0 + chunk1)]

        sdist.sharded_scan(chunks1(min(tokens, 3 * chunk1)), ops1, K, ctx_len, WIDTH, local=True)
        ops1.scan = engine.TopActivationScan(0, WIDTH, n_top, ctx_len, dev)
        barrier()
        e0.record()
        res1 = sdist.sharded_scan(chunks1(), ops1, K, ctx_len, WIDTH, local=True)
        e1.record()
        barrier()
        ms1 = max_over_ranks(e0.elapsed_time(e1))
        same1 = torch.tensor([1.0 if (torch.equal(res1.top_win, res.top_win) and torch.equal(res1.top_vals, res.top_vals))
                              else 0.0], device=dev)
        torch.distributed.all_reduce(same1, op=torch.distributed.ReduceOp.MIN)
        scan["single_gpu_same_run"] = {"ms": ms1, "tokens_per_s": tokens / (ms1 * 1e-3),
                                       "lists_equal_sharded": bool(same1.item() == 1.0)}
        scan["speedup_over_1_gpu_same_run"] = ms1 / sms
        del ops1, res1
        torch.cuda.empty_cache()
    if world > 1 and not args.scan_phases:
        # where a chunk's time goes: one SEQUENTIAL pass (every kernel with the GPU to itself, CUDA events between the
        # phases) over a bounded slice of the same tokens; the timed run above overlaps everything but the GEMM stream
        n_ph = min(tokens, 8 * chunk)
        ops.scan = engine.TopActivationScan(lo, hi, n_top, ctx_len, dev)
        ph = {}
        sdist.sharded_scan(chunks(n_ph), ops, K, ctx_len, WIDTH, phase_times=ph)
        per_m = {k_: round(v_ / n_ph * 1048576, 2) for k_, v_ in ph.items()}
        scan["phase_ms_per_1M_tokens_rank0"] = per_m
        scan["phase_note"] = (f"sequential diagnostic pass over the first {n_ph} tokens (lists start empty: the refinement "
                              f"gathers more than in steady state); pipelined, the chunk time is the GEMM stream plus what "
                              f"the co-resident chain costs it")
        scan["critical_phase"] = max(per_m, key=per_m.get) if per_m else None
        scan["pipelined_over_gemm_alone"] = (sms / tokens * 1048576) / per_m["gemm"] if per_m.get("gemm") else None
    if world > 1 and crosscheck_tokens > 0:
        # the token-parallel form of the same scan (full SAE on every rank, tokens split, no per-chunk exchange, one
        # all-gather + merge at the end) must give the same lists -- checked on a bounded slice of the same tokens
        n_cc = min(tokens, crosscheck_tokens)
        ops.scan = engine.TopActivationScan(lo, hi, n_top, ctx_len, dev)
        res_fs = sdist.sharded_scan(chunks(n_cc), ops, K, ctx_len, WIDTH)
        ops_tp = sdist.EngineOps(sae.encoder.weight.data, sae.encoder.bias.data, sae.b_dec.data, 0, WIDTH, n_top,
                                 ctx_len, dev, planes=args.planes)
        w_lo, w_hi = sdist.token_slice(n_cc // ctx_len, world, rank)
        x_tp = xs[w_lo * ctx_len:w_hi * ctx_len]
        chunk_tp = ops_tp.chunk_tokens(1)
        res_tp = sdist.token_parallel_scan((x_tp[t0:t0 + chunk_tp] for t0 in range(0, x_tp.shape[0], chunk_tp)),
                                           ops_tp, K, ctx_len, WIDTH, w_lo, n_top=n_top)
        same_w = bool(torch.equal(res_tp.top_win, res_fs.top_win))
        rel = ((res_tp.top_vals - res_fs.top_vals).abs() / res_fs.top_vals.abs().clamp_min(1e-30)).max()
        flags = torch.tensor([1.0 if same_w else 0.0], device=dev)
        torch.distributed.all_reduce(flags, op=torch.distributed.ReduceOp.MIN)
        scan["crosscheck_token_parallel"] = {"tokens": n_cc, "same_windows": bool(flags.item() == 1.0),
                                             "max_rel_score_diff": float(rel.item())}
        del ops_tp
    del ops, xs
    torch.cuda.empty_cache()
    return scan


def time_gathers(torch, L, _capi, engine, sae, enc, x, acts, idx, sae_out, peaks, value_mode):
    """refinement and decode kernels, each alone with full grids over the whole 65 536-token batch: achieved HBM GB/s
    from the ALGORITHMIC bytes (rows gathered x 16 KiB; the refinement's row count comes from its own counter)"""
    import ctypes

    check = _capi.check
    T = TOKENS
    st = torch.cuda.current_stream().cuda_stream
    prep = torch.empty(L.saeb_prep_bytes(T, D_IN), dtype=torch.uint8, device=x.device)
    ws = torch.empty(L.saeb_candidates_workspace_bytes(T, D_IN, WIDTH, K, 0), dtype=torch.uint8, device=x.device)
    status = torch.zeros(1, dtype=torch.int32, device=x.device)
    a2, i2 = torch.empty_like(acts), torch.empty_like(idx)
    check(L.saeb_prep_activations(x.data_ptr(), _capi.BF16, T, D_IN, D_IN, prep.data_ptr(), st), "prep")
    check(L.saeb_encode_candidates(prep.data_ptr(), T, 0, T, enc.blob.data_ptr(), D_IN, WIDTH, K, 0, -1, 0.0,
                                   ws.data_ptr(), ws.numel(), st), "gemm")

    def refine(merged):
        check(L.saeb_refine_candidates(x.data_ptr(), _capi.BF16, D_IN, prep.data_ptr(), T, 0, T, enc.blob.data_ptr(),
                                       enc.W_enc.data_ptr(), D_IN, WIDTH, K, 0, -1, 0.0, None, None, None, merged,
                                       a2.data_ptr(), None, i2.data_ptr(), status.data_ptr(), ws.data_ptr(), ws.numel(), 0, value_mode, st), "refine")

    refine(0)   # merges the candidate lists once; the timed calls below reuse the merged lists
    check(L.saeb_set_option(b"stats", 1), "stats")
    refine(1)
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 8)()
    L.saeb_debug_stats(buf)
    rows_gathered = int(buf[7])
    check(L.saeb_set_option(b"stats", 0), "stats")
    ms_ref = time_events(torch, lambda: refine(1), 3)
    sq = torch.zeros((), dtype=torch.float64, device=x.device)
    ms_dec = time_events(torch, lambda: engine.decode(idx, acts, sae.W_dec.data, sae.b_dec.data, x=x, sq_err=sq,
                                                      out=sae_out), 3)
    ref_bytes = rows_gathered * D_IN * 4 + T * (D_IN * 2 + 112 * 12 + K * 12)
    dec_bytes = T * (K * D_IN * 4 + D_IN * 4 + D_IN * 2 + K * 12)
    same = bool(torch.equal(a2, acts) and torch.equal(i2, idx))
    return {"bound": "hbm", "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "refine_kernel": {"ms": ms_ref, "rows_gathered_per_token": rows_gathered / T,
                              "algorithmic_bytes": ref_bytes, "achieved": ref_bytes / (ms_ref * 1e-3) / 1e9,
                              "frac": ref_bytes / (ms_ref * 1e-3) / 1e9 / peaks["hbm_gbs"],
                              "equals_timed_step_output": same},
            "decode_kernel": {"ms": ms_dec, "algorithmic_bytes": dec_bytes, "achieved": dec_bytes / (ms_dec * 1e-3) / 1e9,
                              "frac": dec_bytes / (ms_dec * 1e-3) / 1e9 / peaks["hbm_gbs"]},
            "timed": "each kernel alone over the whole batch (one CTA per token), CUDA events, 3 repeats"}


def run_c5(torch, engine, synth, sae, dev, args, peaks, max_over_ranks, barrier, world):
    """steering hook path (reference features/steering.py:105-124): fp16 hidden stream [1, 32768, 4096] -> encode with
    latent 12345 clamped to 50 -> TopK -> sparse decode -> fp16, through the mirror's hook body"""
    from sae_auto_interp.features.steering import steering_hook_output

    T5, feat, cval = 32768, 12345, 50.0
    h = synth.make_activations(T5, D_IN, dev, seed=7, dtype=torch.float16).unsqueeze(0)
    for _ in range(max(2, min(args.warmup, 3))):
        out = steering_hook_output(sae, h, feat, cval)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = max(3, min(args.steps, 10))
    e0.record()
    for _ in range(n):
        out = steering_hook_output(sae, h, feat, cval)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / n
    step1 = steering_hook_output(sae, h[:, :1], feat, cval)   # generation step (T = 1): no clamp, like the reference
    ms1 = time_events(torch, lambda: steering_hook_output(sae, h[:, :1], feat, cval), 20)
    enc = sae.encode(h, clamp_feature=feat, clamp_value=cval)
    clamped = bool(((enc.top_indices[0] == feat) & (enc.top_acts[0] == cval)).any(-1).all())
    par = fp64_parity_sample(torch, sae, h[0], enc.top_acts[0], enc.top_indices[0], out[0].float(),
                             list(range(0, T5, T5 // 128)), clamp=(feat, cval), out_tol=2e-3)   # fp16 output rounding
    return {"workload": "C5: steering hook, 32768 fp16 tokens, clamp latent 12345 to 50, fp16 reconstruction replaces "
                        "the layer output", "tokens": T5, "ms": ms, "tokens_per_s": world * T5 / (ms * 1e-3),
            "path_frac": T5 / (ms * 1e-3) * FLOPS_PER_TOKEN / 1e12 / peaks["tflops_sustained"],
            "generation_step_ms": ms1, "out_dtype": str(out.dtype), "clamp_in_every_row": clamped,
            "parity_sample": par, "step1_shape": list(step1.shape)}


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
    if args.gemm_stages:
        _capi.check(L.saeb_set_option(b"gemm_stages", args.gemm_stages), "set_option")
    peaks = load_peaks()

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
    sae.refine_values = "all" if args.values == "exact" else "boundary"
    value_mode = sae._value_mode()
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

        ov = OverlappedForward(enc, W_dec_used, sae.b_dec.data, K, chunk=args.chunk, value_mode=value_mode)

    def step():
        sq_err.zero_()
        if ov is not None:   # GEMM launches of chunk c+1 beside the bounded gather grids of chunk c
            ov.run(x, acts, idx, sae_out, sq_err)
        else:
            engine.encode_topk(x, enc, K, out_vals=acts, out_idx=idx, value_mode=value_mode)
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

    # ---- parity of the timed batch (outside the timed region): 256 rows spread over the batch vs fp64
    parity = fp64_parity_sample(torch, sae, x, acts, idx, sae_out, list(range(0, TOKENS, TOKENS // 256)))

    # ---- dominant kernel, live CUDA events on the launching stream (library-side bracket of the GEMM launches)
    k_ms = []
    for _ in range(max(3, min(args.steps, 5))):
        engine.encode_topk(x, enc, K, out_vals=acts, out_idx=idx, value_mode=value_mode)
        k_ms.append(float(L.saeb_profile_last_encode_ms()))
    k_avg = sum(k_ms) / len(k_ms)
    achieved = ENC_FLOPS_PER_TOKEN * TOKENS / (k_avg * 1e-3) / 1e12
    traffic = traffic_src = None
    summ = os.path.join(ROOT, "profiles", "encode_kernel_traffic.json")
    if os.path.exists(summ):
        try:
            tj = json.load(open(summ))
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
        except Exception:
            traffic = None
    n_launch = (TOKENS + WAVE_TOKENS - 1) // WAVE_TOKENS
    # the two HBM-bound gather kernels, each timed alone (full grids) over the whole batch in this run
    gathers = None
    if args.planes == 3:
        try:
            gathers = time_gathers(torch, L, _capi, engine, sae, enc, x, acts, idx, sae_out, peaks, value_mode)
        except Exception as exc:   # diagnostics must never take the benchmark down
            gathers = {"error": repr(exc)[:300]}
    roofline = {"bound": "tensor", "kernel": "encode_topk_kernel (tcgen05 GEMM + fused TopK)", "achieved": achieved,
                "launches_per_step": n_launch, "tokens_per_launch": WAVE_TOKENS,
                "algorithmic_flops_per_launch": ENC_FLOPS_PER_TOKEN * WAVE_TOKENS,
                "avg_launch_ms": k_avg * WAVE_TOKENS / TOKENS,
                "peak": peaks["tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops_sustained"],
                "peak_burst": peaks["tflops_burst"], "frac_of_burst_peak": achieved / peaks["tflops_burst"],
                "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": f"{peaks['source']} bf16 sustained (cuBLAS, MEASURED_PEAKS.json)",
                "kernel_ms": k_avg, "kernel_timed": "alone (no concurrent gathers), CUDA events around the step's "
                                                    f"{n_launch} launches on their stream",
                "mma_passes": 1 if enc.planes >= 3 else enc.planes,
                "path_frac": (value / world) * FLOPS_PER_TOKEN / 1e12 / peaks["tflops_sustained"],
                "path_frac_of_burst_peak": (value / world) * FLOPS_PER_TOKEN / 1e12 / peaks["tflops_burst"],
                "hbm_frac": (value / world) * (2 * D_IN + K * 12 + K * 4 * D_IN + 4 * D_IN
                                               + 2.0 * D_IN * WIDTH * (1 if enc.planes >= 3 else enc.planes) / TOKENS)
                / 1e9 / peaks["hbm_gbs"],
                "gather_kernels": gathers}

    # ---- opt-in (--alt-fp16-decode / --alt-mode4): the same step with half the decode / refinement gather bytes; both
    # inside the 1e-3 bar of BASELINE.json but not the parity-grade default, so reported NEXT to `value`
    alts = None
    if world == 1 and ov is not None and args.planes == 3 and not args.no_alt_modes:
        alts = {}
        from saeb200.overlap import OverlappedForward

        # (tag, packed mode, W_dec dtype, value mode): the other value mode of the default path is always measured
        other = engine.VALUES_EXACT if value_mode == engine.VALUES_BOUNDARY else engine.VALUES_BOUNDARY
        combos = [("exact_values" if other == engine.VALUES_EXACT else "boundary_values", 3, torch.float32, other)]
        if args.alt_fp16_decode:
            combos.append(("fp16_decode", 3, torch.float16, value_mode))
        if args.alt_mode4:
            combos += [("mode4_fp32_decode", 4, torch.float32, value_mode), ("mode4_fp16_decode", 4, torch.float16, value_mode)]
        for tag, pl, wdt, vm in combos:
            try:
                enc_a = sae.packed_encoder(pl)
                Wd = sae.W_dec.data if wdt == torch.float32 else sae.W_dec.data.to(wdt)
                ov_a = OverlappedForward(enc_a, Wd, sae.b_dec.data, K, chunk=args.chunk, value_mode=vm)
                out_a, acts_a, idx_a = torch.empty_like(sae_out), torch.empty_like(acts), torch.empty_like(idx)
                sq_a = torch.zeros_like(sq_err)
                for _ in range(2):
                    ov_a.run(x, acts_a, idx_a, out_a, sq_a)
                torch.cuda.synchronize()

                def step_a():
                    sq_a.zero_()
                    ov_a.run(x, acts_a, idx_a, out_a, sq_a)
                    return (sq_a / engine.total_variance(x)).to(torch.float32)

                ms_a = time_events(torch, step_a, args.steps)
                par = fp64_parity_sample(torch, sae, x, acts_a, idx_a, out_a, list(range(0, TOKENS, TOKENS // 256)))
                same = (torch.sort(idx_a, 1).values == torch.sort(idx, 1).values).all(1)
                alts[tag] = {"value": TOKENS / (ms_a * 1e-3), "unit": "tokens/s", "ms_per_step": ms_a,
                             "path_frac": TOKENS / (ms_a * 1e-3) * FLOPS_PER_TOKEN / 1e12 / peaks["tflops_sustained"],
                             "rows_with_different_topk_set_vs_default": int((~same).sum().item()),
                             "parity_sample": par}
                del ov_a, out_a, acts_a, idx_a
            except Exception as exc:
                alts[tag] = {"error": repr(exc)[:300]}
        torch.cuda.empty_cache()

    # ---- end to end through the reference-facing objects with host buffers (`e2e`, `e2e_full`)
    x_host = synth.make_activations(TOKENS, D_IN, dev, seed=1 + rank, pinned_host=True)
    acts_host = torch.empty((TOKENS, K), dtype=torch.float32, pin_memory=True)
    idx_host = torch.empty((TOKENS, K), dtype=torch.int64, pin_memory=True)
    hf = pipeline.HostForward(sae, TOKENS, chunk=args.chunk)

    def run_e2e(out_host):
        for _ in range(max(1, min(args.warmup, 2))):
            hf.run(x_host, acts_host, idx_host, out_host)
        barrier()
        e0.record()
        for _ in range(args.steps):
            f = hf.run(x_host, acts_host, idx_host, out_host)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        return world * TOKENS * args.steps / (ms * 1e-3), float(f.item())

    v_e2e, fvu_host = run_e2e(None)
    e2e = {"value": v_e2e, "unit": "tokens/s", "h2d_bytes_per_step": hf.h2d_bytes * world,
           "d2h_bytes_per_step": hf.d2h_bytes(False) * world,
           "api": "sae_auto_interp.sae.Sae + saeb200.pipeline.HostForward (pinned x in; TopK acts/indices + FVU out)",
           "fvu": fvu_host}
    e2e_full = None
    try:
        out_host = torch.empty((TOKENS, D_IN), dtype=torch.float32, pin_memory=True)
        v_full, fvu_full = run_e2e(out_host)
        e2e_full = {"value": v_full, "unit": "tokens/s", "h2d_bytes_per_step": hf.h2d_bytes * world,
                    "d2h_bytes_per_step": hf.d2h_bytes(True) * world,
                    "api": "same call, additionally copying the reconstruction sae_out [T, 4096] fp32 back (every "
                           "field of the reference's ForwardOutput)", "fvu": fvu_full,
                    "sae_out_equals_device_run": bool(torch.equal(out_host[:4096], sae_out[:4096].cpu()))}
        del out_host
    except Exception as exc:
        e2e_full = {"error": repr(exc)[:300]}
    del hf, x_host

    # ---- BASELINE configs[4]: steering hook, 32 768 fp16 tokens, clamp latent 12345 to 50, fp16 reconstruction
    c5 = None
    if not args.no_c5:
        try:
            c5 = run_c5(torch, engine, synth, sae, dev, args, peaks, max_over_ranks, barrier, world)
        except Exception as exc:
            c5 = {"error": repr(exc)[:300]}

    # ---- feature-sharded top-activation scans (BASELINE configs[2] and [3])
    del ov, x, acts, idx, sae_out
    torch.cuda.empty_cache()
    scans = {}
    for tag, tokens, n_top, name in (("scan_c3", args.scan_c3_tokens, 5, "C3: 1M tokens, top-5 windows per feature"),
                                     ("scan_c4", args.scan_c4_tokens, 20, "C4: 4M tokens, top-20 windows per feature")):
        if tokens <= 0:
            scans[tag] = None
            continue
        scans[tag] = run_scan(torch, sdist, engine, synth, sae, args, rank, world, dev, barrier, max_over_ranks,
                              tokens, n_top, name, peaks,
                              crosscheck_tokens=args.scan_crosscheck_tokens if tag == "scan_c3" else 0)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        tok_s, threads, secs = cpu_forward_tokens_per_s(repeats=2)
        cpu = {"value": tok_s, "unit": "tokens/s", "cores": threads, "kind": "port",
               "sample": f"oracle port of reference Sae.forward (PyTorch CPU fp32), 2 x {REF_SAMPLE_TOKENS} tokens in "
                         f"{REF_BATCH}-token batches ({secs:.1f} s)"}
        if scans.get("scan_c3") is not None:
            tok_s, threads, secs, nnz = cpu_cache_chain_tokens_per_s(1024, 512)
            scans["scan_c3"]["cpu_baseline"] = {
                "value": tok_s, "unit": "tokens/s", "cores": threads, "kind": "port",
                "sample": f"oracle port of the reference cache chain (pre_acts -> topk -> scatter -> nonzero, "
                          f"features/cache.py:206-218,73-92), 1024 tokens in 512-token batches ({secs:.1f} s, {nnz} cached "
                          f"activations); the per-feature window ranking the reference then runs feature by feature "
                          f"is not included"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "scaling_note": "`value` is the BASELINE metric on the token-parallel axis (SAE replicated, no data-path "
                            "collective); the north-star multi-GPU curve is scan_c4.tokens_per_s (feature-sharded, "
                            "strong scaling over N)",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(world, "3b" if (args.planes == 3 and value_mode == 1) else args.planes,
                                      args.decode_dtype),
            "clocks": clocks, "e2e": e2e, "e2e_full": e2e_full, "gpu_launches": int(launches), "roofline": roofline,
            "parity_sample": parity, "cpu_baseline": cpu, "c5": c5, "scan_c3": scans["scan_c3"],
            "scan_c4": scans["scan_c4"], "fvu": fvu_val, "alt_modes": alts,
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
    ap.add_argument("--scan-c3-tokens", type=int, default=1048576, help="tokens of the C3 scan (top-5); 0 = skip")
    ap.add_argument("--scan-c4-tokens", type=int, default=4194304, help="tokens of the C4 scan (top-20); 0 = skip")
    ap.add_argument("--scan-crosscheck-tokens", type=int, default=262144,
                    help="N > 1: tokens on which the feature-sharded lists are compared with the token-parallel form")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-scan-single", action="store_true",
                    help="N > 1: skip the one-GPU run of the same scan (the denominator of the strong-scaling ratio)")
    ap.add_argument("--no-c5", action="store_true")
    ap.add_argument("--no-alt-modes", action="store_true")
    ap.add_argument("--values", default="boundary", choices=["boundary", "exact"],
                    help="which TopK values the refinement re-evaluates exactly: boundary (default: exact index set, "
                         "exact values only where they decide it) or exact (every value; reported as alt_modes "
                         "otherwise)")
    ap.add_argument("--scan-exchange", default=None, choices=["nccl", "push"],
                    help="per-chunk exchanges of the sharded scan: NCCL all-gathers or the library's own "
                         "peer-memory all-gather (saeb_push_gather)")
    ap.add_argument("--scan-phases", action="store_true", help="per-phase CUDA-event timing of the scan (diagnostic)")
    ap.add_argument("--alt-mode4", action="store_true",
                    help="also measure packed mode 4 (residual-correction refinement), with fp32 and fp16 W_dec")
    ap.add_argument("--alt-fp16-decode", action="store_true",
                    help="also measure the step with an fp16 copy of W_dec")
    ap.add_argument("--no-overlap", action="store_true", help="run the phases of a step back to back on one stream")
    ap.add_argument("--chunk", type=int, default=WAVE_TOKENS,
                    help="tokens per pipeline chunk (multiple of 9472 = one wave of the GEMM grid)")
    ap.add_argument("--gemm-stages", type=int, default=0, help="depth of the GEMM's shared-memory ring (0 = default)")
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
