"""Bring-up probe for the GPU box: each experiment runs in its own process with a timeout so that a hung or trapped
kernel cannot take the whole call down.  Usage: python tools/gpu_probe.py [exp ...] ; results -> gpurun_out/probe.log"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-sae_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def _setup(pair):
    import torch
    from saeb200 import _capi, engine
    L = _capi.lib()
    _capi.check(L.saeb_set_option(b"cta_pair", pair), "set_option")
    return torch, engine


def exp_gemm(pair, planes, T, d, N, xdt="bf16"):
    torch, engine = _setup(pair)
    g = torch.Generator(device="cuda").manual_seed(1)
    W = (torch.rand(N, d, device="cuda", generator=g) * 2 - 1) / d ** 0.5
    be = torch.randn(N, device="cuda", generator=g) * 0.01
    bd = torch.randn(d, device="cuda", generator=g) * 0.1
    x = torch.randn(T, d, device="cuda", generator=g)
    x = {"bf16": x.to(torch.bfloat16), "f16": x.to(torch.float16), "f32": x}[xdt]
    enc = engine.PackedEncoder.pack(W, be, bd, planes)
    torch.cuda.synchronize()
    _, _, dense = engine.encode_topk(x, enc, 16, want_dense=True, want_topk=False)
    torch.cuda.synchronize()
    ref = torch.relu((x.double() - bd.double()) @ W.double().T + be.double())
    err = (dense.double() - ref).abs().max().item()
    rel = err / ref.abs().max().item()
    # error of a plain bf16-weight product for scale
    Wb = W.to(torch.bfloat16).double()
    ref1 = torch.relu(x.double() @ Wb.T + (be.double() - W.double() @ bd.double()))
    err1 = (dense.double() - ref1).abs().max().item()
    bad = (dense.double() - ref).abs() > 1e-2
    return dict(max_abs_err_vs_f64=err, rel=rel, max_abs_err_vs_bf16W=err1, n_bad=int(bad.sum()),
                first_bad=[int(v) for v in torch.nonzero(bad)[0].tolist()] if bad.any() else None,
                dense_absmax=dense.abs().max().item(), ref_absmax=ref.abs().max().item())


def exp_topk(pair, planes, T, d, N, k, xdt="bf16"):
    torch, engine = _setup(pair)
    g = torch.Generator(device="cuda").manual_seed(2)
    W = (torch.rand(N, d, device="cuda", generator=g) * 2 - 1) / d ** 0.5
    be = torch.randn(N, device="cuda", generator=g) * 0.01
    bd = torch.randn(d, device="cuda", generator=g) * 0.1
    x = torch.randn(T, d, device="cuda", generator=g)
    x = {"bf16": x.to(torch.bfloat16), "f16": x.to(torch.float16), "f32": x}[xdt]
    enc = engine.PackedEncoder.pack(W, be, bd, planes)
    vals, idx, _ = engine.encode_topk(x, enc, k)
    torch.cuda.synchronize()
    chunks = []
    bad_rows = 0
    maxrel = 0.0
    for t0 in range(0, T, 1024):
        xs = x[t0:t0 + 1024]
        ref = torch.relu((xs.double() - bd.double()) @ W.double().T + be.double())
        rv, ri = ref.topk(k, dim=-1)
        a = torch.sort(idx[t0:t0 + 1024], dim=-1).values
        b = torch.sort(ri, dim=-1).values
        bad_rows += int((a != b).any(-1).sum())
        maxrel = max(maxrel, ((vals[t0:t0 + 1024].double() - rv).abs() / rv.abs().clamp_min(1e-6)).max().item())
    sorted_ok = bool((vals[:, :-1] >= vals[:, 1:]).all())
    return dict(rows=T, rows_with_set_mismatch=bad_rows, max_rel_val_err=maxrel, sorted_desc=sorted_ok)


def exp_time(pair, planes, T, d, N, k, iters=3, splits=0, dbg=0, hints=1, persist=0, prefetch=2):
    torch, engine = _setup(pair)
    from saeb200 import _capi
    _capi.check(_capi.lib().saeb_set_option(b"prefetch_b", prefetch), "set_option")
    _capi.check(_capi.lib().saeb_set_option(b"debug_tiles", dbg), "set_option")
    _capi.check(_capi.lib().saeb_set_option(b"l2_hints", hints), "set_option")
    _capi.check(_capi.lib().saeb_set_option(b"persist_a", persist), "set_option")
    _capi.check(_capi.lib().saeb_set_option(b"splits", splits), "set_option")
    _capi.check(_capi.lib().saeb_set_option(b"profile", 1), "set_option")
    g = torch.Generator(device="cuda").manual_seed(3)
    W = (torch.rand(N, d, device="cuda", generator=g) * 2 - 1) / d ** 0.5
    be = torch.randn(N, device="cuda", generator=g) * 0.01
    bd = torch.randn(d, device="cuda", generator=g) * 0.1
    enc = engine.PackedEncoder.pack(W, be, bd, planes)
    Wd = W / W.norm(dim=1, keepdim=True)
    del W
    x = torch.randn(T, d, device="cuda", generator=g).to(torch.bfloat16)
    out = {}
    for _ in range(2):
        vals, idx, _ = engine.encode_topk(x, enc, k)
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    te, td, tk = [], [], []
    for _ in range(iters):
        e0.record()
        vals, idx, _ = engine.encode_topk(x, enc, k)
        e1.record()
        tk.append(float(_capi.lib().saeb_profile_last_encode_ms()))
        y = engine.decode(idx, vals, Wd, bd)
        e2.record()
        torch.cuda.synchronize()
        te.append(e0.elapsed_time(e1))
        td.append(e1.elapsed_time(e2))
    flops = 2.0 * T * d * N
    out["encode_ms"] = te
    out["gemm_kernel_ms"] = tk
    if planes == 3:
        out["flagged_rows"] = int(engine.encode_topk.last_status.item())
    out["decode_ms"] = td
    out["encode_tflops_alg"] = flops / (min(te) * 1e-3) / 1e12
    out["decode_GBs"] = (T * k * d * 4 + T * d * 4 + T * k * 12) / (min(td) * 1e-3) / 1e9
    out["tokens_per_s_fwd"] = T / ((min(te) + min(td)) * 1e-3)
    return out


def exp_decode(T, d, N, k):
    torch, engine = _setup(2)
    g = torch.Generator(device="cuda").manual_seed(4)
    Wd = torch.randn(N, d, device="cuda", generator=g)
    bd = torch.randn(d, device="cuda", generator=g)
    lat = torch.rand(T, N, device="cuda", generator=g)
    vals, idx = lat.topk(k)
    y = engine.decode(idx, vals, Wd, bd)
    buf = torch.zeros(T, N, device="cuda", dtype=torch.float64)
    buf.scatter_(-1, idx, vals.double())
    ref = buf @ Wd.double() + bd.double()
    return dict(max_abs_err=(y.double() - ref).abs().max().item(), ref_absmax=ref.abs().max().item())


def exp_overlap(T, chunk, planes=3, overlap=True, l2_hints=1, splits=0, iters=3, persist=1, chunking=1, margin=0):
    torch, engine = _setup(2)
    from saeb200 import _capi, synth
    from saeb200.overlap import OverlappedForward
    L = _capi.lib()
    _capi.check(L.saeb_set_option(b"refine_margin", margin), "set_option")
    _capi.check(L.saeb_set_option(b"persist_a", persist), "set_option")
    _capi.check(L.saeb_set_option(b"chunking", chunking), "set_option")
    _capi.check(L.saeb_set_option(b"l2_hints", l2_hints), "set_option")
    _capi.check(L.saeb_set_option(b"splits", splits), "set_option")
    sae = synth.make_sae(4096, 131072, 64, "cuda", seed=1234)
    sae.encoder_planes = planes
    enc = sae.packed_encoder()
    x = synth.make_activations(T, 4096, "cuda", seed=3)
    acts = torch.empty((T, 64), dtype=torch.float32, device="cuda")
    idx = torch.empty((T, 64), dtype=torch.int64, device="cuda")
    out = torch.empty((T, 4096), dtype=torch.float32, device="cuda")
    sq = torch.zeros((), dtype=torch.float64, device="cuda")
    ov = OverlappedForward(enc, sae.W_dec.data, sae.b_dec.data, 64, chunk=chunk) if (overlap and planes == 3) else None

    def step():
        sq.zero_()
        if ov is not None:
            ov.run(x, acts, idx, out, sq)
        else:
            engine.encode_topk(x, enc, 64, out_vals=acts, out_idx=idx)
            engine.decode(idx, acts, sae.W_dec.data, sae.b_dec.data, x=x, sq_err=sq, out=out)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for _ in range(iters):
        e0.record(); step(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    res = dict(ms=ms, tokens_per_s=T / (min(ms) * 1e-3), flops_frac_of_1418=T / (min(ms) * 1e-3) * 1.0743e9 / 1.4181e15)
    if planes == 3:
        res["flagged"] = int(ov.status.sum().item()) if ov is not None else int(engine.encode_topk.last_status.item())
    # consistency of the overlapped result with the sequential one
    if ov is not None:
        a2, i2, _ = engine.encode_topk(x[:4096], enc, 64)
        res["overlap_equals_sequential"] = bool(torch.equal(a2, acts[:4096]) and torch.equal(i2, idx[:4096]))
    return res


def exp_stats(planes, T=9472, margin_k=64):
    """cycle accounting of one single-wave launch of the fused encode kernel"""
    import ctypes
    torch, engine = _setup(2)
    from saeb200 import _capi, synth
    L = _capi.lib()
    sae = synth.make_sae(4096, 131072, 64, "cuda", seed=1234)
    sae.encoder_planes = planes
    enc = sae.packed_encoder()
    x = synth.make_activations(T, 4096, "cuda", seed=3)
    for _ in range(2):
        engine.encode_topk(x, enc, margin_k)
    torch.cuda.synchronize()
    _capi.check(L.saeb_set_option(b"stats", 1), "set_option")
    _capi.check(L.saeb_set_option(b"profile", 1), "set_option")
    engine.encode_topk(x, enc, margin_k)
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 8)()
    L.saeb_debug_stats(buf)
    v = list(buf)
    n = max(v[6], 1)
    tot = v[5] / n
    return dict(gemm_ms=float(L.saeb_profile_last_encode_ms()), pairs=v[6], kernel_cycles=tot,
                producer_wait_empty=v[0] / n / tot, mma_wait_tmem=v[1] / n / tot, mma_wait_tma=v[2] / n / tot,
                epi_wait_acc=v[3] / n / tot, epi_compaction=v[4] / n / tot)


def exp_scan(T=131072, world=1, rank=0, planes=3):
    """feature-sharded scan of one (simulated) rank: features [rank*N/world, (rank+1)*N/world), no collectives"""
    torch, engine = _setup(2)
    from saeb200 import dist as sdist, synth
    sae = synth.make_sae(4096, 131072, 64, "cuda", seed=1234)
    lo, hi = sdist.shard_range(131072, world, rank)
    ops = sdist.EngineOps(sae.encoder.weight.data[lo:hi], sae.encoder.bias.data[lo:hi], sae.b_dec.data, lo, hi, 20, 64,
                          "cuda", planes=planes)
    x = synth.make_activations(T, 4096, "cuda", seed=5)
    chunk = 18944

    def run():
        ops.scan = engine.TopActivationScan(lo, hi, 20, 64, "cuda")
        sdist.sharded_scan((x[t:t + chunk] for t in range(0, T, chunk)), ops, 64, 64, 131072)
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return dict(ms=ms, tokens_per_s=T / (ms * 1e-3), world=world)


def exp_scan_sim(T=151552, world=8, rank=3):
    """per-rank cost of the feature-sharded scan with a realistic external lower bound (no collectives): the global
    k-th values come from a full single-GPU encode, ext_L = 0.985 * kth"""
    torch, engine = _setup(2)
    from saeb200 import dist as sdist, synth
    sae = synth.make_sae(4096, 131072, 64, "cuda", seed=1234)
    x = synth.make_activations(T, 4096, "cuda", seed=5)
    kth = sae.encode(x).top_acts[:, -1].contiguous()
    lo, hi = sdist.shard_range(131072, world, rank)
    ops = sdist.EngineOps(sae.encoder.weight.data[lo:hi], sae.encoder.bias.data[lo:hi], sae.b_dec.data, lo, hi, 20, 64,
                          "cuda")
    chunk = 37888
    ph = {}

    def run(timed):
        ops.scan = engine.TopActivationScan(lo, hi, 20, 64, "cuda")
        marks = []

        def mark(n):
            if timed:
                e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((n, e))
        mark("start")
        for t in range(0, T, chunk):
            xc = x[t:t + chunk]
            lb = ops.local_bounds(xc, 64); mark("gemm+bounds")
            v, i = ops.local_topk(kth[t:t + chunk] * 0.985); mark("refine")
            ops.scan_update(v, i, t // 64, kth[t:t + chunk]); mark("scan_update")
        ops.scan_finalize(); mark("finalize")
        if timed:
            torch.cuda.synchronize()
            for (n0, e0), (n1, e1) in zip(marks, marks[1:]):
                ph[n1] = ph.get(n1, 0.0) + e0.elapsed_time(e1)
    run(False); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(True); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return dict(ms=ms, tokens_per_s=T / (ms * 1e-3), phases={k: round(v, 2) for k, v in ph.items()},
                flagged=int(ops.status.item()))


def exp_steering(T=32768):
    """BASELINE config 5: fp16 hidden stream [1, T, 4096] -> encode with one latent clamped -> fp16 reconstruction"""
    torch, engine = _setup(2)
    from saeb200 import synth
    from sae_auto_interp.features.steering import steering_hook_output
    sae = synth.make_sae(4096, 131072, 64, "cuda", seed=1234)
    h = synth.make_activations(T, 4096, "cuda", seed=7, dtype=torch.float16).unsqueeze(0)
    for _ in range(2):
        out = steering_hook_output(sae, h, 12345, 50.0)
    step = steering_hook_output(sae, h[:, :1], 12345, 50.0)   # decode step (T = 1): no clamp
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(); out = steering_hook_output(sae, h, 12345, 50.0); e1.record()
    for _ in range(20):
        step = steering_hook_output(sae, h[:, :1], 12345, 50.0)
    e2.record(); torch.cuda.synchronize()
    enc = sae.encode(h, clamp_feature=12345, clamp_value=50.0)
    clamped = bool(((enc.top_indices[0] == 12345) & (enc.top_acts[0] == 50.0)).any(-1).all())
    return dict(prefill_ms=e0.elapsed_time(e1), prefill_tokens_per_s=T / (e0.elapsed_time(e1) * 1e-3),
                decode_step_ms=e1.elapsed_time(e2) / 20, out_dtype=str(out.dtype), clamp_in_every_row=clamped)


def exp_cache(T=65536, seq=2048):
    """cache path (features/cache.py:206-218 + :73-92): hidden [B, seq, d] bf16 -> encode -> COO triples on device"""
    torch, engine = _setup(2)
    from saeb200 import synth
    sae = synth.make_sae(4096, 131072, 64, "cuda", seed=1234)
    h = synth.make_activations(T, 4096, "cuda", seed=8).view(T // seq, seq, 4096)
    filt = engine.make_filter_bitmap(torch.arange(5000, device="cuda"), 131072)   # README: first-5k-feature filter
    res = {}
    for name, bm in (("nofilter", None), ("filter5k", filt)):
        for _ in range(2):
            enc = sae.encode(h)
            loc, act = engine.coo_extract(enc.top_acts, enc.top_indices, seq, filter_bitmap=bm)
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(); enc = sae.encode(h); e1.record()
        loc, act = engine.coo_extract(enc.top_acts, enc.top_indices, seq, filter_bitmap=bm); e2.record()
        torch.cuda.synchronize()
        res[name] = dict(encode_ms=e0.elapsed_time(e1), coo_ms=e1.elapsed_time(e2), nnz=int(loc.shape[0]),
                         tokens_per_s=T / ((e0.elapsed_time(e1) + e1.elapsed_time(e2)) * 1e-3))
    return res


EXPS = {
    "gemm_p1_small": lambda: exp_gemm(1, 1, 256, 128, 512),
    "gemm_p2_small": lambda: exp_gemm(2, 1, 256, 128, 512),
    "gemm_p1_mid": lambda: exp_gemm(1, 1, 300, 4096, 2048 + 64),
    "gemm_p2_mid": lambda: exp_gemm(2, 1, 300, 4096, 2048 + 64),
    "gemm_p1_hilo": lambda: exp_gemm(1, 2, 512, 4096, 4096),
    "gemm_p2_hilo": lambda: exp_gemm(2, 2, 512, 4096, 4096),
    "gemm_p2_f32x": lambda: exp_gemm(2, 2, 512, 1024, 4096, "f32"),
    "gemm_p2_f16x": lambda: exp_gemm(2, 2, 512, 1024, 4096, "f16"),
    "topk_p1": lambda: exp_topk(1, 2, 2048, 1024, 16384, 64),
    "topk_p2": lambda: exp_topk(2, 2, 2048, 1024, 16384, 64),
    "topk_p2_k256": lambda: exp_topk(2, 2, 1024, 1024, 16384, 256),
    "topk_p2_full": lambda: exp_topk(2, 2, 2048, 4096, 131072, 64),
    "decode": lambda: exp_decode(64, 512, 4096, 32),
    "time_p2_2pl": lambda: exp_time(2, 2, 16384, 4096, 131072, 64),
    "time_p2_1pl": lambda: exp_time(2, 1, 16384, 4096, 131072, 64),
    "time_p1_2pl": lambda: exp_time(1, 2, 16384, 4096, 131072, 64),
    "time_p1_1pl": lambda: exp_time(1, 1, 16384, 4096, 131072, 64),
    "time_p2_2pl_64k": lambda: exp_time(2, 2, 65536, 4096, 131072, 64, iters=2),
    "topk_refine": lambda: exp_topk(2, 3, 2048, 1024, 16384, 64),
    "topk_refine_full": lambda: exp_topk(2, 3, 2048, 4096, 131072, 64),
    "topk_refine_f32x": lambda: exp_topk(2, 3, 1024, 1024, 16384, 64, "f32"),
    "time_refine_16k": lambda: exp_time(2, 3, 16384, 4096, 131072, 64),
    "time_refine_64k": lambda: exp_time(2, 3, 65536, 4096, 131072, 64, iters=2),
    "time_refine_64k_s2": lambda: exp_time(2, 3, 65536, 4096, 131072, 64, iters=2, splits=2),
    "t_h0": lambda: exp_time(2, 1, 65536, 4096, 131072, 64, iters=2, hints=0),
    "t_h1": lambda: exp_time(2, 1, 65536, 4096, 131072, 64, iters=2, hints=1),
    "t_h2": lambda: exp_time(2, 1, 65536, 4096, 131072, 64, iters=2, hints=2),
    "t_h3": lambda: exp_time(2, 1, 65536, 4096, 131072, 64, iters=2, hints=3),
    "t_s8": lambda: exp_time(2, 1, 65536, 4096, 131072, 64, iters=2, splits=8),
    "t_9472": lambda: exp_time(2, 1, 9472, 4096, 131072, 64, iters=3, splits=2),
    "t_pf0": lambda: exp_time(2, 3, 9472, 4096, 131072, 64, iters=4, persist=1, prefetch=0),
    "t_pf1": lambda: exp_time(2, 3, 9472, 4096, 131072, 64, iters=4, persist=1, prefetch=1),
    "t_pf2": lambda: exp_time(2, 3, 9472, 4096, 131072, 64, iters=4, persist=1, prefetch=2),
    "t_pf4": lambda: exp_time(2, 3, 9472, 4096, 131072, 64, iters=4, persist=1, prefetch=4),
    "t_p1": lambda: exp_time(2, 3, 9472, 4096, 131072, 64, iters=3, persist=1),
    "t_p2": lambda: exp_time(2, 3, 9472, 4096, 131072, 64, iters=3, persist=2),
    "t_p3": lambda: exp_time(2, 3, 9472, 4096, 131072, 64, iters=3, persist=3),
    "t_p0": lambda: exp_time(2, 3, 9472, 4096, 131072, 64, iters=3, persist=0),
    "t_9472_persist": lambda: exp_time(2, 1, 9472, 4096, 131072, 64, iters=3, splits=2, persist=1),
    "t_9472_persist_h2": lambda: exp_time(2, 1, 9472, 4096, 131072, 64, iters=3, splits=2, persist=1, hints=2),
    "t_4736_persist_s4": lambda: exp_time(2, 1, 4736, 4096, 131072, 64, iters=3, splits=4, persist=1),
    "time_dbg_a": lambda: exp_time(2, 1, 65536, 4096, 131072, 64, iters=2, dbg=1),
    "time_dbg_b": lambda: exp_time(2, 1, 65536, 4096, 131072, 64, iters=2, dbg=2),
    "time_dbg_ab": lambda: exp_time(2, 1, 65536, 4096, 131072, 64, iters=2, dbg=3),
    "time_dbg_0": lambda: exp_time(2, 1, 65536, 4096, 131072, 64, iters=2, dbg=0),
    "time_refine_64k_s3": lambda: exp_time(2, 3, 65536, 4096, 131072, 64, iters=2, splits=3),
    "time_refine_64k_s4": lambda: exp_time(2, 3, 65536, 4096, 131072, 64, iters=2, splits=4),
    "time_refine_64k_s6": lambda: exp_time(2, 3, 65536, 4096, 131072, 64, iters=2, splits=6),
    "time_refine_64k_s8": lambda: exp_time(2, 3, 65536, 4096, 131072, 64, iters=2, splits=8),
    "steering_32k": lambda: exp_steering(),
    "cache_64k": lambda: exp_cache(),
    "scan_sim8": lambda: exp_scan_sim(),
    "scan_w1": lambda: exp_scan(world=1),
    "scan_w8": lambda: exp_scan(world=8),
    "stats_refine": lambda: exp_stats(3),
    "stats_hilo": lambda: exp_stats(2),
    "stats_bf16x1": lambda: exp_stats(1),
    "m32": lambda: exp_overlap(65536, 18944, margin=32),
    "m48": lambda: exp_overlap(65536, 18944, margin=48),
    "m64": lambda: exp_overlap(65536, 18944, margin=64),
    "m24": lambda: exp_overlap(65536, 18944, margin=24),
    "seq3_refine": lambda: exp_overlap(65536, 18944, overlap=False),
    "ov3_c18944": lambda: exp_overlap(65536, 18944),
    "ov3_c9472": lambda: exp_overlap(65536, 9472),
    "ov2_c9472": lambda: exp_overlap(65536, 9472),
    "ov2_c18944": lambda: exp_overlap(65536, 18944),
    "ov2_c28416": lambda: exp_overlap(65536, 28416),
    "ov2_c18944_nopersist": lambda: exp_overlap(65536, 18944, persist=0),
    "seq2_refine": lambda: exp_overlap(65536, 18944, overlap=False),
    "seq2_refine_nochunk": lambda: exp_overlap(65536, 18944, overlap=False, chunking=0),
    "seq2_hilo": lambda: exp_overlap(65536, 18944, planes=2, overlap=False),
    "seq2_hilo_nochunk": lambda: exp_overlap(65536, 18944, planes=2, overlap=False, chunking=0),
    "ov_s2_c18944": lambda: exp_overlap(65536, 18944, splits=2),
    "ov_s3_c18944": lambda: exp_overlap(65536, 18944, splits=3),
    "ov_s4_c18944": lambda: exp_overlap(65536, 18944, splits=4),
    "ov_s4_c9472": lambda: exp_overlap(65536, 9472, splits=4),
    "ov_s2_c9472": lambda: exp_overlap(65536, 9472, splits=2),
    "ov_s6_c18944": lambda: exp_overlap(65536, 18944, splits=6),
    "seq_s4": lambda: exp_overlap(65536, 18944, splits=4, overlap=False),
    "ov_64k_c8k": lambda: exp_overlap(65536, 8192),
    "ov_64k_c16k": lambda: exp_overlap(65536, 16384),
    "ov_64k_c4k": lambda: exp_overlap(65536, 4096),
    "ov_64k_c8k_nohint": lambda: exp_overlap(65536, 8192, l2_hints=0),
    "ov_64k_c8k_hint3": lambda: exp_overlap(65536, 8192, l2_hints=3),
    "seq_64k_refine": lambda: exp_overlap(65536, 8192, overlap=False),
    "seq_64k_refine_nohint": lambda: exp_overlap(65536, 8192, overlap=False, l2_hints=0),
    "seq_64k_hilo": lambda: exp_overlap(65536, 8192, planes=2, overlap=False),
    "seq_64k_hilo_nohint": lambda: exp_overlap(65536, 8192, planes=2, overlap=False, l2_hints=0),
    "time_bf16x1_64k": lambda: exp_time(2, 1, 65536, 4096, 131072, 64, iters=2),
    "time_hilo_64k_s2": lambda: exp_time(2, 2, 65536, 4096, 131072, 64, iters=2, splits=2),
    "time_hilo_64k_s4": lambda: exp_time(2, 2, 65536, 4096, 131072, 64, iters=2, splits=4),
    "time_hilo_64k_s8": lambda: exp_time(2, 2, 65536, 4096, 131072, 64, iters=2, splits=8),
}

if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--child":
        name = sys.argv[2]
        t = time.time()
        res = EXPS[name]()
        res["wall_s"] = round(time.time() - t, 2)
        print("RESULT " + json.dumps({name: res}))
        sys.exit(0)
    names = sys.argv[1:] or list(EXPS)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "probe.log"), "a")
    for n in names:
        try:
            r = subprocess.run([sys.executable, __file__, "--child", n], capture_output=True, text=True, timeout=240)
            lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            msg = lines[-1] if lines else f"FAIL {n} rc={r.returncode}\n--stdout--\n{r.stdout[-1500:]}\n--stderr--\n{r.stderr[-2500:]}"
        except subprocess.TimeoutExpired as e:
            msg = f"TIMEOUT {n}\n{(e.stdout or b'')[-1000:]}\n{(e.stderr or b'')[-1000:]}"
        print(msg, flush=True)
        log.write(msg + "\n")
        log.flush()
