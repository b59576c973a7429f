"""Round-2 probe: does the HBM-bound half of the forward (refine + decode) really run INSIDE the tensor-core-bound GEMM
launches?  Each experiment runs in its own process (timeout) and prints one RESULT line; results -> gpurun_out/.

    python tools/probe_overlap.py [exp ...]
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-sae_b200"))

T, D, N, K = 65536, 4096, 131072, 64
VALUE_MODE = int(os.environ.get("PROBE_VALUE_MODE", "0"))


def _common(stages=0, planes=3):
    import torch
    from saeb200 import _capi, engine, synth
    L = _capi.lib()
    _capi.check(L.saeb_set_option(b"gemm_stages", stages), "set_option")
    sae = synth.make_sae(D, N, K, "cuda", seed=1234)
    sae.encoder_planes = planes
    enc = sae.packed_encoder()
    x = synth.make_activations(T, D, "cuda", seed=3)
    return torch, _capi, engine, L, sae, enc, x


def _time(torch, fn, iters=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms.append(round(e0.elapsed_time(e1), 3))
    return ms


def exp_forward(overlap=True, chunk=9472, ctas_per_sm=1, priority="gemm", stages=0, planes=3, check=True):
    torch, _capi, engine, L, sae, enc, x = _common(stages, planes)
    from saeb200.overlap import OverlappedForward
    acts = torch.empty((T, K), dtype=torch.float32, device="cuda")
    idx = torch.empty((T, K), dtype=torch.int64, device="cuda")
    out = torch.empty((T, D), dtype=torch.float32, device="cuda")
    sq = torch.zeros((), dtype=torch.float64, device="cuda")
    ov = OverlappedForward(enc, sae.W_dec.data, sae.b_dec.data, K, chunk=chunk, ctas_per_sm=ctas_per_sm,
                           priority=priority, value_mode=VALUE_MODE) if overlap else None
    if ov is not None:
        ov.gemm_stages = stages

    def step():
        sq.zero_()
        if ov is not None:
            ov.run(x, acts, idx, out, sq)
        else:
            engine.encode_topk(x, enc, K, out_vals=acts, out_idx=idx, value_mode=VALUE_MODE)
            engine.decode(idx, acts, sae.W_dec.data, sae.b_dec.data, x=x, sq_err=sq, out=out)
    ms = _time(torch, step)
    res = dict(ms=ms, tokens_per_s=round(T / (min(ms) * 1e-3)), path_tflops=round(T / (min(ms) * 1e-3) * 1.0743e9 / 1e12, 1))
    if check and ov is not None:
        n = 20000
        a2, i2, _ = engine.encode_topk(x[:n], enc, K, value_mode=VALUE_MODE)
        o2 = engine.decode(i2, a2, sae.W_dec.data, sae.b_dec.data)
        res["equals_sequential"] = bool(torch.equal(a2, acts[:n]) and torch.equal(i2, idx[:n]) and torch.equal(o2, out[:n]))
        res["flagged"] = int(ov.status.sum().item())
    return res


def exp_parts(stages=0, max_ctas=0, chunk=9472):
    """GEMM launches alone, gathers alone (bounded grid or not), and both at once WITHOUT data dependencies (the
    gathers work on the candidate lists of a previous pass): separates scheduling / co-residency from pipeline logic"""
    torch, _capi, engine, L, sae, enc, x = _common(stages)
    check = _capi.check
    n_chunks = (T + chunk - 1) // chunk
    prep = torch.empty(L.saeb_prep_bytes(T, D), dtype=torch.uint8, device="cuda")
    wsb = L.saeb_candidates_workspace_bytes(chunk, D, N, K, 0)
    ws_a = [torch.empty(wsb, dtype=torch.uint8, device="cuda") for _ in range(n_chunks)]   # read by the gathers
    ws_b = [torch.empty(wsb, dtype=torch.uint8, device="cuda") for _ in range(2)]          # written by the timed GEMMs
    acts = torch.empty((T, K), dtype=torch.float32, device="cuda")
    idx = torch.empty((T, K), dtype=torch.int64, device="cuda")
    out = torch.empty((T, D), dtype=torch.float32, device="cuda")
    sq = torch.zeros((), dtype=torch.float64, device="cuda")
    status = torch.zeros(4, dtype=torch.int32, device="cuda")
    s_g = torch.cuda.Stream(priority=-1)
    s_m = torch.cuda.Stream(priority=0)
    code = engine._code(x)
    check(L.saeb_prep_activations(x.data_ptr(), code, T, D, D, prep.data_ptr(), torch.cuda.current_stream().cuda_stream), "prep")

    def gemm(c, ws, st):
        a, b = c * chunk, min(T, (c + 1) * chunk)
        check(L.saeb_encode_candidates(prep.data_ptr(), T, a, b - a, enc.blob.data_ptr(), D, N, K, 0, -1, 0.0,
                                       ws.data_ptr(), ws.numel(), st.cuda_stream), "gemm")

    def gather(c, ws, st, mc):
        a, b = c * chunk, min(T, (c + 1) * chunk)
        with torch.cuda.stream(st):
            check(L.saeb_refine_candidates(x.data_ptr() + a * D * 2, code, D, prep.data_ptr(), T, a, b - a,
                                           enc.blob.data_ptr(), enc.W_enc.data_ptr(), D, N, K, 0, -1, 0.0, None, None, None,
                                           0, acts[a:b].data_ptr(), None, idx[a:b].data_ptr(), status.data_ptr(), ws.data_ptr(),
                                           ws.numel(), mc, VALUE_MODE, st.cuda_stream), "refine")
            engine.decode(idx[a:b], acts[a:b], sae.W_dec.data, sae.b_dec.data, x=x[a:b], sq_err=sq, out=out[a:b],
                          max_ctas=mc)

    main = torch.cuda.current_stream()
    for c in range(n_chunks):   # candidate lists for the gathers
        gemm(c, ws_a[c], main)
    torch.cuda.synchronize()

    def only_gemm():
        s_g.wait_stream(main)
        for c in range(n_chunks):
            gemm(c, ws_b[c & 1], s_g)
        main.wait_stream(s_g)

    def only_gather():
        s_m.wait_stream(main)
        for c in range(n_chunks):
            gather(c, ws_a[c], s_m, max_ctas)
        main.wait_stream(s_m)

    def both():
        s_g.wait_stream(main)
        s_m.wait_stream(main)
        for c in range(n_chunks):
            gemm(c, ws_b[c & 1], s_g)
            gather(c, ws_a[c], s_m, max_ctas)
        main.wait_stream(s_g)
        main.wait_stream(s_m)

    r = dict(gemm_ms=_time(torch, only_gemm), gather_ms=_time(torch, only_gather), both_ms=_time(torch, both))
    r["sum_ms"] = round(min(r["gemm_ms"]) + min(r["gather_ms"]), 2)
    r["overlap_gain"] = round(1.0 - min(r["both_ms"]) / r["sum_ms"], 3)
    return r


def exp_power(seconds=3.0):
    """board power (nvidia-smi, 100 ms samples) while looping the GEMM launches alone, the gathers alone and both: is
    the step energy-bound under the power cap?"""
    import threading
    torch, _capi, engine, L, sae, enc, x = _common(0)
    check = _capi.check
    chunk = 9472
    n_chunks = (T + chunk - 1) // chunk
    prep = torch.empty(L.saeb_prep_bytes(T, D), dtype=torch.uint8, device="cuda")
    wsb = L.saeb_candidates_workspace_bytes(chunk, D, N, K, 0)
    ws = [torch.empty(wsb, dtype=torch.uint8, device="cuda") for _ in range(n_chunks)]
    acts = torch.empty((T, K), dtype=torch.float32, device="cuda")
    idx = torch.empty((T, K), dtype=torch.int64, device="cuda")
    out = torch.empty((T, D), dtype=torch.float32, device="cuda")
    status = torch.zeros(4, dtype=torch.int32, device="cuda")
    code = engine._code(x)
    st = torch.cuda.current_stream()
    check(L.saeb_prep_activations(x.data_ptr(), code, T, D, D, prep.data_ptr(), st.cuda_stream), "prep")

    def gemm():
        for c in range(n_chunks):
            a, b = c * chunk, min(T, (c + 1) * chunk)
            check(L.saeb_encode_candidates(prep.data_ptr(), T, a, b - a, enc.blob.data_ptr(), D, N, K, 0, -1, 0.0,
                                           ws[c].data_ptr(), ws[c].numel(), st.cuda_stream), "gemm")

    def refine(vm):
        for c in range(n_chunks):
            a, b = c * chunk, min(T, (c + 1) * chunk)
            check(L.saeb_refine_candidates(x.data_ptr() + a * D * 2, code, D, prep.data_ptr(), T, a, b - a,
                                           enc.blob.data_ptr(), enc.W_enc.data_ptr(), D, N, K, 0, -1, 0.0, None, None, None,
                                           0, acts[a:b].data_ptr(), None, idx[a:b].data_ptr(), status.data_ptr(),
                                           ws[c].data_ptr(), ws[c].numel(), 0, vm, st.cuda_stream), "refine")

    def decode():
        engine.decode(idx, acts, sae.W_dec.data, sae.b_dec.data, out=out)

    gemm(); refine(0); decode(); torch.cuda.synchronize()
    res = {}
    for name, fn in (("idle", None), ("gemm", gemm), ("refine_exact", lambda: refine(0)),
                     ("refine_boundary", lambda: refine(1)), ("decode", decode)):
        samples, stop = [], threading.Event()

        def pump():
            import subprocess as sp
            pr = sp.Popen(["nvidia-smi", "--query-gpu=power.draw,clocks.sm", "--format=csv,noheader,nounits", "-lms", "100"],
                          stdout=sp.PIPE, text=True)
            for line in pr.stdout:
                if stop.is_set():
                    break
                try:
                    pw, ck = line.split(",")
                    samples.append((time.time(), float(pw), float(ck)))
                except Exception:
                    pass
            pr.terminate()
        th = threading.Thread(target=pump, daemon=True); th.start()
        time.sleep(0.5)
        t0 = time.time()
        n = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        while time.time() - t0 < seconds:
            if fn is None:
                time.sleep(0.1)
            else:
                fn(); n += 1
                if n % 4 == 0:
                    torch.cuda.synchronize()
        e1.record(); torch.cuda.synchronize()
        t1 = time.time()
        stop.set()
        # nvidia-smi power.draw is a ~1 s moving average: use the second half of the window
        tail = [(p_, c_) for ts, p_, c_ in samples if t0 + 0.6 * (t1 - t0) <= ts <= t1]
        res[name] = dict(ms_per_pass=round(e0.elapsed_time(e1) / max(n, 1), 2), passes=n,
                         power_w=round(sum(p_ for p_, _ in tail) / max(len(tail), 1), 1),
                         sm_mhz=round(sum(c_ for _, c_ in tail) / max(len(tail), 1)))
    return res


def exp_gemm_only(cluster4=1, stages=0):
    """the 7 GEMM launches of a 65 536-token step alone, CTA-pair kernel vs 4-CTA clusters with activation multicast"""
    torch, _capi, engine, L, sae, enc, x = _common(stages)
    check = _capi.check
    check(L.saeb_set_option(b"cluster4", cluster4), "cluster4")
    chunk = 9472
    n_chunks = (T + chunk - 1) // chunk
    prep = torch.empty(L.saeb_prep_bytes(T, D), dtype=torch.uint8, device="cuda")
    wsb = L.saeb_candidates_workspace_bytes(chunk, D, N, K, 0)
    ws = [torch.empty(wsb, dtype=torch.uint8, device="cuda") for _ in range(2)]
    st = torch.cuda.current_stream()
    check(L.saeb_prep_activations(x.data_ptr(), engine._code(x), T, D, D, prep.data_ptr(), st.cuda_stream), "prep")

    def gemm():
        for c in range(n_chunks):
            a, b = c * chunk, min(T, (c + 1) * chunk)
            check(L.saeb_encode_candidates(prep.data_ptr(), T, a, b - a, enc.blob.data_ptr(), D, N, K, 0, -1, 0.0,
                                           ws[c & 1].data_ptr(), ws[c & 1].numel(), st.cuda_stream), "gemm")
    ms = _time(torch, gemm, iters=4)
    return dict(ms=ms, tflops=round(2.0 * T * D * N / (min(ms) * 1e-3) / 1e12, 1),
                max_clusters4=int(L.saeb_query(b"max_clusters4")))


def exp_shard_phases(world=8, rank=3, tokens=1048576 // 4, waves=4):
    """where does the GEMM stream of ONE rank of the feature-sharded scan spend its time?  prep / GEMM launches /
    merge + bounds timed separately (CUDA events, summed over the chunks) for the shard N / world on one GPU"""
    import torch
    from saeb200 import _capi, dist as sdist, engine, synth
    L = _capi.lib()
    check = _capi.check
    sae = synth.make_sae(D, N, K, "cuda", seed=1234)
    lo, hi = sdist.shard_range(N, world, rank)
    enc = engine.PackedEncoder.pack(sae.encoder.weight.data[lo:hi], sae.encoder.bias.data[lo:hi], sae.b_dec.data, 3)
    Ns = hi - lo
    chunk = waves * 9472
    x = synth.make_activations(tokens, D, "cuda", seed=5)
    prep = torch.empty(L.saeb_prep_bytes(chunk, D), dtype=torch.uint8, device="cuda")
    ws = torch.empty(L.saeb_candidates_workspace_bytes(chunk, D, Ns, K, 0), dtype=torch.uint8, device="cuda")
    lb = torch.empty((chunk, K), dtype=torch.float32, device="cuda")
    ub = torch.empty((chunk, K), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    acc = {"prep": 0.0, "gemm": 0.0, "merge+bounds": 0.0}
    n_chunks = 0
    for rep in range(2):
        evs = []
        for t0 in range(0, tokens - chunk + 1, chunk):
            xc = x[t0:t0 + chunk]
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            e[0].record()
            check(L.saeb_prep_activations(xc.data_ptr(), _capi.BF16, chunk, D, D, prep.data_ptr(), st), "prep")
            e[1].record()
            check(L.saeb_encode_candidates(prep.data_ptr(), chunk, 0, chunk, enc.blob.data_ptr(), D, Ns, K, 0, -1, 0.0,
                                           ws.data_ptr(), ws.numel(), st), "gemm")
            e[2].record()
            check(L.saeb_candidate_bounds(prep.data_ptr(), chunk, 0, chunk, enc.blob.data_ptr(), _capi.BF16, D, Ns, K, 0,
                                          -1, lb.data_ptr(), ub.data_ptr(), ws.data_ptr(), ws.numel(), 0, st), "bounds")
            e[3].record()
            evs.append(e)
        torch.cuda.synchronize()
        if rep == 1:
            for e in evs:
                acc["prep"] += e[0].elapsed_time(e[1])
                acc["gemm"] += e[1].elapsed_time(e[2])
                acc["merge+bounds"] += e[2].elapsed_time(e[3])
            n_chunks = len(evs)
    flops = 2.0 * n_chunks * chunk * D * Ns
    return dict(world=world, features=Ns, chunk_tokens=chunk, chunks=n_chunks,
                ms_per_chunk={k_: round(v_ / n_chunks, 3) for k_, v_ in acc.items()},
                gemm_tflops=round(flops / (acc["gemm"] * 1e-3) / 1e12, 1),
                per_1M_tokens_ms={k_: round(v_ / (n_chunks * chunk) * 1048576, 1) for k_, v_ in acc.items()})


def exp_coload(world=8, rank=3, n_chunks=8, waves=4):
    """the GEMM stream of one rank of the 8-way sharded scan (prep + 4 single-wave launches per 37 888-token chunk)
    alone and with a synthetic co-resident load on a second stream: pure ALU, streaming reads, L2-resident reads, at two
    intensities each.  Which resource does a co-resident chain take from the GEMM?"""
    import torch
    from saeb200 import _capi, dist as sdist, engine, synth
    L = _capi.lib()
    check = _capi.check
    sae = synth.make_sae(D, N, K, "cuda", seed=1234)
    lo, hi = sdist.shard_range(N, world, rank)
    enc = engine.PackedEncoder.pack(sae.encoder.weight.data[lo:hi], sae.encoder.bias.data[lo:hi], sae.b_dec.data, 3)
    Ns = hi - lo
    chunk = waves * 9472
    x = synth.make_activations(chunk * n_chunks, D, "cuda", seed=5)
    prep = [torch.empty(L.saeb_prep_bytes(chunk, D), dtype=torch.uint8, device="cuda") for _ in range(2)]
    ws = [torch.empty(L.saeb_candidates_workspace_bytes(chunk, D, Ns, K, 0), dtype=torch.uint8, device="cuda") for _ in range(2)]
    buf = torch.empty(2 << 30, dtype=torch.uint8, device="cuda")
    sink = torch.zeros(1, dtype=torch.float32, device="cuda")
    sg, sa = torch.cuda.Stream(priority=0), torch.cuda.Stream(priority=-1)
    check(L.saeb_set_option(b"gemm_stages", 5), "stages")

    def run(load):
        main = torch.cuda.current_stream()
        sg.wait_stream(main); sa.wait_stream(main)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(sg)
        for c in range(n_chunks):
            xc = x[c * chunk:(c + 1) * chunk]
            st = sg.cuda_stream
            check(L.saeb_prep_activations(xc.data_ptr(), _capi.BF16, chunk, D, D, prep[c & 1].data_ptr(), st), "prep")
            check(L.saeb_encode_candidates(prep[c & 1].data_ptr(), chunk, 0, chunk, enc.blob.data_ptr(), D, Ns, K, 0, -1, 0.0,
                                           ws[c & 1].data_ptr(), ws[c & 1].numel(), st), "gemm")
        e1.record(sg)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if load is not None:
            mode, ctas, iters, reps = load
            a0.record(sa)
            for _ in range(reps):
                check(L.saeb_debug_coload(mode, ctas, iters, buf.data_ptr(), buf.numel(), sink.data_ptr(), sa.cuda_stream), "coload")
            a1.record(sa)
        main.wait_stream(sg); main.wait_stream(sa)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n_chunks, (a0.elapsed_time(a1) if load is not None else 0.0)

    res = {}
    run(None)
    base = min(run(None)[0] for _ in range(3))
    res["gemm_alone_ms_per_chunk"] = round(base, 3)
    # (mode, ctas, iters, launches): sized so that the load lasts about as long as the GEMM stream
    loads = {"alu_2cta": (0, 296, 400000, 40), "alu_4cta": (0, 592, 400000, 40),
             "stream_light": (1, 296, 2000, 40), "stream_heavy": (1, 592, 6000, 40),
             "l2_light": (2, 296, 2000, 40), "l2_heavy": (2, 592, 6000, 40)}
    if os.environ.get("PROBE_COLOAD_SMSP"):
        # pure ALU load from ONE warp position of every 4-warp CTA (CTA-local warp id, then hardware warp slot):
        # the GEMM's TMA producer is CTA warp 0, its MMA issuer warp 1, the epilogue warps 4-7
        loads = {f"alu_2cta_warp{w}": (0 | ((1 << w) << 4), 296, 400000, 40) for w in range(4)}
        loads.update({f"alu_2cta_hwslot{w}": (0 | ((1 << w) << 4) | 256, 296, 400000, 40) for w in range(4)})
        loads["alu_2cta_warps23"] = (0 | (12 << 4), 296, 400000, 40)
        loads["alu_4cta_warps23"] = (0 | (12 << 4), 592, 400000, 40)
        loads["alu_2cta_warps01"] = (0 | (3 << 4), 296, 400000, 40)
    for name, ld in loads.items():
        alone = run_load_alone = None
        # the load alone (duration, and bytes per second for the memory modes)
        main = torch.cuda.current_stream()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(ld[3]):
            check(L.saeb_debug_coload(ld[0], ld[1], ld[2], buf.data_ptr(), buf.numel(), sink.data_ptr(), main.cuda_stream), "coload")
        a1.record(); torch.cuda.synchronize()
        alone = a0.elapsed_time(a1)
        g, a = run(ld)
        gb = ld[1] * 128 * ld[2] * 16 * ld[3] / 1e9 if ld[0] else 0.0
        res[name] = dict(gemm_ms_per_chunk=round(g, 3), slowdown=round(g / base - 1, 3), load_ms_beside=round(a, 1),
                         load_ms_alone=round(alone, 1), load_gb=round(gb, 1),
                         load_gbs_beside=round(gb / (a * 1e-3), 0) if a else None)
    check(L.saeb_set_option(b"gemm_stages", 0), "stages")
    return res


EXPS = {
    "coload": lambda: exp_coload(),
    "ov0": lambda: exp_forward(ctas_per_sm=0),
    "ov0_s5": lambda: exp_forward(ctas_per_sm=0, stages=5),
    "ov4_s5": lambda: exp_forward(ctas_per_sm=4, stages=5),
    "shard8": lambda: exp_shard_phases(8, 3),
    "shard8_w8": lambda: exp_shard_phases(8, 3, waves=8),
    "shard1": lambda: exp_shard_phases(1, 0),
    "gemm_pair": lambda: exp_gemm_only(0),
    "gemm_cl4": lambda: exp_gemm_only(2),
    "gemm_cl4_auto": lambda: exp_gemm_only(1),
    "gemm_cl4_s5": lambda: exp_gemm_only(2, 5),
    "power": lambda: exp_power(),
    "seq": lambda: exp_forward(overlap=False),
    "ov_r1": lambda: exp_forward(chunk=18944, ctas_per_sm=0, priority="mem"),
    "ov_c0_pgemm": lambda: exp_forward(chunk=18944, ctas_per_sm=0, priority="gemm"),
    "ov1": lambda: exp_forward(ctas_per_sm=1),
    "ov1_c18944": lambda: exp_forward(chunk=18944, ctas_per_sm=1),
    "ov1_pnone": lambda: exp_forward(ctas_per_sm=1, priority="none"),
    "ov1_s5": lambda: exp_forward(ctas_per_sm=1, stages=5),
    "ov2_s5": lambda: exp_forward(ctas_per_sm=2, stages=5),
    "ov3_s5": lambda: exp_forward(ctas_per_sm=3, stages=5),
    "ov2_s5_c18944": lambda: exp_forward(chunk=18944, ctas_per_sm=2, stages=5),
    "ov4_s4": lambda: exp_forward(ctas_per_sm=4, stages=4),
    "ov2_s5_m4": lambda: exp_forward(ctas_per_sm=2, stages=5, planes=4),
    "parts_s6_full": lambda: exp_parts(0, 0),
    "parts_s6_c1": lambda: exp_parts(0, 148),
    "parts_s5_c1": lambda: exp_parts(5, 148),
    "parts_s5_c2": lambda: exp_parts(5, 296),
    "parts_s5_c3": lambda: exp_parts(5, 444),
    "parts_s4_c4": lambda: exp_parts(4, 592),
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
    log = open(os.path.join(ROOT, "gpurun_out", "probe_overlap.log"), "a")
    for n in names:
        try:
            r = subprocess.run([sys.executable, __file__, "--child", n], capture_output=True, text=True, timeout=200)
            lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            msg = lines[-1] if lines else f"FAIL {n} rc={r.returncode}\n--stdout--\n{r.stdout[-1500:]}\n--stderr--\n{r.stderr[-2500:]}"
        except subprocess.TimeoutExpired as e:
            msg = f"TIMEOUT {n}\n{(e.stdout or b'')[-1000:]}\n{(e.stderr or b'')[-1000:]}"
        print(msg, flush=True)
        log.write(msg + "\n")
        log.flush()
