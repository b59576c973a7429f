"""ONE rank of the 8-way feature-sharded scan on ONE GPU, with what the peers would send replayed from a recording.

Multi-GPU calls cost 8x the box time, and what limits the sharded scan is the per-rank schedule (does the per-chunk
chain -- merge + bounds, exchange, kth, restricted refinement, exchange, kth, list update -- really run INSIDE the next
chunk's GEMM launches?), not the NVLink transfers (0.02-0.04 ms per exchange).  So:

  phase A  all R logical shards run the real protocol in lockstep on this GPU over C chunks (the choreography of
           tests/test_gpu_parity.py::test_feature_sharded_scan_logical_shards) and the gathered exchange tensors of
           every chunk are recorded ([R, Tc, 2 m1] bounds, [R, Tc, k] member values);
  phase B  rank r alone runs `saeb200.dist.sharded_scan` (the production loop, both streams) with a replay transport:
           an exchange = copy of the rank's live slab into the recorded gathered tensor (a small copy kernel instead of
           the push kernel).  Its per-feature lists must equal phase A's lists of shard r.

Phase B records CUDA events around every operation on both streams (no synchronisation) and prints the timeline of a
few chunks: when each chain kernel ran relative to the GEMM launches of the next chunk.

    python tools/scan_rank_emul.py [--world 8] [--rank 3] [--chunks 8] [--stages 0] [--refine-ctas 0] [--prio high]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-sae_b200"))

D, N, K, CTX = 4096, 131072, 64, 64


class FakeDist:
    """just enough of torch.distributed for saeb200.dist.sharded_scan to take its world > 1 path in one process"""

    class ReduceOp:
        MIN, MAX, SUM = "min", "max", "sum"

    def __init__(self, world, rank):
        self.world, self.rank = world, rank

    def is_available(self):
        return True

    def is_initialized(self):
        return True

    def get_world_size(self, group=None):
        return self.world

    def get_rank(self, group=None):
        return self.rank

    def all_reduce(self, t, op=None, group=None):
        return None

    def all_gather(self, out, t, group=None, async_op=False):
        for o in out:   # final list all-gather: every slab = this rank's lists (only the own slab is looked at)
            o.copy_(t)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--rank", type=int, default=3)
    ap.add_argument("--chunks", type=int, default=8)
    ap.add_argument("--waves", type=int, default=4)
    ap.add_argument("--top", type=int, default=20)
    ap.add_argument("--stages", default="0", help="comma list of GEMM ring depths to time (0 = default)")
    ap.add_argument("--refine-ctas", default="0", help="comma list: bounded refinement grid, CTAs per SM (0 = one per token)")
    ap.add_argument("--prio", default="high", help="comma list of aux-stream priorities: high,low")
    ap.add_argument("--timeline-chunks", type=int, default=2)
    ap.add_argument("--tag", default="")
    ap.add_argument("--shard-margin", type=int, default=-1, help="EngineOps.margin_sharded (-1 = its default)")
    ap.add_argument("--margin", type=int, default=0, help="option refine_margin: candidates kept per row = k + margin "
                    "(0 = default max(48, k/2)); smaller lists are safe (short lists are flagged and recomputed exactly)")
    ap.add_argument("--prep-ahead", type=int, default=0, help="0: activation prep on the GEMM stream")
    ap.add_argument("--sequential", type=int, default=0, help="1: one stream, phases back to back (every span = the "
                    "kernel's time with the GPU to itself)")
    ap.add_argument("--fine", type=int, default=0, help="1: also time the individual library calls of update / refinement")
    ap.add_argument("--skip", default="", help="semicolon list of configurations, each a comma list of chain steps that "
                    "are NOT launched (their recorded outputs are used instead): bounds,gbounds,topk,kth,update,exchange; "
                    "'all' = GEMM stream only.  Marginal cost of each step on the GEMM")
    ap.add_argument("--coresident", type=int, default=1, help="0: round-1 launch shapes of the chain")
    ap.add_argument("--packed-bounds", type=int, default=1, help="0: merge + candidate_bounds + cat instead of the fused kernel")
    ap.add_argument("--scan-warp", type=int, default=1, help="0: CTA-per-token refinement")
    args = ap.parse_args()

    import torch

    from saeb200 import _capi, dist as sdist, engine, synth

    L = _capi.lib()
    _capi.check(L.saeb_set_option(b"refine_margin", args.margin), "refine_margin")
    dev = torch.device("cuda", 0)
    R, r = args.world, args.rank
    sae = synth.make_sae(D, N, K, dev, seed=1234)
    shards = [sdist.shard_range(N, R, i) for i in range(R)]
    num_sms = int(L.saeb_query(b"num_sms"))

    def make_ops(i, prio="high"):
        lo, hi = shards[i]
        return sdist.EngineOps(sae.encoder.weight.data[lo:hi], sae.encoder.bias.data[lo:hi], sae.b_dec.data, lo, hi,
                               args.top, CTX, dev, aux_priority=prio)

    ops_all = [make_ops(i) for i in range(R)]
    chunk = ops_all[0].chunk_tokens(R, args.waves)
    tokens = chunk * args.chunks
    xs = synth.make_activations(tokens, D, dev, seed=99)
    k_local = min(K, shards[0][1] - shards[0][0])
    m1 = sdist.bounds_width(K, k_local, R)
    m1 = min(k_local, max(m1, -(-(K + 1) // R)))

    # ---- phase A: lockstep protocol over all logical shards, exchanges recorded
    G1, G2, REC = [], [], []
    for c in range(args.chunks):
        xc = xs[c * chunk:(c + 1) * chunk]
        bounds = [o.local_bounds(xc, k_local) for o in ops_all]
        g1 = torch.stack([torch.cat([lb[:, :m1], ub[:, :m1]], -1) for lb, ub in bounds], 0)
        ext_L = engine.kth_of_gathered(g1[:, :, :m1], K)
        ext_U = torch.maximum(engine.kth_of_gathered(g1[:, :, m1:], K + 1), g1[:, :, m1:].amin(-1).amax(0))
        outs = [o.local_topk(ext_L, ext_U) for o in ops_all]
        g2 = torch.stack([m for _, m, _ in outs], 0)
        tok_thr = engine.kth_of_gathered(g2, K)
        for o, (v, m, i) in zip(ops_all, outs):
            o.scan_update(v, i, c * chunk // CTX, tok_thr, m)
        G1.append(g1)
        G2.append(g2)
        REC.append(dict(ext=(ext_L.clone(), ext_U.clone()), topk=tuple(t_.clone() for t_ in outs[r]),
                        tok_thr=tok_thr.clone()))
    ref_vals, ref_win = ops_all[r].scan_finalize()
    ref_vals, ref_win = ref_vals.clone(), ref_win.clone()
    del ops_all
    torch.cuda.empty_cache()

    # ---- phase B: rank r through the production loop with the replay transport
    class ReplayOps(sdist.EngineOps):
        def __init__(self, *a, **kw):
            super().__init__(*a, **kw)
            self.exchange = "push"
            self.count = [0, 0]
            self.timeline = []

        def push_gather(self, t, group, channel, slot):
            c = self.count[channel]
            self.count[channel] += 1
            g = (G1, G2)[channel][c]
            if "exchange" not in self.skip:
                with self.span(f"exchange{channel + 1}", c):
                    g[r].copy_(t)   # the live slab: keeps the data dependency of the real exchange
            return g

        skip = frozenset()

        def local_bounds_finish(self, slot=0, coresident=False, pack_m1=None):
            if "bounds" in self.skip:   # views of the slot's (stale) bound lists: right shapes, nothing launched
                T = self._x[slot].shape[0]
                k = self._k[slot]
                lb = self._scratch(self._lb, slot, T * k * 4, dev)[: T * k * 4].view(torch.float32).view(T, k)
                ub = self._scratch(self._ub, slot, T * k * 4, dev)[: T * k * 4].view(torch.float32).view(T, k)
                return lb, ub
            return super().local_bounds_finish(slot, coresident, pack_m1)

        def gathered_bounds(self, gathered, m1, k):
            if "gbounds" in self.skip:
                return REC[self.count[0] - 1]["ext"]
            return super().gathered_bounds(gathered, m1, k)

        def local_topk(self, ext_L=None, ext_U=None, slot=0):
            if "topk" in self.skip:
                return REC[self.count[0] - 1]["topk"]
            return super().local_topk(ext_L, ext_U, slot)

        def kth_of_gathered(self, gathered, kth=None):
            if "kth" in self.skip:
                return REC[self.count[1] - 1]["tok_thr"]
            return super().kth_of_gathered(gathered, kth)

        def scan_update(self, vals, idx, window_base, tok_thr, member=None):
            if "update" in self.skip:
                return
            return super().scan_update(vals, idx, window_base, tok_thr, member)

        def span(self, name, c=None):
            ops = self

            class _S:
                def __enter__(s):
                    s.e0 = torch.cuda.Event(enable_timing=True)
                    s.e0.record()

                def __exit__(s, *exc):
                    e1 = torch.cuda.Event(enable_timing=True)
                    e1.record()
                    ops.timeline.append((name, c, s.e0, e1))
            return _S()

    def wrap(ops, name, key=None):
        fn = getattr(ops, name)
        cnt = [0]

        def f(*a, **kw):
            c = cnt[0]
            cnt[0] += 1
            with ops.span(key or name, c):
                return fn(*a, **kw)
        setattr(ops, name, f)

    def chunks_iter(n):
        for c in range(n):
            yield xs[c * chunk:(c + 1) * chunk]

    fake = FakeDist(R, r)
    real_dist = sdist.dist
    results = []
    ALL = "bounds,gbounds,topk,kth,update,exchange"
    for skip in [frozenset((ALL if sk == "all" else sk).split(",")) - {""} for sk in args.skip.split(";")]:
      for stages in [int(s) for s in args.stages.split(",")]:
        for ctas in [int(s) for s in args.refine_ctas.split(",")]:
            for prio in args.prio.split(","):
                lo, hi = shards[r]
                ops = ReplayOps(sae.encoder.weight.data[lo:hi], sae.encoder.bias.data[lo:hi], sae.b_dec.data, lo, hi,
                                args.top, CTX, dev, aux_priority=prio)
                ops.refine_max_ctas = ctas * num_sms
                ops.skip = skip
                ops.gemm_stages = stages
                ops.coresident = args.coresident != 0
                ops.packed_bounds = args.packed_bounds != 0
                ops.prep_ahead = args.prep_ahead != 0
                if args.shard_margin >= 0:
                    ops.margin_sharded = args.shard_margin
                _capi.check(L.saeb_set_option(b"scan_warp", args.scan_warp), "scan_warp")
                sdist.dist = fake
                try:
                    sdist.sharded_scan(chunks_iter(min(3, args.chunks)), ops, K, CTX, N)   # warm-up
                    ops.scan = engine.TopActivationScan(lo, hi, args.top, CTX, dev)
                    if args.sequential:
                        ops.scan.coresident = ops.coresident
                    ops.count = [0, 0]
                    ops.timeline = []
                    # kth_of_gathered is called 3x per chunk (ext_L, ext_U, tok_thr): one counter covers them
                    for name in ("local_prep", "local_gemm", "local_bounds_finish", "local_topk", "scan_update", "kth_of_gathered",
                                 "gathered_bounds"):
                        wrap(ops, name)
                    if args.fine:   # spans around the individual library calls of the list update / refinement
                        wrap(ops.scan, "flush")
                        ops.scan.span = ops.span
                        for fname in ("saeb_scan_pool_ws", "saeb_scan_pool", "saeb_refine_candidates"):
                            orig = getattr(L, fname)

                            def make(orig=orig, fname=fname):
                                cnt = [0]

                                def f(*a):
                                    cnt[0] += 1
                                    with ops.span(fname, cnt[0] - 1):
                                        return orig(*a)
                                return f
                            setattr(L, fname, make())
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize()
                    e0.record()
                    res = sdist.sharded_scan(chunks_iter(args.chunks), ops, K, CTX, N,
                                             pipelined=False if args.sequential else None)
                    e1.record()
                    torch.cuda.synchronize()
                finally:
                    sdist.dist = real_dist
                ms = e0.elapsed_time(e1)
                same = bool(torch.equal(res.top_win[lo:hi], ref_win) and torch.equal(res.top_vals[lo:hi], ref_vals))
                if skip:
                    same = None
                # seconds per op class, and the timeline of the middle chunks (ms since the start of the timed run)
                per = {}
                for name, c, a, b in ops.timeline:
                    per[name] = per.get(name, 0.0) + a.elapsed_time(b)
                mid = args.chunks // 2
                tl = [(name, c, round(e0.elapsed_time(a), 3), round(e0.elapsed_time(b), 3))
                      for name, c, a, b in ops.timeline]
                tl.sort(key=lambda t: t[2])
                g0 = [t for t in tl if t[0] == "local_gemm" and t[1] == mid]
                g1 = [t for t in tl if t[0] == "local_gemm" and t[1] == mid + args.timeline_chunks]
                if g0 and g1:
                    tl = [t for t in tl if g0[0][2] <= t[2] < g1[0][3]]
                out = {"tag": args.tag, "skipped": sorted(skip), "world": R, "rank": r, "chunks": args.chunks, "chunk_tokens": chunk,
                       "gemm_stages": stages, "refine_ctas_per_sm": ctas, "aux_priority": prio,
                       "sequential": args.sequential, "margin": args.margin, "shard_margin": args.shard_margin, "prep_ahead": args.prep_ahead, "coresident": args.coresident, "scan_warp": args.scan_warp, "packed_bounds": args.packed_bounds,
                       "ms": round(ms, 2), "ms_per_chunk": round(ms / args.chunks, 3),
                       "ms_per_1M_tokens": round(ms / tokens * 1048576, 1),
                       "lists_equal_lockstep": same, "flagged_rows": int(ops.status.item()),
                       "span_ms_per_chunk": {n_: round(v_ / args.chunks, 3) for n_, v_ in per.items()},
                       "timeline": tl}
                print("EMUL " + json.dumps(out), flush=True)
                results.append(out)
                del ops
                torch.cuda.empty_cache()
    _capi.check(L.saeb_set_option(b"scan_warp", 1), "scan_warp")


if __name__ == "__main__":
    main()
