"""Feature-sharded scan, several configurations back to back in ONE process group (multi-GPU calls are expensive):
exchange transport (NCCL all-gathers vs the library's peer-memory push kernel), chunk size, refinement grid shape,
stream priorities.  Every configuration must produce the same per-feature lists as the first one, and the first one
is cross-checked against the token-parallel form on a slice.  Rank 0 prints one JSON line per configuration.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
        tools/scan_sweep.py [--tokens 1048576] [--top 20] [--configs name,name,...]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-sae_b200"))

# name -> (exchange, waves, refine_max_ctas per SM, aux priority, schedule, gemm stages[, packed mode])
CONFIGS = {
    "push_w4_m4": ("push", 4, 0, "high", "streams", 0, 4),
    "nccl_w4_m4": ("nccl", 4, 0, "high", "streams", 0, 4),
    "push_seq_phases_m4": ("push", 4, 0, "high", "sequential", 0, 4),
    "nccl_w4": ("nccl", 4, 0, "high", "streams", 0),
    "nccl_w8": ("nccl", 8, 0, "high", "streams", 0),
    "push_w4": ("push", 4, 0, "high", "streams", 0),
    "push_w8": ("push", 8, 0, "high", "streams", 0),
    "push_w4_low": ("push", 4, 0, "low", "streams", 0),
    "push_w4_low_prepahead": ("push", 4, 0, "low", "streams", 0),
    "push_w4_low_m48": ("push", 4, 0, "low", "streams", 0),
    "push_w4_low_m16": ("push", 4, 0, "low", "streams", 0),
    "push_w4_low_m4": ("push", 4, 0, "low", "streams", 0),
    "nccl_w4_low": ("nccl", 4, 0, "low", "streams", 0),
    "push_w4_s6": ("push", 4, 0, "high", "streams", 6),
    "push_w4_small": ("push", 4, 4, "high", "streams", 0),
    "push_w8_small": ("push", 8, 4, "high", "streams", 0),
    "push_w4_small_low": ("push", 4, 4, "low", "streams", 0),
    "push_w4_small_s5": ("push", 4, 4, "high", "streams", 5),
    "nccl_w4_small": ("nccl", 4, 4, "high", "streams", 0),
    "push_w4_look": ("push", 4, 0, "high", "lookahead", 0),
    "push_w2_small": ("push", 2, 4, "high", "streams", 0),
    "seq_phases": ("nccl", 4, 0, "high", "sequential", 0),
    "push_seq_phases": ("push", 4, 0, "high", "sequential", 0),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=1048576)
    ap.add_argument("--top", type=int, default=20)
    ap.add_argument("--configs", default="nccl_w4,push_w4,push_w4_small,push_w8_small")
    ap.add_argument("--crosscheck-tokens", type=int, default=151552)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    from saeb200 import _capi, dist as sdist, engine, synth

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _capi.lib()
    D, N, K, ctx = 4096, 131072, 64, 64
    sae = synth.make_sae(D, N, K, dev, seed=1234)
    lo, hi = sdist.shard_range(N, world, rank)
    num_sms = int(L.saeb_query(b"num_sms"))
    xs = synth.make_activations(args.tokens, D, dev, seed=99)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ref = None
    for name in args.configs.split(","):
        exch, waves, ctas, prio, sched, stages, *rest = CONFIGS[name]
        planes = rest[0] if rest else 3
        os.environ["SAEB_SCAN_SCHEDULE"] = "streams" if sched == "sequential" else sched
        _capi.check(L.saeb_set_option(b"gemm_stages", stages), "gemm_stages")
        ops = sdist.EngineOps(sae.encoder.weight.data[lo:hi], sae.encoder.bias.data[lo:hi], sae.b_dec.data, lo, hi,
                              args.top, ctx, dev, aux_priority=prio, planes=planes)
        ops.exchange = exch if world > 1 else "nccl"
        ops.refine_max_ctas = ctas * num_sms
        if stages:
            ops.gemm_stages = stages
        ops.prep_ahead = "prepahead" in name
        if "_m48" in name:
            ops.margin_sharded = 0
        if "_m4" in name and "_m48" not in name:
            ops.margin_sharded = 4
        if "_m16" in name:
            ops.margin_sharded = 16
        chunk = ops.chunk_tokens(world, waves)

        def chunks(limit=args.tokens):
            for t0 in range(0, min(args.tokens, limit), chunk):
                yield xs[t0:min(args.tokens, limit, t0 + chunk)]

        out = {"config": name, "world": world, "tokens": args.tokens, "chunk_tokens": chunk, "exchange": exch,
               "refine_max_ctas": ctas * num_sms, "aux_priority": prio, "schedule": sched, "gemm_stages": stages,
               "packed_mode": planes}
        try:
            sdist.sharded_scan(chunks(3 * chunk), ops, K, ctx, N)   # warm-up
            ops.scan = engine.TopActivationScan(lo, hi, args.top, ctx, dev)
            phases = {} if sched == "sequential" else None
            sync()
            e0.record()
            res = sdist.sharded_scan(chunks(), ops, K, ctx, N, phase_times=phases)
            e1.record()
            sync()
            ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            out.update(ms=round(float(ms.item()), 2), tokens_per_s=round(args.tokens / (float(ms.item()) * 1e-3)),
                       exchange_used=ops.exchange, flagged_rows=int(ops.status.item()),
                       transport=(ops._push.transport if ops._push is not None else None))
            if phases:
                out["phase_ms_rank0"] = {k_: round(v_, 2) for k_, v_ in phases.items()}
                if world > 1:   # every rank's alone-times: a slow GPU sets the pace of the lockstep protocol
                    keys = sorted(phases)
                    mine = torch.tensor([phases[k_] for k_ in keys], dtype=torch.float32, device=dev)
                    allp = [torch.empty_like(mine) for _ in range(world)]
                    dist.all_gather(allp, mine)
                    out["phase_ms_per_rank"] = {k_: [round(float(a[i]), 1) for a in allp] for i, k_ in enumerate(keys)}
            if ref is None:
                ref = res
                if world > 1 and args.crosscheck_tokens > 0:
                    n_cc = min(args.tokens, args.crosscheck_tokens) // chunk * chunk or chunk
                    ops.scan = engine.TopActivationScan(lo, hi, args.top, ctx, dev)
                    res_fs = sdist.sharded_scan(chunks(n_cc), ops, K, ctx, N)
                    ops_tp = sdist.EngineOps(sae.encoder.weight.data, sae.encoder.bias.data, sae.b_dec.data, 0, N,
                                             args.top, ctx, dev)
                    w_lo, w_hi = sdist.token_slice(n_cc // ctx, world, rank)
                    x_tp = xs[w_lo * ctx:w_hi * ctx]
                    ctp = ops_tp.chunk_tokens(1)
                    res_tp = sdist.token_parallel_scan((x_tp[t0:t0 + ctp] for t0 in range(0, x_tp.shape[0], ctp)),
                                                       ops_tp, K, ctx, N, w_lo, n_top=args.top)
                    same = torch.tensor([1.0 if (torch.equal(res_tp.top_win, res_fs.top_win)
                                                 and torch.equal(res_tp.top_vals, res_fs.top_vals)) else 0.0], device=dev)
                    dist.all_reduce(same, op=dist.ReduceOp.MIN)
                    out["crosscheck_token_parallel"] = {"tokens": n_cc, "identical_lists": bool(same.item() == 1.0)}
                    del ops_tp
            else:
                same = torch.tensor([1.0 if (torch.equal(res.top_win, ref.top_win)
                                             and torch.equal(res.top_vals, ref.top_vals)) else 0.0], device=dev)
                if world > 1:
                    dist.all_reduce(same, op=dist.ReduceOp.MIN)
                out["same_lists_as_first_config"] = bool(same.item() == 1.0)
                if same.item() != 1.0:
                    out["windows_equal_frac"] = float((res.top_win == ref.top_win).float().mean().item())
                    out["max_rel_val_diff"] = float(((res.top_vals - ref.top_vals).abs()
                                                     / ref.top_vals.abs().clamp_min(1e-30)).max().item())
        except Exception as exc:   # one broken configuration must not waste the multi-GPU call
            out["error"] = repr(exc)[:400]
        if rank == 0:
            print("SWEEP " + json.dumps(out), flush=True)
        del ops
        torch.cuda.empty_cache()
    _capi.check(L.saeb_set_option(b"gemm_stages", 0), "gemm_stages")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
