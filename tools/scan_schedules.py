"""Feature-sharded scan only, every schedule back to back in one process group (diagnostic; bench.py reports the default).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/scan_schedules.py \
        [--tokens 1048576] [--schedules lookahead,streams,sequential]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-sae_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=1048576)
    ap.add_argument("--schedules", default="lookahead,streams")
    ap.add_argument("--top", type=int, default=20)
    ap.add_argument("--reserve-sms", type=int, default=0, help="SMs the GEMM grid leaves free (set NCCL_MAX_CTAS too)")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    from saeb200 import dist as sdist, engine, synth

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    D, N, K, ctx = 4096, 131072, 64, 64
    sae = synth.make_sae(D, N, K, dev, seed=1234)
    lo, hi = sdist.shard_range(N, world, rank)
    ops = sdist.EngineOps(sae.encoder.weight.data[lo:hi], sae.encoder.bias.data[lo:hi], sae.b_dec.data, lo, hi,
                          args.top, ctx, dev)
    ops.reserve_sms = args.reserve_sms
    del sae
    chunk = ops.chunk_tokens(world)
    xs = synth.make_activations(args.tokens, D, dev, seed=99)

    def chunks(limit):
        for t0 in range(0, min(args.tokens, limit), chunk):
            yield xs[t0:t0 + chunk]

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = {"world": world, "tokens": args.tokens, "chunk_tokens": chunk}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ref = None
    for sch in args.schedules.split(","):
        kw = {}
        phases = None
        if sch == "sequential":
            phases = {}
            kw["phase_times"] = phases
            os.environ.pop("SAEB_SCAN_SCHEDULE", None)
        else:
            os.environ["SAEB_SCAN_SCHEDULE"] = sch
        ops.scan = engine.TopActivationScan(lo, hi, args.top, ctx, dev)
        sdist.sharded_scan(chunks(4 * chunk), ops, K, ctx, N, **kw)   # warm-up
        ops.scan = engine.TopActivationScan(lo, hi, args.top, ctx, dev)
        sync()
        e0.record()
        res = sdist.sharded_scan(chunks(args.tokens), ops, K, ctx, N, **kw)
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms.item())
        same = True
        if ref is None:
            ref = (res.top_vals.clone(), res.top_win.clone())
        else:   # every schedule must produce the same per-feature lists
            same = bool(torch.equal(ref[0], res.top_vals) and torch.equal(ref[1], res.top_win))
        out[sch] = {"ms": round(ms, 2), "tokens_per_s": round(args.tokens / (ms * 1e-3)), "same_lists": same}
        if phases:
            out[sch]["phase_ms_rank0"] = {k: round(v, 2) for k, v in phases.items()}
    if rank == 0:
        print("SCAN " + json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
