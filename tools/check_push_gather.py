#!/usr/bin/env python
"""Multi-GPU check + timing of the peer-memory all-gather (saeb_push_gather) against NCCL's all_gather_into_tensor.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_push_gather.py [--rows 37888] [--iters 50] [--no-multicast]

1. correctness: 40 exchanges on two channels (widths 24 and 64) and alternating slots must equal the NCCL result bit
   for bit on every rank (fresh random data each time, so a stale or torn slab is caught);
2. timing (CUDA events, max over ranks): NCCL vs push, alone on an idle GPU and with a tensor-core GEMM of this
   library's fused encoder in flight on another stream (the situation of the pipelined scan);
3. the same for the kth-largest kernel that consumes the gathered lists (register-resident vs the first version).
Rank 0 prints one JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-sae_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=37888)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--no-multicast", action="store_true")
    args = ap.parse_args()
    from saeb200 import _capi, engine, synth
    from saeb200.p2p import PushExchange

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    widths = (24, 64)
    px = PushExchange(None, dev, args.rows, widths, multicast=not args.no_multicast)
    gen = torch.Generator(device=dev).manual_seed(100 + rank)

    # ---- 1. correctness
    bad = 0
    for it in range(40):
        c = it & 1
        T = args.rows if it % 5 else args.rows // 2   # shorter last-chunk shapes too
        t = torch.rand(T, widths[c], device=dev, generator=gen)
        ref = torch.empty((world, T, widths[c]), device=dev)
        dist.all_gather_into_tensor(ref, t)
        got = px.gather(t, c, slot=(it >> 1) & 1)
        bad += int(not torch.equal(got, ref))
    bad_t = torch.tensor([bad], device=dev)
    dist.all_reduce(bad_t)

    def timed(fn, iters):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- 2. exchange timing, idle GPU
    res = {"world": world, "rows": args.rows, "transport": px.transport, "mismatching_exchanges": int(bad_t.item())}
    for c, w in enumerate(widths):
        t = torch.rand(args.rows, w, device=dev, generator=gen)
        out = torch.empty((world, args.rows, w), device=dev)
        res[f"nccl_ms_w{w}"] = timed(lambda: dist.all_gather_into_tensor(out, t), args.iters)
        res[f"push_ms_w{w}"] = timed(lambda: px.gather(t, c, 0), args.iters)

    # ---- with the fused encoder GEMM in flight on another stream
    d, N, k = 4096, 131072 // world, 64
    sae = synth.make_sae(d, N, k, dev, seed=7)
    sae.encoder_planes = 3
    enc = sae.packed_encoder()
    x = synth.make_activations(args.rows, d, dev, seed=8)
    L = _capi.lib()
    prep = torch.empty(L.saeb_prep_bytes(args.rows, d), dtype=torch.uint8, device=dev)
    ws = torch.empty(L.saeb_candidates_workspace_bytes(args.rows, d, N, k, 0), dtype=torch.uint8, device=dev)
    sg = torch.cuda.Stream(dev)
    sa = torch.cuda.Stream(dev, priority=-1)
    _capi.check(L.saeb_prep_activations(x.data_ptr(), engine._code(x), args.rows, d, d, prep.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream), "prep")
    torch.cuda.synchronize()

    def gemm():
        _capi.check(L.saeb_encode_candidates(prep.data_ptr(), args.rows, 0, args.rows, enc.blob.data_ptr(), d, N, k, 0,
                                             -1, 0.0, ws.data_ptr(), ws.numel(), sg.cuda_stream), "encode_candidates")

    def under_gemm(fn):
        def run():
            sg.wait_stream(torch.cuda.current_stream())
            sa.wait_stream(torch.cuda.current_stream())
            gemm()
            with torch.cuda.stream(sa):
                fn()
                fn()
            torch.cuda.current_stream().wait_stream(sg)
            torch.cuda.current_stream().wait_stream(sa)
        return run

    t64 = torch.rand(args.rows, 64, device=dev, generator=gen)
    out64 = torch.empty((world, args.rows, 64), device=dev)
    res["gemm_alone_ms"] = timed(under_gemm(lambda: None), 10)
    res["gemm_plus_2_nccl_ms"] = timed(under_gemm(lambda: dist.all_gather_into_tensor(out64, t64)), 10)
    res["gemm_plus_2_push_ms"] = timed(under_gemm(lambda: px.gather(t64, 1, 0)), 10)

    # ---- 3. kth-largest kernel on the gathered lists
    g = torch.rand(world, args.rows, 64, device=dev, generator=gen)
    for impl in (1, 0):
        _capi.check(L.saeb_set_option(b"kth_impl", impl), "set_option")
        res[f"kth_ms_impl{impl}"] = timed(lambda: engine.kth_of_gathered(g, 64), args.iters)
    _capi.check(L.saeb_set_option(b"kth_impl", 1), "set_option")
    if rank == 0:
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
