"""CUDA-event timeline of the overlapped forward (saeb200.overlap.OverlappedForward): when did the refinement and the
decode of chunk c run relative to the GEMM launches of chunk c+1?  Events are recorded on both streams around every
library call (no synchronisation inside the step); prints one JSON line with the intervals in ms since the step began.

    python tools/forward_timeline.py [--ctas-per-sm 2] [--stages 5]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-sae_b200"))

T, D, N, K = 65536, 4096, 131072, 64


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ctas-per-sm", type=int, default=2)
    ap.add_argument("--stages", type=int, default=5)
    ap.add_argument("--value-mode", type=int, default=1)
    args = ap.parse_args()
    import torch
    from saeb200 import _capi, engine, synth
    from saeb200.overlap import OverlappedForward

    L = _capi.lib()
    sae = synth.make_sae(D, N, K, "cuda", seed=1234)
    enc = sae.packed_encoder()
    x = synth.make_activations(T, D, "cuda", seed=3)
    acts = torch.empty((T, K), dtype=torch.float32, device="cuda")
    idx = torch.empty((T, K), dtype=torch.int64, device="cuda")
    out = torch.empty((T, D), dtype=torch.float32, device="cuda")
    sq = torch.zeros((), dtype=torch.float64, device="cuda")
    ov = OverlappedForward(enc, sae.W_dec.data, sae.b_dec.data, K, ctas_per_sm=args.ctas_per_sm, value_mode=args.value_mode)
    ov.gemm_stages = args.stages
    for _ in range(3):
        ov.run(x, acts, idx, out, sq)
    torch.cuda.synchronize()

    spans = []

    def wrap(obj, name, label):
        orig = getattr(obj, name)
        cnt = [0]

        def f(*a, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = orig(*a, **kw)
            e1.record()
            spans.append((label, cnt[0], e0, e1))
            cnt[0] += 1
            return r
        setattr(obj, name, f)

    wrap(L, "saeb_encode_candidates", "gemm")
    wrap(L, "saeb_refine_candidates", "merge+refine")
    import saeb200.overlap as ovm
    wrap(ovm.engine, "decode", "decode")
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sq.zero_()
    t0.record()
    ov.run(x, acts, idx, out, sq)
    t1.record()
    torch.cuda.synchronize()
    tl = sorted(((lab, c, round(t0.elapsed_time(a), 3), round(t0.elapsed_time(b), 3)) for lab, c, a, b in spans),
                key=lambda s: s[2])
    gem = {c: (a, b) for lab, c, a, b in tl if lab == "gemm"}
    inside = []
    for lab, c, a, b in tl:
        if lab == "gemm" or c + 1 not in gem:
            continue
        g0, g1 = gem[c + 1]
        ov_ms = max(0.0, min(b, g1) - max(a, g0))
        inside.append(ov_ms / max(b - a, 1e-9))
    print("TIMELINE " + json.dumps({"ctas_per_sm": args.ctas_per_sm, "gemm_stages": args.stages,
                                     "value_mode": args.value_mode, "step_ms": round(t0.elapsed_time(t1), 3),
                                     "gather_time_inside_next_gemm_frac": round(sum(inside) / max(len(inside), 1), 3),
                                     "spans_ms": tl}))


if __name__ == "__main__":
    main()
