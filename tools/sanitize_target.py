"""Small-shape pass over every kernel family, meant to run under compute-sanitizer (race / memory
checking).  Each step checks its result against a torch computation so that a sanitizer-clean but wrong run fails too.

    compute-sanitizer --tool memcheck  python tools/sanitize_target.py
    compute-sanitizer --tool racecheck python tools/sanitize_target.py
(tools/sanitize.sh runs both with a time limit and keeps the summaries under gpurun_out/.)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-sae_b200"))

import torch

from saeb200 import engine, synth
from saeb200.engine import TopActivationScan
from saeb200.overlap import OverlappedForward

DEV = torch.device("cuda:0")


def main():
    T, d, N, k = 300, 256, 2048, 16
    sae = synth.make_sae(d, N, k, DEV, seed=5)
    x = synth.make_activations(T, d, DEV, seed=6)
    pre = torch.relu((x.float() - sae.b_dec.data) @ sae.encoder.weight.data.T + sae.encoder.bias.data)
    ref_v, ref_i = pre.topk(k)
    for planes in (3, 4, 2):
        for vm in ((engine.VALUES_EXACT, engine.VALUES_BOUNDARY) if planes != 2 else (engine.VALUES_EXACT,)):
            v, i, _ = engine.encode_topk(x, sae.packed_encoder(planes), k, value_mode=vm)
            assert torch.equal(torch.sort(i, 1).values, torch.sort(ref_i, 1).values), (planes, vm)
            assert torch.allclose(v.sort(1).values, ref_v.sort(1).values, rtol=1e-3, atol=1e-5)
    v, i, _ = engine.encode_topk(x, sae.packed_encoder(3), k)
    sq = torch.zeros((), dtype=torch.float64, device=DEV)
    out = engine.decode(i, v, sae.W_dec.data, sae.b_dec.data, x=x, sq_err=sq)
    ref_out = (v.unsqueeze(-1) * sae.W_dec.data[i]).sum(1) + sae.b_dec.data
    assert torch.allclose(out, ref_out, rtol=1e-4, atol=1e-5)
    out2 = engine.decode(i, v, sae.W_dec.data, sae.b_dec.data, max_ctas=7)
    assert torch.equal(out, out2)
    ov = OverlappedForward(sae.packed_encoder(3), sae.W_dec.data, sae.b_dec.data, k, chunk=128, ctas_per_sm=1)
    a2, i2, o2 = torch.empty_like(v), torch.empty_like(i), torch.empty_like(out)
    ov.run(x, a2, i2, o2)
    torch.cuda.synchronize()
    assert torch.equal(a2, v) and torch.equal(i2, i) and torch.equal(o2, out)
    loc, act = engine.coo_extract(v.view(3, 100, k), i.view(3, 100, k), 100)
    assert loc.shape[0] == int((v > 1e-5).sum())
    arena = engine.CooArena(DEV, capacity=256)
    arena.append(v.view(3, 100, k), i.view(3, 100, k), 100)
    l2, a3 = arena.tensors()
    assert torch.equal(l2, loc) and torch.equal(a3, act)
    scan = TopActivationScan(0, N, 4, 20, DEV)
    scan.update(v, i, 0)
    sv, sw = scan.finalize()
    assert int((sw[:, 0] >= 0).sum()) > 0
    g = torch.rand(4, 50, 16, device=DEV)
    kth = engine.kth_of_gathered(g)
    assert torch.equal(kth, g.permute(1, 0, 2).reshape(50, 64).topk(16).values[:, -1])
    d_acts, dW = engine.decode_backward(i, v, sae.W_dec.data, torch.randn(T, d, device=DEV))
    assert d_acts.shape == v.shape and dW.shape == sae.W_dec.shape
    mean = engine.mean_activations(x, sae.packed_encoder(2), chunk_tokens=128)
    assert torch.allclose(mean, pre.mean(0), rtol=1e-3, atol=1e-5)
    feats, offs, sc, wins = engine.coo_top_windows(loc, act, 3, ctx_len=20, seq_len=100)
    assert feats.numel() > 0 and int(offs[-1]) == sc.numel() == wins.numel()
    print("sanitize target ok")


if __name__ == "__main__":
    main()
