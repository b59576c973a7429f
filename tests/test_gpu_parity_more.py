"""GPU parity tests, second file: decoder backward, attribution patching, packed mode 4, the full-size scan, 16-bit
decoder copies and the image scan.  Written at the end of round 1 without GPU access as non-strict xfail; all of them
passed on the driver's B200 at round end (GPUTEST_r01.json: 22 xpassed), so they are plain strict tests now."""
import os

import numpy as np
import pytest
import torch

import sae_oracle as O

from conftest import GOLDEN

pytestmark = [pytest.mark.gpu]
DEV = torch.device("cuda:0")


def test_decode_backward_matches_reference_autograd():
    """saeb_decode_backward_acts / _weight through the autograd seam vs the reference's own gradients."""
    from sae_auto_interp.sae import utils as seam

    g = np.load(os.path.join(GOLDEN, "decode_backward.npz"))
    idx = torch.from_numpy(g["top_idx"]).to(DEV)
    go = torch.from_numpy(g["grad_out"]).to(DEV)
    vals = torch.from_numpy(g["top_vals"]).to(DEV).requires_grad_(True)
    W = torch.from_numpy(g["W_dec"]).to(DEV).requires_grad_(True)
    out = seam.decoder_impl(idx, vals, W.mT)
    np.testing.assert_allclose(out.detach().cpu().numpy(), g["out"], rtol=1e-5, atol=1e-6)
    out.backward(go)
    np.testing.assert_allclose(vals.grad.cpu().numpy(), g["d_vals"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(W.grad.cpu().numpy(), g["d_W_dec"], rtol=1e-4, atol=1e-5)   # atomic summation order


@pytest.mark.parametrize("T,d,N,k", [(64, 4096, 2048, 64), (33, 100, 300, 7), (5, 30, 40, 40)])
def test_decode_backward_vs_oracle(T, d, N, k):
    """vector (d % 4 == 0) and scalar paths, zero activations skipped by the weight gradient only"""
    from saeb200 import engine

    gen = torch.Generator().manual_seed(T + d)
    W = torch.randn(N, d, generator=gen)
    idx = torch.stack([torch.randperm(N, generator=gen)[:k] for _ in range(T)])
    vals = torch.rand(T, k, generator=gen)
    vals[::3, 0] = 0.0
    go = torch.randn(T, d, generator=gen)
    ref_a, ref_w = O.decode_backward(idx, vals, W, go)
    d_acts, dW = engine.decode_backward(idx.to(DEV), vals.to(DEV), W.to(DEV), go.to(DEV))
    torch.testing.assert_close(d_acts.cpu(), ref_a, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(dW.cpu(), ref_w, rtol=1e-4, atol=1e-4)
    assert int(engine.decode_backward.last_err_flag.item()) == 0


def test_attribution_patching_matches_reference_golden():
    """features/patching on the device: fused encode with the latent clamped to 0, decode through the autograd seam
    (its backward kernels run because W_dec requires grad), attribution values vs the reference's."""
    from functools import partial

    from sae_auto_interp.features.patching import attribution_for_feature, get_logit_diff
    from sae_auto_interp.sae import Sae, SaeConfig
    from test_host_logic import _ToyLogitLM

    g = np.load(os.path.join(GOLDEN, "attribution.npz"))
    N, d = g["W_enc"].shape
    sae = Sae(d, SaeConfig(num_latents=N, k=int(g["k"])), device=DEV)
    with torch.no_grad():
        sae.encoder.weight.copy_(torch.from_numpy(g["W_enc"]))
        sae.encoder.bias.copy_(torch.from_numpy(g["b_enc"]))
        sae.W_dec.copy_(torch.from_numpy(g["W_dec"]))
        sae.b_dec.copy_(torch.from_numpy(g["b_dec"]))
    model = _ToyLogitLM(40, d, seed=42).to(DEV)
    inputs = {"input_ids": torch.from_numpy(g["input_ids"]).to(DEV)}
    metric = partial(get_logit_diff, answer_token_indices=torch.from_numpy(g["answers"]).to(DEV))
    sae_dict, m2n = {"layers.0": sae}, {model.layers[0]: "layers.0"}
    for f in g["features"].tolist():
        att = attribution_for_feature(model, inputs, sae_dict, m2n, metric, f)["layers.0"]
        np.testing.assert_allclose(att.float().numpy(), g[f"att_{f}"], rtol=5e-3, atol=2e-4)
    assert sae.W_dec.grad is not None and sae.W_dec.grad.shape == sae.W_dec.shape


@pytest.mark.parametrize("xdt", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("T,d,N,k", [(1024, 1024, 16384, 64), (300, 520, 2048 + 40, 32), (64, 4096, 8192, 256),
                                     (100, 50, 100, 10)])
def test_mode4_residual_refine_vs_oracle(T, d, N, k, xdt):
    """packed mode 4 (fp16 hi plane on the tensor cores + fp16 residual plane in the refinement): same parity bar as
    the default mode; fp32 activations silently take the exact route."""
    from test_gpu_parity import _assert_topk_parity, _sae_from_params

    p = O.init_params(d, N, k, seed=300 + T)
    x = torch.randn(T, d, generator=torch.Generator().manual_seed(T + 1)).to(xdt)
    sae = _sae_from_params(p, 4)
    sae.refine_values = "all"   # every value corrected: this test pins the grade of the corrected values
    enc = sae.encode(x.to(DEV))
    _assert_topk_parity(p, x.float(), enc.top_acts, enc.top_indices)
    # tighter than the 1e-3 bar: the corrected values are fp32-GEMM grade
    ref = O.encode(p, x.float())
    ri, rv = O.canonical_topk(ref.top_acts, ref.top_indices)
    gi, gv = O.canonical_topk(enc.top_acts.cpu(), enc.top_indices.cpu())
    same = (gi == ri).all(-1)
    np.testing.assert_allclose(gv[same], rv[same], rtol=2e-5, atol=1e-6)


def test_mode4_full_width_forward():
    """BASELINE config 2 shape on a row sample: index sets (tie-audited), values, reconstruction and FVU in mode 4,
    through the chunked two-stream forward as well (T >= 2 chunks)."""
    from test_gpu_parity import REL, _assert_topk_parity, _sae_from_params

    p = O.init_params(4096, 131072, 64, seed=1234)
    x = torch.randn(384, 4096, generator=torch.Generator().manual_seed(5)).to(torch.bfloat16)
    sae = _sae_from_params(p, 4)
    out = sae(x.to(DEV))
    assert _assert_topk_parity(p, x.float(), out.latent_acts, out.latent_indices) <= 4
    ref = O.forward(p, x.float())
    rel = (out.sae_out.cpu() - ref.sae_out).norm(dim=1) / ref.sae_out.norm(dim=1)
    assert float(rel.max()) < REL
    np.testing.assert_allclose(float(out.fvu), float(ref.fvu), rtol=REL)
    # mode 3 and mode 4 agree on the same device (identical index sets on every untied row)
    sae3 = _sae_from_params(p, 3)
    o3 = sae3(x.to(DEV))
    i3, _ = O.canonical_topk(o3.latent_acts.cpu(), o3.latent_indices.cpu())
    i4, _ = O.canonical_topk(out.latent_acts.cpu(), out.latent_indices.cpu())
    assert int((i3 != i4).any(-1).sum()) <= 4
    sae.overlap_chunk = 128   # force the chunked two-stream path on this small batch
    out2 = sae(x.to(DEV))
    assert torch.equal(out2.latent_indices, out.latent_indices)
    torch.testing.assert_close(out2.sae_out, out.sae_out, rtol=1e-5, atol=1e-6)


def test_full_size_scan_properties():
    """C3 / C4 shape on one GPU (131 072 features, 131 072 tokens, top-20 windows of 64 tokens): properties of the
    per-feature lists that need no oracle run.  (Verified kernels; the test itself was written without a GPU.)"""
    from saeb200 import synth
    from saeb200.engine import TopActivationScan

    N, d, k, ctx, n_top, T = 131072, 4096, 64, 64, 20, 131072
    sae = synth.make_sae(d, N, k, DEV, seed=1234)
    x = synth.make_activations(T, d, DEV, seed=9)
    enc = sae.encode(x)
    acts, idx = enc.top_acts, enc.top_indices

    def run(step):
        scan = TopActivationScan(0, N, n_top, ctx, DEV)
        for t0 in range(0, T, step):
            scan.update(acts[t0:t0 + step], idx[t0:t0 + step], t0 // ctx)
        s, w = scan.finalize()
        assert int(scan.overflow.item()) == 0
        return s.clone(), w.clone()

    s, w = run(ctx * 256)
    # (1) every list is sorted (score desc, window asc), window ids are valid and distinct, empty slots trail
    assert bool((s[:, :-1] >= s[:, 1:]).all())
    filled = w >= 0
    assert bool((w[filled] < T // ctx).all()) and bool((s[filled] > 1e-5).all()) and bool((s[~filled] == 0).all())
    assert bool((filled[:, :-1] | ~filled[:, 1:]).all())
    tie = (s[:, :-1] == s[:, 1:]) & filled[:, 1:]
    assert bool((w[:, :-1][tie] < w[:, 1:][tie]).all())
    srt = torch.sort(torch.where(filled, w, torch.arange(-n_top, 0, device=DEV).expand_as(w)), dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())
    # (2) the lists do not depend on how the token stream is chunked
    s2, w2 = run(ctx * 96)
    assert torch.equal(s, s2) and torch.equal(w, w2)
    # (3) every listed score is the max of that feature's TopK activations inside the listed window, and no window of
    #     the sample beats a full list's last entry without being in it (checked on a sample of features)
    feats = torch.randint(0, N, (64,), generator=torch.Generator().manual_seed(1)).tolist()
    a_w = acts.view(T // ctx, ctx * k)
    i_w = idx.view(T // ctx, ctx * k)
    for f in feats:
        pooled = torch.where(i_w == f, a_w, torch.zeros_like(a_w)).amax(dim=1)      # [n_windows]
        pooled = torch.where(pooled > 1e-5, pooled, torch.zeros_like(pooled))
        ref_v, ref_w = torch.sort(pooled, descending=True, stable=True)             # ties: smaller window first
        n = int((ref_v[:n_top] > 0).sum())
        assert torch.equal(s[f, :n], ref_v[:n]) and torch.equal(w[f, :n], ref_w[:n])
        assert bool((w[f, n:] == -1).all())
    # (4) feature-sharded lists concatenate to the unsharded ones
    lo, hi = 3 * N // 8, 4 * N // 8
    sc = TopActivationScan(lo, hi, n_top, ctx, DEV)
    for t0 in range(0, T, ctx * 256):
        sc.update(acts[t0:t0 + ctx * 256], idx[t0:t0 + ctx * 256], t0 // ctx)
    ss, sw = sc.finalize()
    assert torch.equal(ss, s[lo:hi]) and torch.equal(sw, w[lo:hi])


@pytest.mark.parametrize("wdt,tol", [(torch.float16, 5e-4), (torch.bfloat16, 4e-3)])
def test_decode_with_16bit_weight_copy(wdt, tol):
    """decode from an fp16 / bf16 copy of W_dec (half the gather bytes): row-wise relative error of the reconstruction
    vs the fp32 decode -- fp16 stays inside the 1e-3 bar, bf16 does not (hard part 1)."""
    from saeb200 import engine

    g = torch.Generator().manual_seed(11)
    N, d, k, T = 8192, 4096, 64, 256
    W = torch.randn(N, d, generator=g)
    W /= W.norm(dim=1, keepdim=True)
    b = torch.randn(d, generator=g) * 0.1
    vals = torch.rand(T, k, generator=g) * 3
    idx = torch.stack([torch.randperm(N, generator=g)[:k] for _ in range(T)])
    ref = engine.decode(idx.to(DEV), vals.to(DEV), W.to(DEV), b.to(DEV))
    out = engine.decode(idx.to(DEV), vals.to(DEV), W.to(DEV).to(wdt), b.to(DEV))
    rel = ((out - ref).norm(dim=1) / ref.norm(dim=1)).max().item()
    assert rel < tol
    torch.testing.assert_close(ref.cpu(), O.eager_decode(idx, vals, W.mT) + b, rtol=1e-4, atol=1e-4)


def test_image_scan_matches_the_per_feature_route():
    """TopImageScan (image_pool_kernel + scan_merge_kernel) vs the host route of the image constructor: per feature,
    mean over the base image tokens -> ranking (features/constructors.py:109-122); chunked updates, a feature shard"""
    from saeb200.engine import TopImageScan
    from sae_auto_interp.features.constructors import image_scores

    N, d, k, tpi, n_base, n_img, n_top = 1024, 64, 8, 40, 24, 70, 6
    p = O.init_params(d, N, k, seed=61)
    x = torch.randn(n_img * tpi, d, generator=torch.Generator().manual_seed(62)).to(torch.bfloat16)
    from test_gpu_parity import _sae_from_params

    enc = _sae_from_params(p).encode(x.to(DEV))
    acts, idx = enc.top_acts, enc.top_indices
    lo, hi = 128, 640
    scan = TopImageScan(lo, hi, n_top, tpi, n_base, DEV, bucket_cap=32)
    for i0 in range(0, n_img, 25):
        i1 = min(n_img, i0 + 25)
        scan.update(acts[i0 * tpi:i1 * tpi], idx[i0 * tpi:i1 * tpi], i0)
    s, w = (t.cpu() for t in scan.finalize())
    assert int(scan.overflow.item()) == 0
    a_c, i_c = acts.cpu(), idx.cpu()
    pos, img = torch.arange(n_img * tpi) % tpi, torch.arange(n_img * tpi) // tpi
    for f in range(lo, hi, 7):
        tok, col = ((i_c == f) & (a_c > 1e-5)).nonzero(as_tuple=True)
        sc = image_scores(torch.stack([img[tok], pos[tok]], 1), a_c[tok, col], n_img, n_base)
        order = sorted((i for i in range(n_img) if sc[i] > 1e-5), key=lambda i: (-float(sc[i]), i))[:n_top]
        assert w[f - lo, :len(order)].tolist() == order and bool((w[f - lo, len(order):] == -1).all())
        torch.testing.assert_close(s[f - lo, :len(order)], sc[order], rtol=1e-6, atol=1e-9)


def test_feature_sharded_image_scan_logical_shards():
    """The image form of the scan under feature sharding (EngineOps(image_n_base=...): TopImageScan per shard, every
    member of the token's global TopK exact, membership threshold from the gathered values): 4 logical shards on one
    device reproduce the unsharded TopImageScan over the SAE's own TopK output."""
    from saeb200 import dist as sdist, engine
    from saeb200.engine import TopImageScan
    from test_gpu_parity import _sae_from_params

    N, d, k, tpi, n_base, n_img, n_top, R = 2048, 256, 16, 24, 16, 40, 5, 4
    p = O.init_params(d, N, k, seed=91)
    x = torch.randn(n_img * tpi, d, generator=torch.Generator().manual_seed(92)).to(torch.bfloat16).to(DEV)
    sae = _sae_from_params(p)
    sae.refine_values = "all"
    enc = sae.encode(x)
    ref = TopImageScan(0, N, n_top, tpi, n_base, DEV)
    ref.update(enc.top_acts, enc.top_indices, 0)
    ref_s, ref_w = (t.cpu() for t in ref.finalize())
    shards = [sdist.shard_range(N, R, r) for r in range(R)]
    ops = [sdist.EngineOps(p.W_enc[lo:hi].to(DEV), p.b_enc[lo:hi].to(DEV), p.b_dec.to(DEV), lo, hi, n_top, tpi, DEV,
                           image_n_base=n_base) for lo, hi in shards]
    m1, step = 8, tpi * 10
    for c0 in range(0, x.shape[0], step):
        xc = x[c0:c0 + step]
        bounds = [o.local_bounds(xc, k) for o in ops]
        g = torch.cat([torch.stack([lb[:, :m1] for lb, _ in bounds], 0), torch.stack([ub[:, :m1] for _, ub in bounds], 0)], -1)
        ext_L, ext_U = engine.gathered_bounds(g.contiguous(), m1, k)
        outs = [o.local_topk(ext_L, ext_U) for o in ops]
        tok_thr = engine.kth_of_gathered(torch.stack([m for _, m, _ in outs], 0), k)
        for o, (v, m, i) in zip(ops, outs):
            o.scan_update(v, i, c0 // tpi, tok_thr, m)
    parts = [o.scan_finalize() for o in ops]
    got_s = torch.cat([a for a, _ in parts]).cpu()
    got_w = torch.cat([b for _, b in parts]).cpu()
    assert torch.equal(got_w, ref_w)
    torch.testing.assert_close(got_s, ref_s, rtol=1e-6, atol=1e-9)
    assert int((ref_w >= 0).sum()) > N        # the comparison is not vacuous


@pytest.mark.parametrize("planes", [3, 4])
def test_overlapped_forward_with_bounded_gather_grids_equals_sequential(planes):
    """The two-stream forward (GEMM launches of chunk c+1 beside persistent, bounded gather grids of chunk c) must give
    bit-identical TopK / reconstruction / residual to the one-stream path, and both must match the oracle."""
    from saeb200 import engine, synth
    from saeb200.overlap import OverlappedForward

    T, d, N, k = 2000, 1024, 16384, 32
    sae = synth.make_sae(d, N, k, DEV, seed=11)
    sae.encoder_planes = planes
    enc = sae.packed_encoder()
    x = synth.make_activations(T, d, DEV, seed=12)
    a0, i0, _ = engine.encode_topk(x, enc, k)
    sq0 = torch.zeros((), dtype=torch.float64, device=DEV)
    o0 = engine.decode(i0, a0, sae.W_dec.data, sae.b_dec.data, x=x, sq_err=sq0)
    for chunk, cps in ((512, 1), (768, 2), (4096, 1)):
        ov = OverlappedForward(enc, sae.W_dec.data, sae.b_dec.data, k, chunk=chunk, ctas_per_sm=cps)
        a1 = torch.empty_like(a0); i1 = torch.empty_like(i0); o1 = torch.empty_like(o0)
        sq1 = torch.zeros((), dtype=torch.float64, device=DEV)
        ov.run(x, a1, i1, o1, sq1)
        torch.cuda.synchronize()
        assert torch.equal(a1, a0) and torch.equal(i1, i0) and torch.equal(o1, o0)
        assert abs(float(sq1) - float(sq0)) <= 1e-9 * abs(float(sq0))
        assert int(ov.status.sum().item()) == 0
    p = O.SaeParams(sae.encoder.weight.data.cpu(), sae.encoder.bias.data.cpu(), sae.W_dec.data.cpu(),
                    sae.b_dec.data.cpu(), k)
    ref = O.forward(p, x[:256].float().cpu())
    gi, gv = O.canonical_topk(a0[:256].cpu(), i0[:256].cpu())
    ri, rv = O.canonical_topk(ref.latent_acts, ref.latent_indices)
    assert np.array_equal(gi, ri)
    np.testing.assert_allclose(gv, rv, rtol=1e-3, atol=1e-5)


def test_decode_persistent_grid_equals_one_cta_per_token():
    from saeb200 import engine

    gen = torch.Generator().manual_seed(5)
    T, d, N, k = 333, 512, 4096, 24
    W = torch.randn(N, d, generator=gen).to(DEV)
    b = torch.randn(d, generator=gen).to(DEV)
    idx = torch.stack([torch.randperm(N, generator=gen)[:k] for _ in range(T)]).to(DEV)
    vals = torch.rand(T, k, generator=gen).to(DEV)
    x = torch.randn(T, d, generator=gen).to(DEV).to(torch.bfloat16)
    sq0 = torch.zeros((), dtype=torch.float64, device=DEV)
    o0 = engine.decode(idx, vals, W, b, x=x, sq_err=sq0)
    for mc in (1, 7, 148, 1000):
        sq1 = torch.zeros((), dtype=torch.float64, device=DEV)
        o1 = engine.decode(idx, vals, W, b, x=x, sq_err=sq1, max_ctas=mc)
        assert torch.equal(o0, o1)
        assert abs(float(sq1) - float(sq0)) <= 1e-12 * abs(float(sq0))


def test_boundary_values_mode_full_width():
    """BASELINE config 2 shape, value mode "boundary" vs "all" on the same rows: identical index sets on every row,
    re-evaluated members bit-equal to the exact mode, the others within 1e-3 relative of it (measured ~5e-5), and
    far fewer rows gathered."""
    import ctypes

    from saeb200 import _capi, engine, synth

    L = _capi.lib()
    sae = synth.make_sae(4096, 131072, 64, DEV, seed=1234)
    x = synth.make_activations(4096, 4096, DEV, seed=21)
    enc = sae.packed_encoder()
    rows = {}
    res = {}
    for name, vm in (("all", engine.VALUES_EXACT), ("boundary", engine.VALUES_BOUNDARY)):
        _capi.check(L.saeb_set_option(b"stats", 1), "stats")
        res[name] = engine.encode_topk(x, enc, 64, value_mode=vm)[:2]
        torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * 8)()
        L.saeb_debug_stats(buf)
        rows[name] = int(buf[7]) / x.shape[0]
        _capi.check(L.saeb_set_option(b"stats", 0), "stats")
        assert int(engine.encode_topk.last_status.item()) == 0
    (va, ia), (vb, ib) = res["all"], res["boundary"]
    sa, sb = torch.sort(ia, 1), torch.sort(ib, 1)
    assert torch.equal(sa.values, sb.values), "value mode changed a TopK index set"
    ga, gb = torch.gather(va, 1, sa.indices), torch.gather(vb, 1, sb.indices)
    rel = ((ga - gb).abs() / ga.abs()).max().item()
    assert rel < 1e-3, rel
    assert float((ga == gb).float().mean()) > 0.02          # the boundary members were re-evaluated ...
    assert rows["boundary"] < 0.35 * rows["all"], rows       # ... and only they
    assert bool((vb[:, :-1] >= vb[:, 1:]).all())


@pytest.mark.parametrize("T,d,N,k", [(1024, 1024, 16384, 64), (700, 4096, 8192, 32), (2500, 512, 4096 + 512, 16)])
def test_cluster_of_two_pairs_multicast_equals_pair_kernel(T, d, N, k):
    """The 4-CTA-cluster form of the fused GEMM (two CTA pairs share the activation tile through TMA multicast) must
    give bit-identical TopK output to the CTA-pair form, and both must match the oracle.  `splits = 2` forces the
    two-split decomposition the cluster kernel needs on these small shapes."""
    from saeb200 import _capi, engine, synth

    L = _capi.lib()
    sae = synth.make_sae(d, N, k, DEV, seed=31)
    enc = sae.packed_encoder()
    x = synth.make_activations(T, d, DEV, seed=32)
    res = {}
    try:
        _capi.check(L.saeb_set_option(b"splits", 2), "splits")
        for mode in (0, 2):
            _capi.check(L.saeb_set_option(b"cluster4", mode), "cluster4")
            res[mode] = engine.encode_topk(x, enc, k, value_mode=engine.VALUES_EXACT)[:2]
            torch.cuda.synchronize()
            assert int(engine.encode_topk.last_status.item()) == 0
    finally:
        L.saeb_set_option(b"splits", 0)
        L.saeb_set_option(b"cluster4", 0)
    assert torch.equal(res[0][0], res[2][0]) and torch.equal(res[0][1], res[2][1])
    p = O.SaeParams(sae.encoder.weight.data.cpu(), sae.encoder.bias.data.cpu(), sae.W_dec.data.cpu(),
                    sae.b_dec.data.cpu(), k)
    n = min(T, 300)
    ref = O.encode(p, x[:n].float().cpu())
    gi, gv = O.canonical_topk(res[2][0][:n].cpu(), res[2][1][:n].cpu())
    ri, rv = O.canonical_topk(ref.top_acts, ref.top_indices)
    assert np.array_equal(gi, ri)
    np.testing.assert_allclose(gv, rv, rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize("tag", ["image", "text"])
def test_probe_activations_matches_reference_golden(tag):
    """mean-over-tokens top features + their activation maps (tools/probe_activations.py:109-126) through the
    engine (dense GEMM store reduced by saeb_column_sums, maps by saeb_feature_maps) vs the reference's tensors."""
    from sae_auto_interp.probe_activations import make_hook, probe_hidden
    from test_gpu_parity import _params, _sae_from_params

    g = np.load(os.path.join(GOLDEN, "probe.npz"))
    sae = _sae_from_params(_params(g))
    hidden = torch.from_numpy(g[f"{tag}_hidden"]).to(DEV)
    interval, drop = g[f"{tag}_interval"].tolist(), bool(g[f"{tag}_drop"])
    idx, acts = probe_hidden(sae, hidden, interval, drop_first=drop)
    assert idx.dtype == torch.int64 and acts.shape == tuple(g[f"{tag}_acts"].shape)
    # the feature set and its order by mean activation (descending); means closer than 1e-5 relative may swap
    ref_idx, ref_mean = g[f"{tag}_indices"], g[f"{tag}_mean"]
    if not np.array_equal(idx.numpy(), ref_idx):
        assert sorted(idx.tolist()) == sorted(ref_idx.tolist())
        for a, b in zip(idx.tolist(), ref_idx.tolist()):
            assert abs(ref_mean[a] - ref_mean[b]) <= 1e-5 * abs(ref_mean[b])
    order = [idx.tolist().index(f) for f in ref_idx.tolist()]
    np.testing.assert_allclose(acts.numpy()[order], g[f"{tag}_acts"], rtol=1e-3, atol=1e-5)
    # the same through the forward hook the tool registers
    sink = {}
    make_hook(sae, interval, sink, drop_first=drop)(None, None, (hidden, None))
    assert torch.equal(sink["topk_indices"], idx) and torch.equal(sink["topk_acts"], acts)


def test_mean_activations_streams_token_chunks():
    """column means are accumulated chunk by chunk: any chunking gives the same means (fp64 accumulation), equal to
    the dense tensor's own mean"""
    from saeb200 import engine, synth

    sae = synth.make_sae(256, 2048, 16, DEV, seed=71)
    x = synth.make_activations(1000, 256, DEV, seed=72)
    enc2 = sae.packed_encoder(2)
    m1 = engine.mean_activations(x, enc2, chunk_tokens=4096)
    m2 = engine.mean_activations(x, enc2, chunk_tokens=333)
    dense = sae.pre_acts(x)
    torch.testing.assert_close(m1, m2, rtol=1e-6, atol=1e-9)
    torch.testing.assert_close(m1, dense.double().mean(0).float(), rtol=1e-6, atol=1e-9)
    maps = engine.feature_maps(x, sae.encoder.weight.data, sae.encoder.bias.data, sae.b_dec.data,
                               torch.tensor([5, 77, 2047], device=DEV))
    torch.testing.assert_close(maps, dense[:, [5, 77, 2047]].T.contiguous(), rtol=1e-3, atol=1e-5)
    with pytest.raises(engine.SaebError):
        engine.feature_maps(x, sae.encoder.weight.data, sae.encoder.bias.data, sae.b_dec.data,
                            torch.tensor([2048], device=DEV))


class _ToyLM(torch.nn.Module):
    """the toy host model of the cache fixture (oracle/gen_golden.py::ToyLM): embedding -> one hooked linear layer"""

    def __init__(self, emb, layer_w):
        super().__init__()
        self.emb = torch.nn.Embedding.from_pretrained(torch.from_numpy(emb), freeze=True)
        self.layers = torch.nn.ModuleList([torch.nn.Linear(layer_w.shape[1], layer_w.shape[0], bias=False)])
        with torch.no_grad():
            self.layers[0].weight.copy_(torch.from_numpy(layer_w))

    @property
    def device(self):
        return self.emb.weight.device

    def forward(self, ids):
        return self.layers[0](self.emb(ids))


def _run_fixture_cache(tmp_path, n_splits=4):
    from sae_auto_interp.features import FeatureCache
    from test_gpu_parity import _params, _sae_from_params

    g = np.load(os.path.join(GOLDEN, "cache_chain.npz"))
    sae = _sae_from_params(_params(g))
    model = _ToyLM(g["emb"], g["layer_w"]).to(DEV)
    tokens = torch.from_numpy(g["tokens"])
    fc = FeatureCache(model, None, {"layers.0": sae}, batch_size=2, shard_size=100)
    fc.run(16, [{"input_ids": tokens[i]} for i in range(tokens.shape[0])])
    return fc, g, tokens


def test_device_cache_arena_writes_the_reference_split_files(tmp_path):
    """FeatureCache on the device end to end: triples accumulate in the HBM arena (no host copy per batch), save_splits
    buckets + sorts them on the device, and the files are byte-for-byte the ones the reference wrote for the same run
    (fixture: FeatureCache.run -> save_splits -> concate_safetensors of the unmodified reference)."""
    from safetensors.torch import load_file

    fc, g, _ = _run_fixture_cache(tmp_path)
    arena = fc.cache._arenas["layers.0"]
    assert arena.syncs <= 2, "the arena must not synchronise per batch"      # save() only
    assert fc.cache.feature_locations["layers.0"].device.type == "cpu"
    assert np.array_equal(fc.cache.feature_locations["layers.0"].numpy(), g["nofilter_locations"])
    assert fc.cache.device_tensors("layers.0")[0].is_cuda
    fc.save_splits(4, str(tmp_path), rank=0)
    fc.concate_safetensors(4, str(tmp_path))
    files = sorted(os.listdir(tmp_path / "layers.0"))
    assert files == sorted(g["split_files"].tolist())
    for f in files:
        data = load_file(str(tmp_path / "layers.0" / f))
        assert np.array_equal(data["locations"].numpy(), g[f"split_{f}_locations"])
        np.testing.assert_allclose(data["activations"].numpy(), g[f"split_{f}_activations"], rtol=1e-3)


def test_coo_arena_grows_without_losing_entries():
    from saeb200 import engine

    gen = torch.Generator().manual_seed(3)
    arena = engine.CooArena(DEV, capacity=64)
    locs, acts = [], []
    for b in range(5):
        v = torch.rand(3, 16, 8, generator=gen).to(DEV)
        v[v < 0.3] = 0.0
        i = torch.stack([torch.randperm(500, generator=gen)[:8] for _ in range(48)]).view(3, 16, 8).to(DEV)
        arena.append(v, i, 16, row_offset=10 * b)
        l, a = engine.coo_extract(v, i, 16, row_offset=10 * b)
        locs.append(l)
        acts.append(a)
    loc, act = arena.tensors()
    assert torch.equal(loc, torch.cat(locs)) and torch.equal(act, torch.cat(acts))
    assert arena.capacity >= loc.shape[0]


def test_loader_device_route_equals_host_route(tmp_path):
    """FeatureDataset.load with the window constructor bound through functools.partial (what the explain launchers
    do): the device route (one ranking pass per split file for all its features) must build the same examples as the
    per-feature host route, for every feature of the fixture's cache."""
    from functools import partial

    from sae_auto_interp.config import FeatureConfig
    from sae_auto_interp.features import FeatureDataset, pool_max_activation_windows

    fc, g, tokens = _run_fixture_cache(tmp_path)
    fc.fix_split_bounds = True
    fc.save_splits(4, str(tmp_path), rank=0)
    fc.concate_safetensors(4, str(tmp_path))
    cfg = FeatureConfig(width=64, example_ctx_len=4, min_examples=0, max_examples=5, n_splits=4)
    big = torch.zeros(100 + tokens.shape[0], tokens.shape[1], dtype=torch.long)
    big[100:] = tokens
    ds = FeatureDataset(str(tmp_path), cfg, modules=["layers.0"])
    ctor = partial(pool_max_activation_windows, tokens=big, cfg=cfg)
    host = ds.load(collate=True, constructor=ctor, device=None)
    ds2 = FeatureDataset(str(tmp_path), cfg, modules=["layers.0"])
    dev = ds2.load(collate=True, constructor=ctor, device=DEV)
    assert [r.feature.feature_index for r in host] == [r.feature.feature_index for r in dev] and len(host) > 20
    n_checked = 0
    for rh, rd in zip(host, dev):
        assert len(rh.examples) == len(rd.examples)
        sh = [float(e.activations.max()) for e in rh.examples]
        sd = [float(e.activations.max()) for e in rd.examples]
        assert sh == sd == sorted(sd, reverse=True)      # same pooled scores, descending
        # same windows, except among windows TIED at the cut-off score (repeated tokens give identical activations;
        # torch.topk's choice among equal values is unspecified, the device route takes the smaller window id)
        cut = sh[-1] if sh else 0.0
        key = lambda ex: (ex.tokens.tolist(), ex.activations.tolist())
        above_h = sorted(key(e) for e in rh.examples if float(e.activations.max()) > cut)
        above_d = sorted(key(e) for e in rd.examples if float(e.activations.max()) > cut)
        assert above_h == above_d
        n_checked += len(above_h)
    assert n_checked > 50


def test_image_loader_device_route_equals_host_route():
    """image constructor: ranking by the mean over the base image tokens on the device (sequential file-order sums,
    bit-equal to the host's) -> same images, same ImageExamples as the host route"""
    from sae_auto_interp.config import FeatureConfig
    from sae_auto_interp.features import pool_max_activations_windows_image
    from sae_auto_interp.features.constructors import image_scores
    from sae_auto_interp.features.features import Feature, FeatureRecord
    from sae_auto_interp.features.loader import BufferOutput
    from saeb200 import engine
    from synth_images import synth_image_cache

    ds, loc, act = synth_image_cache()
    feats, offs, scores, wins = engine.coo_top_windows(
        torch.cat([loc, torch.full((loc.shape[0], 1), 7)], 1) if loc.shape[1] == 2 else loc, act, 6 + 50, n_base=576,
        device=DEV)
    assert feats.tolist() == [7]
    ref = image_scores(loc, act, len(ds), 576)
    pos = ref > 0
    order = torch.argsort(-ref[pos], stable=True)
    want_ids = pos.nonzero().squeeze(1)[order][: 56]
    assert torch.equal(wins.cpu(), want_ids) and torch.equal(scores.cpu(), ref[want_ids])
    cfg = FeatureConfig(width=64, max_examples=6)
    bo = BufferOutput(Feature("layers.0", 7), loc, act)
    rec_h, rec_d = FeatureRecord(bo.feature), FeatureRecord(bo.feature)
    pool_max_activations_windows_image(rec_h, bo, ds, cfg, None)
    pool_max_activations_windows_image(rec_d, bo, ds, cfg, None, ranked=(scores.cpu(), wins.cpu()))
    assert len(rec_h.examples) == len(rec_d.examples) == 6
    for eh, ed in zip(rec_h.examples, rec_d.examples):
        assert torch.equal(eh.activations, ed.activations)
        assert np.array_equal(np.asarray(eh.image), np.asarray(ed.image))


def test_fused_encode_is_differentiable_like_the_reference_chain():
    """Sae.encode under autograd (SparseEncode): gradients w.r.t. the input and the encoder parameters equal those of
    the reference chain nn.Linear -> relu -> topk -> decode (sae/sae.py:172-191) evaluated with torch ops on the
    device; two SAEs chained (the multi-layer attribution-patching situation): the gradient reaches the first one's
    reconstruction through the second one's encoder."""
    from saeb200 import synth

    T, d, N, k = 96, 256, 2048, 16
    sae1, sae2 = synth.make_sae(d, N, k, DEV, seed=81), synth.make_sae(d, N, k, DEV, seed=82)
    for s in (sae1, sae2):
        s.refine_values = "all"
        s.requires_grad_(True)
        s.train_encoder = True
    x = synth.make_activations(T, d, DEV, seed=83, dtype=torch.float32).requires_grad_(True)
    wsum = torch.randn(T, d, generator=torch.Generator().manual_seed(84)).to(DEV)

    def fused(inp):
        r1 = sae1.decode(*sae1.encode(inp))
        r1.retain_grad()
        r2 = sae2.decode(*sae2.encode(r1))
        return r1, (r2 * wsum).sum()

    def dense(sae, inp):
        pre = torch.relu((inp - sae.b_dec) @ sae.encoder.weight.T + sae.encoder.bias)
        v, i = pre.topk(k, dim=-1)
        return (v.unsqueeze(-1) * sae.W_dec[i]).sum(1) + sae.b_dec

    r1, loss = fused(x)
    loss.backward()
    got = {"x": x.grad.clone(), "r1": r1.grad.clone(), "W1": sae1.encoder.weight.grad.clone(),
           "b1": sae1.encoder.bias.grad.clone(), "bd2": sae2.b_dec.grad.clone(), "Wd1": sae1.W_dec.grad.clone()}
    for t in (x, sae1.encoder.weight, sae1.encoder.bias, sae2.b_dec, sae1.W_dec, sae1.b_dec, sae2.encoder.weight,
              sae2.encoder.bias, sae2.W_dec):
        t.grad = None
    d1 = dense(sae1, x)
    d1.retain_grad()
    ((dense(sae2, d1)) * wsum).sum().backward()
    want = {"x": x.grad, "r1": d1.grad, "W1": sae1.encoder.weight.grad, "b1": sae1.encoder.bias.grad,
            "bd2": sae2.b_dec.grad, "Wd1": sae1.W_dec.grad}
    assert got["r1"] is not None and float(got["r1"].abs().max()) > 0     # gradient crossed the second SAE's encoder
    for name in got:
        err = (got[name] - want[name]).norm() / want[name].norm()
        assert float(err) < 1e-3, (name, float(err))
