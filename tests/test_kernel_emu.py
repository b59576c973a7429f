"""CPU: the device code of `multimodal-sae_b200/csrc/kernels_*.cuh` executed on the host by a small emulation of the CUDA
execution model (tests/emu/cuda_emu.h: CUDA threads as host threads, warp / block collectives as barriers).

There is no GPU in the build container; this tier checks the LOGIC of the simple SIMT kernels -- indexing, warp
reductions, shared-memory phases, scale conventions -- against the oracle before they ever run on a B200 (the tcgen05 /
TMA kernel cannot be emulated this way and is covered by the -m gpu parity tests only).  Nothing here is a product
path: the emulation library is built from the kernel headers into a temporary directory and only used by this file.
"""
import ctypes
import os
import shutil
import subprocess
from ctypes import c_float, c_int, c_longlong, c_void_p

import numpy as np
import pytest
import torch

import sae_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "multimodal-sae_b200", "csrc")
EMU = os.path.join(ROOT, "tests", "emu")
CUDA_INC = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
HEADERS = ["common.cuh"] + sorted(f for f in os.listdir(CSRC) if f.startswith("kernels_") and f.endswith(".cuh"))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_fp16.h")):
        pytest.skip("needs g++ and the CUDA toolkit headers")
    build = tmp_path_factory.mktemp("emu_build")
    for h in HEADERS:   # shared memory becomes ordinary static storage shared by the block's host threads
        src = open(os.path.join(CSRC, h)).read()
        open(build / h, "w").write(src.replace("extern __shared__", "extern").replace("__shared__", "static"))
    lib = build / "libemu.so"
    cmd = ["g++", "-std=c++20", "-O1", "-shared", "-fPIC", "-Wno-attributes", f"-I{build}", f"-I{EMU}", f"-I{CUDA_INC}",
           os.path.join(EMU, "emu_kernels.cpp"), "-o", str(lib), "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    handle = ctypes.CDLL(str(lib))
    handle.path = str(lib)
    return handle


def _p(a):
    """pointer to a numpy array the CALLER keeps alive (never pass a temporary: its memory is gone after this line)"""
    return c_void_p(a.ctypes.data)


def _bf16_raw(t: torch.Tensor) -> np.ndarray:
    return t.to(torch.bfloat16).view(torch.int16).numpy().copy()


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("vpl", [4, 16, 64, 0])
def test_kth_largest_kernels(emu, vpl):
    """every register tier of kth_gathered_reg_kernel and the memory-resident kernel vs torch.topk"""
    cases = {4: [(4, 13, 16, 16), (2, 7, 3, 6), (5, 9, 13, 65), (8, 5, 16, 1)],
             16: [(8, 11, 24, 64), (8, 9, 64, 64), (3, 10, 100, 17)],
             64: [(8, 6, 256, 100), (16, 5, 64, 64)],
             0: [(3, 5, 700, 64), (4, 9, 16, 16)]}[vpl]
    gen = torch.Generator().manual_seed(24 + vpl)
    for R, T, m, kth in cases:
        g = torch.randn(R, T, m, generator=gen)
        g[:, ::3, m // 2:] = 0.0
        ref = g.clamp_min(0).permute(1, 0, 2).reshape(T, R * m).topk(kth).values[:, -1].numpy()
        ga = np.ascontiguousarray(g.numpy())
        out = np.full(T, -1.0, np.float32)
        emu.emu_kth(c_int(vpl), _p(ga), c_int(R), c_longlong(T), c_int(m), c_int(kth), _p(out))
        assert np.array_equal(out, ref), (R, T, m, kth)
        if vpl == 0 or kth < 4:
            continue
        # member values of the sharded refinement: 3e38 for most entries, a few exact values (with duplicates), zeros;
        # rows with >= kth, kth - 1 ... kth - 9 and far fewer "certain" entries exercise every branch of the fast path
        g2 = torch.zeros(R, T, m)
        for t in range(T):
            flat = torch.zeros(R * m)
            n_sure = max(0, kth - (t % 11) + 1) if t % 5 else max(0, kth // 3)
            n_sure = min(n_sure, R * m - 3)
            perm = torch.randperm(R * m, generator=gen)
            flat[perm[:n_sure]] = 3.0e38
            n_oth = min(R * m - n_sure, 3 + t % 9)
            vals_ = torch.rand(n_oth, generator=gen) * 5 + 0.1
            if n_oth >= 2:
                vals_[1] = vals_[0]                       # a tie
            flat[perm[n_sure:n_sure + n_oth]] = vals_
            g2[:, t, :] = flat.view(R, m)
        ref2 = g2.permute(1, 0, 2).reshape(T, R * m).topk(kth).values[:, -1].numpy()
        out2 = np.full(T, -1.0, np.float32)
        emu.emu_kth(c_int(vpl), _p(np.ascontiguousarray(g2.numpy())), c_int(R), c_longlong(T), c_int(m), c_int(kth), _p(out2))
        assert np.array_equal(out2, ref2), (R, T, m, kth, "member values")


@pytest.mark.parametrize("T,d,N,k", [(5, 64, 40, 7), (3, 50, 30, 30), (2, 260, 64, 9)])
def test_decode_backward_kernels(emu, T, d, N, k):
    """decode_bwd_acts_kernel / decode_bwd_weight_kernel (vector and scalar paths, zero activations, an out-of-range
    index) vs the oracle's restatement of TritonDecoder.backward"""
    gen = torch.Generator().manual_seed(T * d)
    W = torch.randn(N, d, generator=gen)
    idx = torch.stack([torch.randperm(N, generator=gen)[:k] for _ in range(T)])
    vals = torch.rand(T, k, generator=gen)
    vals[::2, 0] = 0.0
    go = torch.randn(T, d, generator=gen)
    ref_a, ref_w = O.decode_backward(idx, vals, W, go)
    Wn, idn, vn, gn = (np.ascontiguousarray(a.numpy()) for a in (W, idx, vals, go))
    d_vals = np.full((T, k), np.nan, np.float32)
    dW = np.zeros((N, d), np.float32)
    err = np.zeros(1, np.int32)
    emu.emu_decode_bwd_acts(_p(gn), c_longlong(d), _p(idn), c_longlong(T), c_int(k), _p(Wn), c_longlong(d),
                            c_longlong(N), _p(d_vals), _p(err))
    emu.emu_decode_bwd_weight(_p(gn), c_longlong(d), _p(idn), _p(vn), c_longlong(T), c_int(k), c_longlong(d),
                              c_longlong(N), _p(dW), _p(err))
    np.testing.assert_allclose(d_vals, ref_a.numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(dW, ref_w.numpy(), rtol=1e-5, atol=1e-5)
    assert err[0] == 0
    bad = idn.copy()
    bad[0, 1] = N + 5
    emu.emu_decode_bwd_acts(_p(gn), c_longlong(d), _p(bad), c_longlong(T), c_int(k), _p(Wn), c_longlong(d),
                            c_longlong(N), _p(d_vals), _p(err))
    assert err[0] == 1 and d_vals[0, 1] == 0.0


# ---------------------------------------------------------------------------------------------
def _pipeline_inputs(emu, d, N, k, T, margin, seed, p=None, x=None):
    """pack (mode 4) + activation prep through the emulated kernels, and the candidate lists the fused GEMM epilogue +
    merge kernel would hand to the refinement (approximate values from the fp16 planes, top-K2 per row)."""
    if p is None:
        p = O.init_params(d, N, k, seed=seed)
        x = torch.randn(T, d, generator=torch.Generator().manual_seed(seed + 1)).to(torch.bfloat16)
    d_pad = (d + 7) // 8 * 8
    W = np.ascontiguousarray(p.W_enc.numpy())
    hi = np.zeros((N, d_pad), np.float16)
    lo = np.zeros((N, d_pad), np.float16)
    bias, wnorm, dnorm = (np.zeros(N, np.float32) for _ in range(3))
    trailer = np.zeros(64, np.float32)
    b_enc, b_dec = np.ascontiguousarray(p.b_enc.numpy()), np.ascontiguousarray(p.b_dec.numpy())
    emu.emu_pack(_p(W), _p(b_enc), _p(b_dec), c_longlong(N), c_longlong(d), c_longlong(d_pad),
                 _p(hi), _p(lo), _p(bias), _p(wnorm), _p(dnorm), _p(trailer))
    xraw = _bf16_raw(x)
    x16 = np.zeros((T, d_pad), np.float16)
    row_scale, xnorm, xdnorm = (np.zeros(T, np.float32) for _ in range(3))
    emu.emu_prep_x_bf16(_p(xraw), c_longlong(T), c_longlong(d), c_longlong(d), c_longlong(d_pad), _p(x16),
                        _p(row_scale), _p(xnorm), _p(xdnorm))
    # what the tensor cores + epilogue produce: (x16 . hi) * row_scale * w_unscale + folded bias
    acc = x16.astype(np.float64) @ hi.astype(np.float64).T
    a = (acc * row_scale[:, None].astype(np.float64) * float(trailer[0]) + bias[None, :]).astype(np.float32)
    K2 = k + margin
    order = np.lexsort((np.arange(N)[None, :].repeat(T, 0), -a), axis=1)[:, :K2]   # (value desc, index asc)
    cand_idx = order.astype(np.int64)
    cand_vals = np.take_along_axis(a, order, 1).astype(np.float32)
    assert (cand_vals > 0).all(), "the toy shape must give every row K2 positive candidates"
    return dict(p=p, x=x, xraw=xraw, W=W, hi=hi, lo=lo, bias=bias, wnorm=wnorm, dnorm=dnorm, trailer=trailer,
                row_scale=row_scale, xnorm=xnorm, xdnorm=xdnorm, cand_idx=np.ascontiguousarray(cand_idx),
                cand_vals=np.ascontiguousarray(cand_vals), K2=K2, d_pad=d_pad, a=a)


def test_pack_and_prep_kernels(emu):
    """w_stats / pack_w_f16 / pack_w_lo / prep_x_f16: folded bias, norms, power-of-two scales, hi + lo = W"""
    s = _pipeline_inputs(emu, d=100, N=96, k=6, T=5, margin=8, seed=3)
    p, W = s["p"], s["W"].astype(np.float64)
    unscale = float(s["trailer"][0])
    assert unscale > 0 and np.log2(unscale) == int(np.log2(unscale))
    np.testing.assert_allclose(s["bias"], (p.b_enc.double() - p.W_enc.double() @ p.b_dec.double()).numpy(), rtol=1e-6,
                               atol=1e-7)
    assert (s["wnorm"] >= np.linalg.norm(W, axis=1) * (1 - 1e-7)).all()
    np.testing.assert_allclose(s["wnorm"], np.linalg.norm(W, axis=1), rtol=1e-5)
    hi, lo = s["hi"].astype(np.float64)[:, :100], s["lo"].astype(np.float64)[:, :100]
    assert 2 ** 13 <= np.abs(s["hi"].astype(np.float32)).max() < 2 ** 14
    assert np.abs(hi * unscale - W).max() > 1e-6 * np.abs(W).max()                  # fp16 alone is lossy ...
    assert np.abs((hi + lo / 2048) * unscale - W).max() < 3e-7 * np.abs(W).max()     # ... hi + lo is not
    np.testing.assert_allclose(s["dnorm"], np.linalg.norm(hi * unscale - W, axis=1), rtol=1e-4)
    assert (s["hi"][:, 100:] == 0).all() and (s["lo"][:, 100:] == 0).all()           # zero padded rows
    x = s["x"].float().numpy()
    assert np.array_equal(s["hi"].shape, (96, 104))
    np.testing.assert_allclose(s["xnorm"], np.linalg.norm(x.astype(np.float64), axis=1), rtol=2e-5)
    assert (s["xdnorm"] == 0).all()                                                    # bf16 rows scale exactly
    prep = np.zeros((5, 104), np.float16)
    emu.emu_prep_x_bf16(_p(s["xraw"]), c_longlong(5), c_longlong(100), c_longlong(100), c_longlong(104), _p(prep),
                        _p(s["row_scale"]), _p(s["xnorm"]), _p(s["xdnorm"]))
    assert np.array_equal(prep.astype(np.float32)[:, :100] * s["row_scale"][:, None], x)


def _run_refine(emu, s, k, *, lo, ext_lower=None, threads=256, clamp=(-1, 0.0), value_mode=0, c_eps=2.0 ** -14,
                ext_upper=None, feat_thr=None, want_member=False):
    T, K2 = s["cand_vals"].shape
    d, N = s["W"].shape[1], s["W"].shape[0]
    out_vals = np.full((T, k), np.nan, np.float32)
    out_idx = np.full((T, k), -7, np.int64)
    status = np.zeros(64, np.int32)
    flag_rows = np.full(max(T, 64), -1, np.int32)
    out_member = np.full((T, k), np.nan, np.float32)
    emu.emu_refine_bf16(_p(s["xraw"]), c_longlong(T), c_longlong(d), _p(s["W"]), c_longlong(d), c_longlong(N),
                        _p(s["bias"]), _p(s["wnorm"]), _p(s["dnorm"]), _p(s["trailer"]), _p(s["xnorm"]), _p(s["xdnorm"]),
                        c_float(c_eps), _p(s["cand_vals"]), _p(s["cand_idx"]), c_int(K2), c_int(k),
                        c_longlong(clamp[0]), c_float(clamp[1]), _p(out_vals), _p(out_idx), _p(status), _p(flag_rows),
                        None if ext_lower is None else _p(ext_lower), _p(s["lo"]) if lo else None,
                        c_longlong(s["d_pad"]), c_int(threads), c_int(value_mode),
                        None if ext_upper is None else _p(ext_upper), None if feat_thr is None else _p(feat_thr),
                        _p(out_member) if want_member else None)
    if want_member:
        return out_vals, out_idx, int(status[0]), out_member
    return out_vals, out_idx, int(status[0])


@pytest.mark.parametrize("lo", [False, True])
@pytest.mark.parametrize("d,N,k,margin", [(64, 256, 6, 10), (100, 300, 8, 12)])
def test_refinement_kernels_end_to_end(emu, lo, d, N, k, margin):
    """refine_kernel (exact fp32 re-evaluation) and refine_lo_kernel (residual-plane correction) on candidate lists
    built from the fp16 planes: TopK index sets equal the oracle's fp32 encode, rows ordered (value desc, index asc),
    values fp32 grade, no row flagged; the feature-sharded variant (ext_lower, 128 threads) agrees."""
    T = 6
    s = _pipeline_inputs(emu, d=d, N=N, k=k, T=T, margin=margin, seed=11 + d)
    ref = O.encode(s["p"], s["x"].float())
    ri, rv = O.canonical_topk(ref.top_acts, ref.top_indices)
    # the single fp16 pass alone is not good enough to pass as the result
    assert not np.allclose(np.sort(s["cand_vals"][:, :k], 1), np.sort(rv, 1), rtol=1e-6)
    vals, idx, flagged = _run_refine(emu, s, k, lo=lo)
    assert flagged == 0
    gi, gv = O.canonical_topk(torch.from_numpy(vals), torch.from_numpy(idx))
    assert np.array_equal(gi, ri)
    np.testing.assert_allclose(gv, rv, rtol=3e-6, atol=1e-7)
    assert (vals[:, :-1] >= vals[:, 1:]).all()
    tie = vals[:, :-1] == vals[:, 1:]
    assert (idx[:, :-1][tie] < idx[:, 1:][tie]).all()
    # feature-sharded call: a lower bound of the "global" k-th value just below the true one changes nothing
    ext = (np.sort(rv, 1)[:, 0] * 0.999).astype(np.float32)
    v2, i2, _ = _run_refine(emu, s, k, lo=lo, ext_lower=ext, threads=128)
    assert np.array_equal(i2, idx) and np.array_equal(v2, vals)
    # ... and one just above the 3rd best value leaves only the latents whose upper bound still reaches it: a prefix of
    # the true list (at least the two best), then zeros
    ext_hi = (np.sort(rv, 1)[:, -3] * 1.0005).astype(np.float32)
    v3, i3, _ = _run_refine(emu, s, k, lo=lo, ext_lower=ext_hi, threads=128)
    n_kept = (v3 > 0).sum(1)
    assert (n_kept >= 2).all() and (n_kept < k).all()
    for r in range(T):
        assert np.array_equal(v3[r, :n_kept[r]], vals[r, :n_kept[r]]) and np.array_equal(i3[r, :n_kept[r]], idx[r, :n_kept[r]])
        assert (v3[r, n_kept[r]:] == 0).all()
    # steering clamp: the GEMM epilogue overrides the clamped column before the candidate selection; the refinement
    # takes that value verbatim (no re-evaluation) and the latent leads every row
    f = int(idx[0, -1])
    a = s["a"].copy()
    a[:, f] = 50.0
    order = np.lexsort((np.arange(N)[None, :].repeat(T, 0), -a), axis=1)[:, :s["K2"]]
    sc = dict(s, cand_idx=np.ascontiguousarray(order.astype(np.int64)),
              cand_vals=np.ascontiguousarray(np.take_along_axis(a, order, 1).astype(np.float32)))
    v4, i4, _ = _run_refine(emu, sc, k, lo=lo, clamp=(f, 50.0))
    assert (i4[:, 0] == f).all() and (v4[:, 0] == 50.0).all()
    for r in range(T):   # the other k - 1 entries are the row's best latents apart from f
        rest = [(v, i) for v, i in zip(vals[r], idx[r]) if i != f][:k - 1]
        assert [int(i) for i in i4[r, 1:]] == [int(i) for _, i in rest]


@pytest.mark.parametrize("lo", [False, True])
@pytest.mark.parametrize("d,N,k,margin", [(64, 256, 6, 10), (100, 300, 8, 12), (128, 512, 16, 24)])
def test_refinement_boundary_only_mode(emu, lo, d, N, k, margin):
    """value_mode 1: only candidates whose error interval straddles the k-th boundary are re-evaluated.  The index SET
    must still equal the oracle's on every row; members that were re-evaluated carry the exact value, the others the
    tensor-core value, which lies within the rigorous per-candidate bound (and far inside 1e-3 relative)."""
    T = 8
    s = _pipeline_inputs(emu, d=d, N=N, k=k, T=T, margin=margin, seed=31 + d)
    ref = O.encode(s["p"], s["x"].float())
    ri, rv = O.canonical_topk(ref.top_acts, ref.top_indices)
    v0, i0, _ = _run_refine(emu, s, k, lo=lo)
    # a wider accumulation slack than the real 2^-14 widens every error interval, so that some rows of these toy shapes
    # do have candidates straddling the k-th boundary (the bound stays valid: it only gets looser)
    v1, i1, flagged = _run_refine(emu, s, k, lo=lo, value_mode=1, c_eps=2.0 ** -9)
    assert flagged == 0
    g0, _ = O.canonical_topk(torch.from_numpy(v0), torch.from_numpy(i0))
    g1, gv1 = O.canonical_topk(torch.from_numpy(v1), torch.from_numpy(i1))
    assert np.array_equal(g1, ri) and np.array_equal(g0, ri)
    np.testing.assert_allclose(gv1, rv, rtol=1e-3, atol=1e-6)
    assert (v1[:, :-1] >= v1[:, 1:]).all()
    # some values must have stayed approximate (else the mode did nothing), and every value is either the exact one or
    # the candidate's approximate one
    exact_by_id = [{int(i): float(v) for i, v in zip(i0[r], v0[r])} for r in range(T)]
    approx_by_id = [{int(i): float(v) for i, v in zip(s["cand_idx"][r], s["cand_vals"][r])} for r in range(T)]
    n_exact = n_approx = 0
    for r in range(T):
        for i, v in zip(i1[r], v1[r]):
            if float(v) == exact_by_id[r][int(i)]:
                n_exact += 1
            else:
                assert float(v) == approx_by_id[r][int(i)]
                n_approx += 1
    assert n_approx > 0 and n_exact > 0, (n_approx, n_exact)
    # with an external lower bound (feature-sharded call) the mode is ignored: exact values everywhere
    ext = (np.sort(rv, 1)[:, 0] * 0.999).astype(np.float32)
    v2, i2, _ = _run_refine(emu, s, k, lo=lo, ext_lower=ext, threads=128, value_mode=1)
    assert np.array_equal(v2, v0) and np.array_equal(i2, i0)


@pytest.mark.parametrize("xdt,lo", [("f16", False), ("f16", True), ("f32", False)])
def test_refinement_with_fp16_and_fp32_activations(emu, xdt, lo):
    """fp16 activations (the steering path's hidden stream) survive the row scaling exactly, so both refinement
    variants apply; fp32 activations are rounded to 11 bits on their way into the tensor cores (a non-zero
    ||x - fp16(x)|| enters the bound) and only the exact route is used.  TopK sets vs the oracle in both cases."""
    d, N, k, T, margin = 64, 256, 6, 5, 12
    p = O.init_params(d, N, k, seed=91)
    xf = torch.randn(T, d, generator=torch.Generator().manual_seed(92))
    x = xf.to(torch.float16) if xdt == "f16" else xf
    d_pad = d
    W = np.ascontiguousarray(p.W_enc.numpy())
    hi, lo_pl = np.zeros((N, d_pad), np.float16), np.zeros((N, d_pad), np.float16)
    bias, wnorm, dnorm = (np.zeros(N, np.float32) for _ in range(3))
    trailer = np.zeros(64, np.float32)
    b_enc, b_dec = np.ascontiguousarray(p.b_enc.numpy()), np.ascontiguousarray(p.b_dec.numpy())
    emu.emu_pack(_p(W), _p(b_enc), _p(b_dec), c_longlong(N), c_longlong(d), c_longlong(d_pad), _p(hi), _p(lo_pl),
                 _p(bias), _p(wnorm), _p(dnorm), _p(trailer))
    xraw = np.ascontiguousarray(x.view(torch.int16).numpy() if xdt == "f16" else x.numpy())
    x16 = np.zeros((T, d_pad), np.float16)
    row_scale, xnorm, xdnorm = (np.zeros(T, np.float32) for _ in range(3))
    prep = emu.emu_prep_x_f16 if xdt == "f16" else emu.emu_prep_x_f32
    prep(_p(xraw), c_longlong(T), c_longlong(d), c_longlong(d), c_longlong(d_pad), _p(x16), _p(row_scale), _p(xnorm),
         _p(xdnorm))
    if xdt == "f16":
        assert (xdnorm == 0).all() and np.array_equal(x16.astype(np.float32) * row_scale[:, None], x.float().numpy())
    else:
        assert (xdnorm > 0).all()
        err = np.linalg.norm(x16.astype(np.float64) * row_scale[:, None] - x.double().numpy(), axis=1)
        assert (xdnorm >= err * (1 - 1e-6)).all() and (xdnorm < 1.01 * err + 1e-12).all()
    acc = x16.astype(np.float64) @ hi.astype(np.float64).T
    a = (acc * row_scale[:, None].astype(np.float64) * float(trailer[0]) + bias[None, :]).astype(np.float32)
    K2 = k + margin
    order = np.lexsort((np.arange(N)[None, :].repeat(T, 0), -a), axis=1)[:, :K2]
    cand_idx = np.ascontiguousarray(order.astype(np.int64))
    cand_vals = np.ascontiguousarray(np.take_along_axis(a, order, 1).astype(np.float32))
    out_vals = np.full((T, k), np.nan, np.float32)
    out_idx = np.full((T, k), -7, np.int64)
    status = np.zeros(64, np.int32)
    flag_rows = np.full(64, -1, np.int32)
    common = (c_longlong(T), c_longlong(d), _p(W), c_longlong(d), c_longlong(N), _p(bias), _p(wnorm), _p(dnorm),
              _p(trailer), _p(xnorm), _p(xdnorm), c_float(2.0 ** -14), _p(cand_vals), _p(cand_idx), c_int(K2), c_int(k),
              _p(out_vals), _p(out_idx), _p(status), _p(flag_rows))
    if xdt == "f16":
        emu.emu_refine_f16(_p(xraw), *common, _p(lo_pl) if lo else None, c_longlong(d_pad))
    else:
        emu.emu_refine_f32(_p(xraw), *common)
    assert status[0] == 0
    ref = O.encode(p, x.float())
    ri, rv = O.canonical_topk(ref.top_acts, ref.top_indices)
    gi, gv = O.canonical_topk(torch.from_numpy(out_vals), torch.from_numpy(out_idx))
    assert np.array_equal(gi, ri)
    np.testing.assert_allclose(gv, rv, rtol=3e-6, atol=1e-7)


def test_candidate_bounds_kernel(emu):
    """per-token k best lower bounds a_j - eps_j (descending, floored at 0) of candidate_bounds_kernel"""
    k = 6
    s = _pipeline_inputs(emu, d=64, N=256, k=k, T=5, margin=10, seed=5)
    T, K2 = s["cand_vals"].shape
    lb = np.full((T, k), np.nan, np.float32)
    emu.emu_candidate_bounds(_p(s["cand_vals"]), _p(s["cand_idx"]), c_longlong(T), c_int(K2), c_int(k), _p(s["wnorm"]),
                             _p(s["dnorm"]), _p(s["xnorm"]), _p(s["xdnorm"]), c_float(2.0 ** -14), c_longlong(-1), _p(lb), None)
    wn, dn = s["wnorm"][s["cand_idx"]], s["dnorm"][s["cand_idx"]]
    xn, xdn = s["xnorm"][:, None], s["xdnorm"][:, None]
    eps = np.float32(1.001) * (xn * dn + xdn * wn) + np.float32(2.0 ** -14) * xn * wn
    want = -np.sort(-np.maximum(s["cand_vals"] - eps, 0), axis=1)[:, :k]
    np.testing.assert_allclose(lb, want, rtol=1e-6, atol=1e-7)
    # every bound really is a lower bound of the exact activation of some latent, and the k-th one of the k-th value
    exact = np.sort(O.pre_acts(s["p"], s["x"].float()).numpy(), axis=1)[:, ::-1]
    assert (lb <= exact[:, :k] + 1e-7).all() and (lb[:, -1] > 0).all()


# ---------------------------------------------------------------------------------------------
# kernels that are verified on the GPU as well: kept under emulation as regression tests of their logic
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("slots,k,cnt,ties", [(8, 64, 250, False), (8, 112, 256, False), (8, 16, 230, True),
                                               (16, 200, 500, False), (32, 300, 1000, True)])
def test_compact_row_keeps_a_superset_of_the_top_k(emu, slots, k, cnt, ties):
    """compact_row (epilogue of the fused tcgen05 kernel): histogram threshold with the exact tie-capped fallback.
    Whatever path it takes, the compacted list keeps at least k entries including every entry above the k-th value,
    in arrival order, and the threshold it returns excludes nothing that could still belong to the top k."""
    rng = np.random.default_rng(slots + k)
    vals = rng.random(cnt).astype(np.float32) + 0.01
    if ties:
        vals[rng.random(cnt) < 0.9] = 0.5   # heavy ties force the exact path
    buf = np.zeros((32 * slots, 2), np.uint32)
    buf[:cnt, 0] = vals.view(np.uint32)
    buf[:cnt, 1] = np.arange(cnt) * 3 + 1
    before = buf[:cnt].copy()
    thr = np.zeros(1, np.float32)
    n_out = np.zeros(1, np.int32)
    emu.emu_compact_row(c_int(slots), _p(buf), c_int(cnt), c_int(k), _p(thr), _p(n_out))
    n = int(n_out[0])
    kept = buf[:n]
    assert k <= n <= cnt and n <= 32 * slots - 32
    kth = np.sort(vals)[-k]
    kept_vals = kept[:, 0].view(np.float32)
    assert (kept_vals >= thr[0]).all() and thr[0] <= kth
    gt = before[before[:, 0].view(np.float32) > kth]
    assert set(map(tuple, gt)) <= set(map(tuple, kept))                       # nothing above the k-th value is lost
    assert np.sort(kept_vals)[-k:].tolist() == np.sort(vals)[-k:].tolist()     # the top-k multiset survives
    pos = {tuple(e): i for i, e in enumerate(before)}
    order = [pos[tuple(e)] for e in kept]
    assert order == sorted(order)                                               # arrival order is preserved


@pytest.mark.parametrize("T,S,CAP,k,N", [(11, 2, 256, 64, 5000), (5, 3, 256, 7, 1200), (4, 1, 512, 200, 3000)])
def test_topk_merge_kernel(emu, T, S, CAP, k, N):
    """topk_merge_kernel: exact top-k over the S candidate lists of a row, canonical order (value desc, index asc),
    ties at the k-th value resolved towards the smaller index, zero padding on unused indices for short rows"""
    rng = np.random.default_rng(T * S)
    cand = np.zeros((T, S, CAP, 2), np.uint32)
    cnt = np.zeros((T, S), np.int32)
    want_v, want_i = np.zeros((T, k), np.float32), np.zeros((T, k), np.int64)
    for t in range(T):
        rows = []
        for s in range(S):
            n = int(rng.integers(k // S + 1, CAP)) if t != 1 else 2   # row 1 is short: fewer than k candidates
            cols = rng.choice(np.arange(s * (N // S), (s + 1) * (N // S)), size=n, replace=False)
            v = np.round(rng.random(n).astype(np.float32) * 8, 1 if t == 2 else 6) + np.float32(0.5)   # row 2: many ties
            cand[t, s, :n, 0] = v.view(np.uint32)
            cand[t, s, :n, 1] = cols
            cnt[t, s] = n
            rows += list(zip(v.tolist(), cols.tolist()))
        rows.sort(key=lambda e: (-e[0], e[1]))
        top = rows[:k]
        want_v[t, :len(top)] = [e[0] for e in top]
        want_i[t, :len(top)] = [e[1] for e in top]
        if len(top) < k:   # padded with value 0 on the smallest unused indices
            used, fill, j = {e[1] for e in top}, [], 0
            while len(fill) < k - len(top):
                if j not in used:
                    fill.append(j)
                j += 1
            want_i[t, len(top):] = fill
    out_v = np.full((T, k), np.nan, np.float32)
    out_i = np.full((T, k), -1, np.int64)
    emu.emu_topk_merge(_p(cand), _p(cnt), c_int(T), c_int(S), c_int(CAP), c_int(k), c_int(N), _p(out_v), _p(out_i))
    assert np.array_equal(out_v, want_v) and np.array_equal(out_i, want_i)


@pytest.mark.parametrize("T,S,CAP,K2,m1,N", [(9, 2, 256, 112, 24, 5000), (6, 2, 128, 20, 5, 900), (5, 1, 512, 128, 40, 4000)])
def test_select_bounds_kernel(emu, T, S, CAP, K2, m1, N):
    """scan_select_bounds_kernel (one register-resident kernel) = topk_merge_kernel + candidate_bounds_kernel: the same
    K2 selected candidates (ties at the K2-th value by smallest column; any order), and as bound lists the m1 largest
    lower / upper bounds of the row (any order, zero padded); gathered_bounds_kernel on such unsorted lists gives the
    k-th largest lower bound and max((k+1)-th largest upper bound, per shard the smallest bound sent)"""
    rng = np.random.default_rng(T * S + K2)
    cand = np.zeros((T, S, CAP, 2), np.uint32)
    cnt = np.zeros((T, S), np.int32)
    for t in range(T):
        for s_ in range(S):
            n = int(rng.integers(K2 // S + 1, CAP)) if t != 1 else 3      # row 1 is short: fewer than K2 candidates
            cols = rng.choice(np.arange(s_ * (N // S), (s_ + 1) * (N // S)), size=n, replace=False)
            v = np.round(rng.random(n).astype(np.float32) * 8, 1 if t == 2 else 6) + np.float32(0.5)   # row 2: many ties
            cand[t, s_, :n, 0] = v.view(np.uint32)
            cand[t, s_, :n, 1] = cols
            cnt[t, s_] = n
    wnorm = (rng.random(N).astype(np.float32) + 0.5)
    dnorm = (wnorm * rng.random(N).astype(np.float32) * 2e-3).astype(np.float32)
    xnorm = (rng.random(T).astype(np.float32) * 3 + 1)
    xdnorm = np.zeros(T, np.float32)
    c_eps = np.float32(2.0 ** -9)
    # reference route: merge (sorted) -> bounds (sorted, k = K2 columns)
    mv, mi = np.zeros((T, K2), np.float32), np.zeros((T, K2), np.int64)
    emu.emu_topk_merge(_p(cand), _p(cnt), c_int(T), c_int(S), c_int(CAP), c_int(K2), c_int(N), _p(mv), _p(mi))
    lb, ub = np.zeros((T, K2), np.float32), np.zeros((T, K2), np.float32)
    emu.emu_candidate_bounds(_p(mv), _p(mi), c_longlong(T), c_int(K2), c_int(K2), _p(wnorm), _p(dnorm), _p(xnorm),
                             _p(xdnorm), c_float(c_eps), c_longlong(-1), _p(lb), _p(ub))
    ov, oi = np.full((T, K2), np.nan, np.float32), np.full((T, K2), -1, np.int64)
    exch = np.full((T, 2 * m1), np.nan, np.float32)
    emu.emu_select_bounds(_p(cand), _p(cnt), c_int(T), c_int(S), c_int(CAP), c_int(K2), c_int(m1), _p(wnorm), _p(dnorm),
                          _p(xnorm), _p(xdnorm), c_float(c_eps), c_longlong(-1), _p(ov), _p(oi), _p(exch))
    for t in range(T):
        want = sorted((float(v), int(i)) for v, i in zip(mv[t], mi[t]) if v > 0)
        got = sorted((float(v), int(i)) for v, i in zip(ov[t], oi[t]) if v > 0)
        assert got == want, t
        assert (ov[t] >= 0).all() and ((oi[t] == 0) | (ov[t] > 0)).all()
        assert sorted(exch[t, :m1].tolist(), reverse=True) == lb[t, :m1].tolist(), t
        assert sorted(exch[t, m1:].tolist(), reverse=True) == ub[t, :m1].tolist(), t
    # gathered_bounds on R copies-with-noise of the unsorted payload vs numpy on the sorted lists
    R, k = 3, min(2 * m1, 3 * m1 - 1)
    g = np.stack([exch * np.float32(1.0 + 0.01 * r) for r in range(R)], 0).astype(np.float32)
    ext_L, ext_U = np.zeros(T, np.float32), np.zeros(T, np.float32)
    emu.emu_gathered_bounds(_p(np.ascontiguousarray(g)), c_int(R), c_longlong(T), c_int(m1), c_int(k), _p(ext_L), _p(ext_U))
    lo_all = np.sort(g[:, :, :m1].transpose(1, 0, 2).reshape(T, -1), 1)[:, ::-1]
    up_all = np.sort(g[:, :, m1:].transpose(1, 0, 2).reshape(T, -1), 1)[:, ::-1]
    assert np.array_equal(ext_L, lo_all[:, k - 1])
    assert np.array_equal(ext_U, np.maximum(up_all[:, k], g[:, :, m1:].min(-1).max(0)))


@pytest.mark.parametrize("tag", ["nofilter", "filter"])
def test_coo_extract_kernels(emu, tag):
    """coo_count / scan / emit: the reference's (row, pos, feature) triples in torch.nonzero order (features/cache.py:
    73-92), threshold and filter bitmap included; T spans two scan chunks"""
    from saeb200.engine import make_filter_bitmap

    N, k, seq, rows = 300, 6, 37, 31    # T = 1147 tokens > SCAN_CHUNK
    T = seq * rows
    gen = torch.Generator().manual_seed(8)
    vals = torch.rand(T, k, generator=gen)
    vals[vals < 0.15] = 0.0
    vals[5, 2] = 5e-6                    # below the 1e-5 threshold
    idx = torch.stack([torch.randperm(N, generator=gen)[:k] for _ in range(T)])
    sel = torch.tensor([1, 5, 15, 16, 31, 40, 63, 200, 299]) if tag == "filter" else None
    dense = torch.zeros(rows, seq, N).scatter_(-1, idx.view(rows, seq, k), vals.view(rows, seq, k))
    ref_loc, ref_act = O.get_nonzeros(dense, sel)
    ref_loc = ref_loc.clone()
    ref_loc[:, 0] += 100
    bitmap = None if sel is None else np.ascontiguousarray(make_filter_bitmap(sel, N).numpy())
    loc = np.full((T * k, 3), -1, np.int64)
    act = np.full(T * k, np.nan, np.float32)
    nnz = np.zeros(1, np.int64)
    vn, inn = np.ascontiguousarray(vals.numpy()), np.ascontiguousarray(idx.numpy())
    emu.emu_coo_extract(_p(vn), _p(inn), c_longlong(T),
                        c_int(k), c_float(1e-5), None if bitmap is None else _p(bitmap), c_longlong(seq),
                        c_longlong(100), _p(loc), _p(act), _p(nnz))
    n = int(nnz[0])
    assert n == ref_loc.shape[0]
    assert np.array_equal(loc[:n], ref_loc.numpy()) and np.array_equal(act[:n], ref_act.numpy())


def test_scan_pool_and_merge_kernels(emu):
    """scan_pool_kernel + scan_merge_kernel over several chunks and flushes vs the oracle's scan, for a feature shard
    with a per-token threshold (the feature-sharded form)"""
    N, k, ctx, n_top, n_win, cap = 120, 5, 8, 3, 20, 8
    T = ctx * n_win
    gen = torch.Generator().manual_seed(9)
    vals = torch.rand(T, k, generator=gen)
    vals[vals < 0.1] = 0.0
    idx = torch.stack([torch.randperm(N, generator=gen)[:k] for _ in range(T)])
    tok_thr = (vals.max(1).values * 0.5).contiguous()                 # drops the smaller half of every token's entries
    masked = torch.where(vals >= tok_thr[:, None], vals, torch.zeros_like(vals))
    lo, hi = 30, 100
    ref_s, ref_w = O.scan_top_windows(masked, idx, N, ctx, n_top)
    F = hi - lo
    top_vals = np.zeros((F, n_top), np.float32)
    top_win = np.full((F, n_top), -1, np.int64)
    feat_thr = np.full(F, 1e-5, np.float32)
    bucket = np.zeros((F, cap, 2), np.uint32)
    bucket_cnt = np.zeros(F, np.int32)
    overflow = np.zeros(1, np.int32)
    vn, inn, tn = (np.ascontiguousarray(a.numpy()) for a in (vals, idx, tok_thr))
    for w0 in range(0, n_win, cap):                                    # at most bucket_cap windows between two merges
        t0, t1 = w0 * ctx, min(T, (w0 + cap) * ctx)
        emu.emu_scan_pool(_p(vn[t0:t1]), _p(inn[t0:t1]), c_longlong(t1 - t0), c_int(k), c_int(ctx), c_float(1e-5),
                          c_longlong(lo), c_longlong(hi), c_longlong(w0), _p(tn[t0:t1]), _p(feat_thr), _p(bucket),
                          _p(bucket_cnt), c_int(cap), _p(overflow), None)
        emu.emu_scan_merge(_p(bucket), _p(bucket_cnt), c_int(cap), c_longlong(F), c_int(n_top), c_float(1e-5),
                           _p(top_vals), _p(top_win), _p(feat_thr))
    assert overflow[0] == 0
    assert np.array_equal(top_vals, ref_s[lo:hi]) and np.array_equal(top_win, ref_w[lo:hi])


def test_image_pool_kernel_and_merge(emu):
    """image_pool_kernel + scan_merge_kernel: per-feature best images by the mean activation over the first n_base
    positions of every image row (features/constructors.py:109-122), for a feature shard, over several calls and
    flushes with fewer persistent CTAs than images; the scratch hash must come back clean after every call"""
    from sae_auto_interp.features.constructors import image_scores

    N, k, tpi, n_base, n_img, n_top, cap, grid, slots = 150, 6, 14, 9, 23, 4, 8, 3, 256
    T = n_img * tpi
    gen = torch.Generator().manual_seed(31)
    vals = torch.rand(T, k, generator=gen)
    vals[vals < 0.2] = 0.0
    idx = torch.stack([torch.randperm(N, generator=gen)[:k] for _ in range(T)])
    lo, hi = 20, 130
    F = hi - lo
    # the reference's per-feature route: COO entries of a feature -> mean over the base positions -> ranking
    pos = torch.arange(T) % tpi
    img = torch.arange(T) // tpi
    want_s = np.zeros((F, n_top), np.float32)
    want_i = np.full((F, n_top), -1, np.int64)
    for f in range(lo, hi):
        hit = (idx == f) & (vals > 1e-5)
        tok, col = hit.nonzero(as_tuple=True)
        sc = image_scores(torch.stack([img[tok], pos[tok]], 1), vals[tok, col], n_img, n_base).numpy()
        order = sorted((i for i in range(n_img) if sc[i] > 1e-5), key=lambda i: (-sc[i], i))[:n_top]
        want_s[f - lo, :len(order)] = sc[order]
        want_i[f - lo, :len(order)] = order
    top_vals = np.zeros((F, n_top), np.float32)
    top_img = np.full((F, n_top), -1, np.int64)
    feat_thr = np.full(F, 1e-5, np.float32)
    bucket = np.zeros((F, cap, 2), np.uint32)
    bucket_cnt = np.zeros(F, np.int32)
    overflow = np.zeros(1, np.int32)
    hkeys = np.full((grid, slots), 0xFFFFFFFF, np.uint32)
    hsums = np.zeros((grid, slots), np.uint64)
    hlist = np.zeros((grid, slots), np.int32)
    vn, inn = np.ascontiguousarray(vals.numpy()), np.ascontiguousarray(idx.numpy())
    for i0 in range(0, n_img, cap):
        i1 = min(n_img, i0 + cap)
        emu.emu_image_pool(_p(vn[i0 * tpi:]), _p(inn[i0 * tpi:]), c_longlong(i1 - i0), c_longlong(tpi), c_int(k),
                           c_int(n_base), c_float(1e-5), c_longlong(lo), c_longlong(hi), c_longlong(i0), None,
                           _p(feat_thr), _p(hkeys), _p(hsums), _p(hlist), c_int(slots), _p(bucket), _p(bucket_cnt),
                           c_int(cap), _p(overflow), c_int(grid))
        assert (hkeys == 0xFFFFFFFF).all() and (hsums == 0).all()
        emu.emu_scan_merge(_p(bucket), _p(bucket_cnt), c_int(cap), c_longlong(F), c_int(n_top), c_float(1e-5),
                           _p(top_vals), _p(top_img), _p(feat_thr))
    assert overflow[0] == 0
    assert np.array_equal(top_img, want_i)
    np.testing.assert_allclose(top_vals, want_s, rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("w16,scalar,d", [(0, 0, 64), (1, 0, 64), (0, 1, 50)])
def test_decode_kernels(emu, w16, scalar, d):
    """decode_kernel (fp32 and fp16 weight rows) and decode_scalar_kernel: gather decode + bias + residual sum of
    squares, zero activations skipped, an out-of-range index flagged"""
    N, k, T = 90, 9, 7
    gen = torch.Generator().manual_seed(d + w16)
    W = torch.randn(N, d, generator=gen)
    b = torch.randn(d, generator=gen)
    vals = torch.rand(T, k, generator=gen)
    vals[::2, 1] = 0.0
    idx = torch.stack([torch.randperm(N, generator=gen)[:k] for _ in range(T)])
    x = torch.randn(T, d, generator=gen).to(torch.bfloat16)
    Wd = W.to(torch.float16) if w16 else W
    ref = O.sparse_decode(idx, vals, Wd.float()) + b
    Wn = np.ascontiguousarray(Wd.view(torch.int16).numpy() if w16 else Wd.numpy())
    out = np.full((T, d), np.nan, np.float32)
    sq = np.zeros(1, np.float64)
    err = np.zeros(1, np.int32)
    idn, vn = np.ascontiguousarray(idx.numpy()), np.ascontiguousarray(vals.numpy())
    bn, xraw = np.ascontiguousarray(b.numpy()), _bf16_raw(x)
    emu.emu_decode(_p(idn), _p(vn), c_longlong(T), c_int(k), _p(Wn), c_int(w16), c_longlong(d), c_longlong(N),
                   _p(bn), _p(out), _p(xraw), _p(sq), _p(err), c_int(scalar))
    np.testing.assert_allclose(out, ref.numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(sq[0], float(((ref - x.float()) ** 2).sum()), rtol=1e-5)
    assert err[0] == 0
    idn[0, 0] = N
    emu.emu_decode(_p(idn), _p(vn), c_longlong(T), c_int(k), _p(Wn), c_int(w16), c_longlong(d), c_longlong(N), None,
                   _p(out), None, None, _p(err), c_int(scalar))
    assert err[0] == 1


def test_dense_topk_kernel(emu):
    """dense_topk_kernel (Sae.select_topk on dense tensors, and the refine fallback): (value desc, index asc), ties at
    the k-th value towards the smaller index"""
    N, k, T = 500, 12, 2
    gen = torch.Generator().manual_seed(12)
    dense = torch.relu(torch.randn(T, N, generator=gen))
    dense[1] = torch.round(dense[1] * 4) / 4                          # a row full of ties
    want_v, want_i = np.zeros((T, k), np.float32), np.zeros((T, k), np.int64)
    for t in range(T):
        order = sorted(range(N), key=lambda j: (-float(dense[t, j]), j))[:k]
        want_i[t], want_v[t] = order, dense[t, order].numpy()
    out_v = np.full((T, k), np.nan, np.float32)
    out_i = np.full((T, k), -1, np.int64)
    dn = np.ascontiguousarray(dense.numpy())
    emu.emu_dense_topk(_p(dn), c_longlong(T), c_longlong(N), c_longlong(N), c_int(k),
                       _p(out_v), _p(out_i), c_int(256))
    assert np.array_equal(out_v, want_v) and np.array_equal(out_i, want_i)


@pytest.mark.parametrize("stage_x", [1, 0])
def test_refine_fallback_kernels(emu, stage_x):
    """exact_rows_kernel + dense_topk_kernel (first 64 flagged rows, redirected through the row map) and
    overflow_rows_kernel (the rest): flagged tokens come out as the exact fp32 TopK, the others are left untouched"""
    d, N, k, T = 32, 96, 8, 80
    p = O.init_params(d, N, k, seed=17)
    x = torch.randn(T, d, generator=torch.Generator().manual_seed(18)).to(torch.bfloat16)
    folded = (p.b_enc.double() - p.W_enc.double() @ p.b_dec.double()).float().numpy()
    flagged = np.array([t for t in range(T) if t % 7 != 3], np.int32)          # 69 rows: 64 wide + 5 overflow
    status = np.array([len(flagged)] + [0] * 63, np.int32)
    flag_rows = np.zeros(T, np.int32)
    flag_rows[:len(flagged)] = flagged
    out_v = np.full((T, k), -5.0, np.float32)
    out_i = np.full((T, k), -5, np.int64)
    dense = np.zeros((64, N), np.float32)
    xraw, Wn = _bf16_raw(x), np.ascontiguousarray(p.W_enc.numpy())
    emu.emu_refine_fallback(_p(xraw), c_longlong(d), _p(Wn), c_longlong(d),
                            c_longlong(N), _p(folded), _p(status), _p(flag_rows), c_longlong(-1), c_float(0.0),
                            _p(dense), c_int(k), _p(out_v), _p(out_i), c_int(2), c_int(128), c_int(stage_x))
    ref = O.encode(p, x.float())
    ri, rv = O.canonical_topk(ref.top_acts, ref.top_indices)
    gi, gv = O.canonical_topk(torch.from_numpy(out_v[flagged]), torch.from_numpy(out_i[flagged]))
    assert np.array_equal(gi, ri[flagged])
    np.testing.assert_allclose(gv, rv[flagged], rtol=2e-6, atol=1e-7)
    assert (out_v[flagged][:, :-1] >= out_v[flagged][:, 1:]).all()
    untouched = np.setdiff1d(np.arange(T), flagged)
    assert (out_v[untouched] == -5.0).all() and (out_i[untouched] == -5).all()


def test_total_variance_kernels(emu):
    """colstats_kernel + totvar_kernel: sum((x - x.mean(0))**2), the FVU denominator of sae/sae.py:204"""
    T, d = 150, 70
    x = (torch.randn(T, d, generator=torch.Generator().manual_seed(19)) * 3 + 1).to(torch.bfloat16)
    scratch = np.zeros(2 * d, np.float64)
    out = np.zeros(1, np.float64)
    xraw = _bf16_raw(x)
    emu.emu_total_variance_bf16(_p(xraw), c_longlong(T), c_longlong(d), c_longlong(d), _p(scratch), _p(out))
    xd = x.double()
    np.testing.assert_allclose(out[0], float(((xd - xd.mean(0)) ** 2).sum()), rtol=1e-10)


@pytest.mark.parametrize("lo", [False, True])
def test_forward_chain_on_the_emulator(emu, lo):
    """The SIMT half of the forward chained on the emulator, fed with what the tensor-core kernel produces (emulated
    in numpy from the packed fp16 planes): per-split candidate lists -> topk_merge_kernel -> refine[_lo]_kernel ->
    decode_kernel (+ residual sum of squares) -> colstats / totvar.  TopK sets, reconstruction and FVU vs the oracle's
    forward (reference Sae.forward, sae/sae.py:193-247)."""
    d, N, k, T, margin, S, CAP = 64, 512, 8, 6, 12, 2, 256
    s = _pipeline_inputs(emu, d=d, N=N, k=k, T=T, margin=margin, seed=71)
    K2 = s["K2"]
    a = s["a"]
    # candidate lists as the epilogue leaves them: per (row, split) a superset of the split's top-K2, in column order
    cand = np.zeros((T, S, CAP, 2), np.uint32)
    cnt = np.zeros((T, S), np.int32)
    for t in range(T):
        for sp in range(S):
            cols = np.arange(sp * N // S, (sp + 1) * N // S)
            v = a[t, cols]
            thr = max(np.sort(v)[-K2], 0.0)
            keep = cols[(v >= thr) & (v > 0)][:CAP]
            cand[t, sp, :len(keep), 0] = a[t, keep].view(np.uint32)
            cand[t, sp, :len(keep), 1] = keep
            cnt[t, sp] = len(keep)
    mvals = np.zeros((T, K2), np.float32)
    midx = np.zeros((T, K2), np.int64)
    emu.emu_topk_merge(_p(cand), _p(cnt), c_int(T), c_int(S), c_int(CAP), c_int(K2), c_int(N), _p(mvals), _p(midx))
    assert np.array_equal(mvals, s["cand_vals"]) and np.array_equal(midx, s["cand_idx"])
    sm = dict(s, cand_vals=mvals, cand_idx=midx)
    vals, idx, flagged = _run_refine(emu, sm, k, lo=lo)
    assert flagged == 0
    p, x = s["p"], s["x"]
    ref = O.forward(p, x.float())
    ri, rv = O.canonical_topk(ref.latent_acts, ref.latent_indices)
    gi, gv = O.canonical_topk(torch.from_numpy(vals), torch.from_numpy(idx))
    assert np.array_equal(gi, ri)
    np.testing.assert_allclose(gv, rv, rtol=3e-6, atol=1e-7)
    out = np.zeros((T, d), np.float32)
    sq = np.zeros(1, np.float64)
    err = np.zeros(1, np.int32)
    Wd, bd = np.ascontiguousarray(p.W_dec.numpy()), np.ascontiguousarray(p.b_dec.numpy())
    emu.emu_decode(_p(idx), _p(vals), c_longlong(T), c_int(k), _p(Wd), c_int(0), c_longlong(d), c_longlong(N), _p(bd),
                   _p(out), _p(s["xraw"]), _p(sq), _p(err), c_int(0))
    np.testing.assert_allclose(out, ref.sae_out.numpy(), rtol=1e-5, atol=1e-6)
    scratch = np.zeros(2 * d, np.float64)
    tv = np.zeros(1, np.float64)
    emu.emu_total_variance_bf16(_p(s["xraw"]), c_longlong(T), c_longlong(d), c_longlong(d), _p(scratch), _p(tv))
    np.testing.assert_allclose(sq[0] / tv[0], float(ref.fvu), rtol=1e-5)


def test_feature_sharded_scan_chain_on_the_emulator(emu):
    """The per-chunk data path of the feature-sharded scan with 2 logical shards, every SIMT kernel on the emulator:
    candidate_bounds -> (gather of the first m1 bound columns) -> kth-largest -> refinement restricted by the global
    lower bound -> (gather of the exact local values) -> kth-largest -> scan_pool with the per-token threshold ->
    scan_merge.  The concatenated per-feature lists must equal the unsharded oracle scan: the cache keeps a latent
    only if it is in the token's GLOBAL top-k (features/cache.py:210-218)."""
    d, N, k, ctx, n_top, margin, R, m1 = 64, 512, 8, 8, 3, 12, 2, 6
    T = ctx * 6
    p = O.init_params(d, N, k, seed=81)
    x = torch.randn(T, d, generator=torch.Generator().manual_seed(82)).to(torch.bfloat16)
    shards = []
    for r in range(R):
        lo, hi = r * N // R, (r + 1) * N // R
        sub = O.SaeParams(p.W_enc[lo:hi].contiguous(), p.b_enc[lo:hi].contiguous(), p.W_dec[lo:hi].contiguous(),
                          p.b_dec, k)
        sh = _pipeline_inputs(emu, d, hi - lo, k, T, margin, 0, p=sub, x=x)
        sh["lo"], sh["hi"] = lo, hi
        shards.append(sh)
    # exchange 1: the shards' k best lower bounds, first m1 columns only
    lbs = []
    for sh in shards:
        lb = np.zeros((T, k), np.float32)
        emu.emu_candidate_bounds(_p(sh["cand_vals"]), _p(sh["cand_idx"]), c_longlong(T), c_int(sh["K2"]), c_int(k),
                                 _p(sh["wnorm"]), _p(sh["dnorm"]), _p(sh["xnorm"]), _p(sh["xdnorm"]),
                                 c_float(2.0 ** -14), c_longlong(-1), _p(lb), None)
        lbs.append(lb[:, :m1])
    g1 = np.ascontiguousarray(np.stack(lbs, 0))
    ext_L = np.zeros(T, np.float32)
    emu.emu_kth(c_int(4), _p(g1), c_int(R), c_longlong(T), c_int(m1), c_int(k), _p(ext_L))
    ref = O.encode(p, x.float())
    kth_true = ref.top_acts.min(1).values.numpy()
    assert (ext_L <= kth_true).all() and (ext_L > 0.9 * kth_true).all()      # a valid, tight lower bound
    # restricted refinement per shard, then exchange 2: the exact local values
    outs = []
    for sh in shards:
        v, i, flagged = _run_refine(emu, sh, k, lo=False, ext_lower=ext_L, threads=128)
        assert flagged == 0
        outs.append((v, i + sh["lo"]))
    n_eval = sum(int((v > 0).sum()) for v, _ in outs)
    assert n_eval < 0.8 * R * k * T                                            # far fewer than k per shard and token
    g2 = np.ascontiguousarray(np.stack([v for v, _ in outs], 0))
    tok_thr = np.zeros(T, np.float32)
    emu.emu_kth(c_int(4), _p(g2), c_int(R), c_longlong(T), c_int(k), c_int(k), _p(tok_thr))
    np.testing.assert_allclose(tok_thr, kth_true, rtol=2e-6)
    # per-feature lists of each shard, filtered by the token's global k-th value
    parts = []
    for sh, (v, i) in zip(shards, outs):
        F = sh["hi"] - sh["lo"]
        top_vals = np.zeros((F, n_top), np.float32)
        top_win = np.full((F, n_top), -1, np.int64)
        feat_thr = np.full(F, 1e-5, np.float32)
        bucket = np.zeros((F, 8, 2), np.uint32)
        bucket_cnt = np.zeros(F, np.int32)
        overflow = np.zeros(1, np.int32)
        vc, ic = np.ascontiguousarray(v), np.ascontiguousarray(i)
        emu.emu_scan_pool(_p(vc), _p(ic), c_longlong(T), c_int(k), c_int(ctx), c_float(1e-5), c_longlong(sh["lo"]),
                          c_longlong(sh["hi"]), c_longlong(0), _p(tok_thr), _p(feat_thr), _p(bucket), _p(bucket_cnt),
                          c_int(8), _p(overflow), None)
        emu.emu_scan_merge(_p(bucket), _p(bucket_cnt), c_int(8), c_longlong(F), c_int(n_top), c_float(1e-5),
                           _p(top_vals), _p(top_win), _p(feat_thr))
        assert overflow[0] == 0
        parts.append((top_vals, top_win))
    ref_s, ref_w = O.scan_top_windows(ref.top_acts, ref.top_indices, N, ctx, n_top)
    np.testing.assert_allclose(np.concatenate([a for a, _ in parts]), ref_s, rtol=3e-6, atol=1e-7)
    assert np.array_equal(np.concatenate([b for _, b in parts]), ref_w)


def test_refinement_scan_mode_unsharded(emu):
    """value_mode 2 without sharding: only the members of the TopK whose feature can still take them (upper bound >=
    feat_thr[feature]) come out, with EXACT values; the others are not gathered at all."""
    d, N, k, margin, T = 100, 300, 8, 12, 8
    s = _pipeline_inputs(emu, d=d, N=N, k=k, T=T, margin=margin, seed=131)
    v0, i0, _ = _run_refine(emu, s, k, lo=False)                         # exact mode: the reference answer
    feat_thr = np.where(np.arange(N) % 2 == 0, np.float32(1e-5), np.float32(1e30)).astype(np.float32)
    v2, i2, flagged = _run_refine(emu, s, k, lo=False, value_mode=2, feat_thr=feat_thr, c_eps=2.0 ** -9)
    assert flagged == 0
    n_unwanted = 0
    for r in range(T):
        exact = {int(i): float(v) for i, v in zip(i0[r], v0[r])}
        got = {int(i): float(v) for i, v in zip(i2[r], v2[r]) if v > 0}
        assert all(got.get(i) == v for i, v in exact.items() if i % 2 == 0), r      # every wanted member, exact value
        assert all(i in exact and exact[i] == v for i, v in got.items()), r        # nothing but members, exact values
        n_unwanted += sum(1 for i in got if i % 2 == 1)   # only those whose membership had to be decided by value
    assert n_unwanted < 0.25 * sum(1 for r in range(T) for i in i0[r] if int(i) % 2 == 1)
    # without thresholds the mode is the exact mode (every member is re-evaluated)
    v3, i3, _ = _run_refine(emu, s, k, lo=False, value_mode=2)
    assert np.array_equal(v3, v0) and np.array_equal(i3, i0)


@pytest.mark.parametrize("threads", [128, -1])
def test_refinement_scan_mode_two_shards(emu, threads):
    """value_mode 2 under feature sharding (2 logical shards): bounds lists (lower AND upper) -> global L and U ->
    per-shard refinement that writes member values -> k-th largest member value = membership threshold.  The entries
    that pass it are exactly the oracle's global TopK, every value exact; with per-feature thresholds only the wanted
    members are gathered.  threads = -1: the warp-per-token kernel (refine_scan_warp_kernel) the launcher uses for
    these calls -- same entries with bit-identical values as the CTA-per-token kernel, in candidate order."""
    d, N, k, margin, R = 64, 512, 12, 12, 2
    T = 24
    p = O.init_params(d, N, k, seed=181)
    x = torch.randn(T, d, generator=torch.Generator().manual_seed(182)).to(torch.bfloat16)
    ref = O.encode(p, x.float())
    ri, rv = O.canonical_topk(ref.top_acts, ref.top_indices)
    shards = []
    for r in range(R):
        lo, hi = r * N // R, (r + 1) * N // R
        sub = O.SaeParams(p.W_enc[lo:hi].contiguous(), p.b_enc[lo:hi].contiguous(), p.W_dec[lo:hi].contiguous(),
                          p.b_dec, k)
        sh = _pipeline_inputs(emu, d, hi - lo, k, T, margin, 0, p=sub, x=x)
        sh["lo"], sh["hi"] = lo, hi
        shards.append(sh)
    c_eps = 2.0 ** -11      # wider intervals than the real ones: some candidates must be undecided on these toy shapes
    lbs, ubs = [], []
    for sh in shards:
        lb, ub = np.zeros((T, k), np.float32), np.zeros((T, k), np.float32)
        emu.emu_candidate_bounds(_p(sh["cand_vals"]), _p(sh["cand_idx"]), c_longlong(T), c_int(sh["K2"]), c_int(k),
                                 _p(sh["wnorm"]), _p(sh["dnorm"]), _p(sh["xnorm"]), _p(sh["xdnorm"]),
                                 c_float(c_eps), c_longlong(-1), _p(lb), _p(ub))
        assert (ub >= lb).all() and (np.diff(ub, axis=1) <= 0).all()
        lbs.append(lb)
        ubs.append(ub)
    all_lb, all_ub = np.concatenate(lbs, 1), np.concatenate(ubs, 1)
    ext_L = np.ascontiguousarray(-np.sort(-all_lb, 1)[:, k - 1])
    ext_U = np.ascontiguousarray(-np.sort(-all_ub, 1)[:, k])
    for feat_thr_on in (False, True):
        outs, n_gathered = [], 0
        for sh in shards:
            F = sh["hi"] - sh["lo"]
            thr = (np.where(np.arange(F) % 2 == 0, np.float32(1e-5), np.float32(1e30)).astype(np.float32)
                   if feat_thr_on else None)
            v, i, flagged, mem = _run_refine(emu, sh, k, lo=False, ext_lower=ext_L, ext_upper=ext_U, value_mode=2,
                                             threads=threads, c_eps=c_eps, feat_thr=thr, want_member=True)
            assert flagged == 0
            if threads < 0:   # the same (id, value, member value) entries as the CTA-per-token kernel, row by row
                v_c, i_c, _, m_c = _run_refine(emu, sh, k, lo=False, ext_lower=ext_L, ext_upper=ext_U, value_mode=2,
                                               threads=128, c_eps=c_eps, feat_thr=thr, want_member=True)
                for r in range(T):
                    a_ = sorted((int(ii), float(vv), float(mm)) for vv, ii, mm in zip(v[r], i[r], mem[r]) if mm > 0)
                    b_ = sorted((int(ii), float(vv), float(mm)) for vv, ii, mm in zip(v_c[r], i_c[r], m_c[r]) if mm > 0)
                    # (the warp kernel takes the external lower bound alone when there is one: it may hand out a few
                    # more boundary candidates than the CTA kernel, never fewer entries, and the same values)
                    assert set(b_) <= set(a_), r
                    assert all(m_ < 3e38 for _, _, m_ in set(a_) - set(b_)), r
                    assert ((v[r] == 0) & (i[r] == 0))[mem[r] == 0].all()
            assert ((mem == 0) | (mem >= 3e38) | (mem == v)).all()          # padding | certain | undecided (exact)
            outs.append((v, i + sh["lo"], mem))
            n_gathered += int((v > 0).sum())
        allm = np.concatenate([m for _, _, m in outs], 1)
        tok_thr = -np.sort(-allm, 1)[:, k - 1]
        n_sure = sum(int((m >= 3e38).sum()) for _, _, m in outs)
        n_maybe = sum(int(((m > 0) & (m < 3e38)).sum()) for _, _, m in outs)
        assert n_sure > 0 and n_maybe > 0
        for r in range(T):
            members = {}
            for v, i, m in outs:
                for vv, ii, mm in zip(v[r], i[r], m[r]):
                    if mm > 0 and mm >= tok_thr[r]:
                        members[int(ii)] = float(vv)
            assert sorted(members) == sorted(int(j) for j in ri[r]), r      # exactly the global TopK
            exact = {int(j): float(w) for j, w in zip(ri[r], rv[r])}
            for j, vv in members.items():
                if feat_thr_on and (j - (0 if j < N // R else N // R)) % 2 == 1:
                    # unwanted member: not gathered (0) unless its membership had to be decided by its exact value
                    assert vv == 0.0 or abs(vv - exact[j]) <= 3e-6 * exact[j]
                else:
                    assert abs(vv - exact[j]) <= 3e-6 * exact[j]            # wanted member: exact value
        if feat_thr_on:
            assert n_gathered < 0.8 * n_all
        n_all = n_gathered


# ---------------------------------------------------------------------------------------------
def _push_rank(lib_path, names, total_bytes, rank, R, chunks, rows, widths, blocks, q):
    """one emulated rank of the peer-memory exchange: a process that maps all R symmetric buffers (at its own
    addresses, like peer mappings) and runs push_gather_kernel for both exchanges of every chunk"""
    import random
    import time
    from ctypes import c_size_t, c_uint32
    from multiprocessing import shared_memory

    try:
        lib = ctypes.CDLL(lib_path)
        shms = [shared_memory.SharedMemory(name=n) for n in names]
        bases = (c_void_p * R)(*[ctypes.addressof(ctypes.c_char.from_buffer(s.buf)) for s in shms])
        mine = np.frombuffer(shms[rank].buf, dtype=np.uint8)
        flags_bytes, off, region = 4096, 4096, {}
        for c, w in enumerate(widths):
            region[c] = off
            off += (R * rows * w * 4 + 1023) // 1024 * 1024
        counters = np.zeros(len(widths), np.int32)
        rng = random.Random(rank)
        seq = [0] * len(widths)
        for chunk in range(chunks):
            T = rows if chunk % 4 else rows // 2          # a shorter chunk now and then: slab offsets follow T
            for c, w in enumerate(widths):
                seq[c] += 1
                src = np.full((T, w), 1000 * chunk + 10 * c + rank, np.float32) + np.arange(w, dtype=np.float32)[None, :] / 64
                lib.emu_push_gather(_p(src), c_size_t(src.nbytes), bases, c_int(R), c_int(rank), c_size_t(region[c]),
                                    c_size_t(0), c_int(c), c_uint32(seq[c]), c_void_p(counters[c:].ctypes.data),
                                    c_int(blocks))
                if rng.random() < 0.5:
                    time.sleep(rng.random() * 2e-3)       # the consumer kernel runs some time after the exchange
                got = mine[region[c]:region[c] + R * T * w * 4].view(np.float32).reshape(R, T, w)
                want = (np.array([1000 * chunk + 10 * c + r for r in range(R)], np.float32)[:, None, None]
                        + np.arange(w, dtype=np.float32)[None, None, :] / 64).repeat(T, 1)
                if not np.array_equal(got, want):
                    q.put((rank, f"chunk {chunk} channel {c}: stale or torn slab"))
                    return
                if counters[c] != 0:
                    q.put((rank, "the CTA counter was not restored"))
                    return
        q.put((rank, "ok"))   # the mappings go away with the process
    except Exception as exc:   # noqa: BLE001
        q.put((rank, repr(exc)))


@pytest.mark.filterwarnings("ignore:This process .* is multi-threaded:DeprecationWarning")
@pytest.mark.parametrize("R,blocks", [(2, 1), (4, 3)])
def test_push_gather_kernel_ranks_as_processes(emu, R, blocks):
    """push_gather_kernel itself (slab offsets, peer walk, last-CTA pattern, per-(channel, source) sequence flags with
    release / acquire, wait loop) with R emulated ranks as processes over shared memory, alternating the two channels
    like the scan does; every rank checks every gathered region after every exchange."""
    import multiprocessing as mp
    from multiprocessing import shared_memory

    rows, widths, chunks = 48, (8, 16), 12
    total = 4096 + sum((R * rows * w * 4 + 1023) // 1024 * 1024 for w in widths)
    shms = [shared_memory.SharedMemory(create=True, size=total) for _ in range(R)]
    try:
        for s in shms:
            np.frombuffer(s.buf, dtype=np.uint8)[:] = 0
        ctx = mp.get_context("fork")
        q = ctx.Queue()
        procs = [ctx.Process(target=_push_rank, args=(emu.path, [s.name for s in shms], total, r, R, chunks, rows,
                                                      widths, blocks, q)) for r in range(R)]
        for pr in procs:
            pr.start()
        results = [q.get(timeout=120) for _ in range(R)]
        for pr in procs:
            pr.join(30)
        assert sorted(results) == [(r, "ok") for r in range(R)], results
    finally:
        for s in shms:
            s.close()
            s.unlink()
