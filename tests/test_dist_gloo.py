"""CPU, world_size 2 over gloo: the multi-GPU choreography of saeb200.dist (feature-sharded scan with the per-token
threshold exchange and the single final all-gather; token-parallel slicing).  The per-rank compute steps are injected
oracle-backed ops -- the CUDA ops are exercised by the -m gpu tests; here the collectives, offsets and merge are."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import sae_oracle as O


class OracleOps:
    """Same interface as saeb200.dist.EngineOps, computed with the oracle on CPU."""

    def __init__(self, p, lo, hi, n_top, ctx_len):
        self.p, self.feat_lo, self.feat_hi, self.n_top, self.ctx_len = p, lo, hi, n_top, ctx_len
        self.sub = O.SaeParams(p.W_enc[lo:hi], p.b_enc[lo:hi], p.W_dec[lo:hi], p.b_dec, p.k)
        self.acts, self.idx, self.thr = [], [], []
        self._v, self._i = {}, {}
        self.first_window = None

    def local_bounds(self, x, k, slot=0):
        pa = O.pre_acts(self.sub, x.float())
        self._v[slot], self._i[slot] = pa.topk(k, sorted=True)
        v = self._v[slot]
        return v * (1 - 1e-3), v * (1 + 1e-3)   # lower / upper bounds, like the engine's a_j -/+ eps_j

    def local_topk(self, ext_L=None, ext_U=None, slot=0):
        """same contract as EngineOps.local_topk: sharded calls also return the member values (3e38 = certainly in the
        token's global TopK, exact value = undecided, 0 = cannot be in it)"""
        v = self._v[slot]
        if ext_L is None:
            return v, None, self._i[slot] + self.feat_lo
        lb, ub = v * (1 - 1e-3), v * (1 + 1e-3)
        sure = (lb > ext_U[:, None]) & (v > 0)
        maybe = ~sure & (ub >= ext_L[:, None]) & (v > 0)
        member = torch.where(sure, torch.full_like(v, 3.0e38), torch.where(maybe, v, torch.zeros_like(v)))
        vals = torch.where(sure | maybe, v, torch.zeros_like(v))
        return vals, member, self._i[slot] + self.feat_lo

    def kth_of_gathered(self, gathered, kth=None):
        R, T, m = gathered.shape
        kth = m if kth is None else kth
        return gathered.clamp_min(0).permute(1, 0, 2).reshape(T, R * m).topk(kth).values[:, -1].contiguous()

    def scan_update(self, vals, idx, window_base, tok_thr, member=None):
        if tok_thr is not None:
            decide = vals if member is None else member
            vals = torch.where(decide >= tok_thr[:, None], vals, torch.zeros_like(vals))
        if self.first_window is None:
            self.first_window = window_base   # chunks arrive in order: later ones continue this numbering
        self.acts.append(vals)
        self.idx.append(idx)

    def scan_finalize(self):
        acts, idx = torch.cat(self.acts), torch.cat(self.idx)
        s, w = O.scan_top_windows(acts, idx - self.feat_lo, self.feat_hi - self.feat_lo, self.ctx_len, self.n_top)
        w = np.where(w >= 0, w + (self.first_window or 0), w)
        return torch.from_numpy(s), torch.from_numpy(w)


class PackedOracleOps(OracleOps):
    """the ops of the pipelined CUDA path: GEMM and bounds are separate steps and the bounds arrive as ONE unsorted
    payload [T, 2*m1] (saeb_candidate_bounds_packed); saeb200.dist must not rely on any order inside a shard's lists"""

    packed_bounds = True

    def local_gemm(self, x, k, slot=0, prepped=False):
        self._lbub = super().local_bounds(x, k, slot)

    def local_bounds_finish(self, slot=0, coresident=False, pack_m1=None):
        lb, ub = self._lbub
        if pack_m1 is None:
            return lb, ub
        g = torch.Generator().manual_seed(1000 + slot + lb.shape[0])
        perm_l, perm_u = torch.randperm(pack_m1, generator=g), torch.randperm(pack_m1, generator=g)
        return torch.cat([lb[:, :pack_m1][:, perm_l], ub[:, :pack_m1][:, perm_u]], -1).contiguous()


def _worker(rank, world, port, exact, pipelined, out_dir, lb_width=None, packed=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "multimodal-sae_b200"))
    sys.path.insert(0, os.path.join(root, "oracle"))
    from saeb200 import dist as sdist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N, d, k, ctx, n_top = 96, 32, 6, 8, 3
        p = O.init_params(d, N, k, seed=77)
        x = torch.randn(ctx * 12, d, generator=torch.Generator().manual_seed(78)).to(torch.bfloat16)
        lo, hi = sdist.shard_range(N, world, rank)
        ops = (PackedOracleOps if packed else OracleOps)(p, lo, hi, n_top, ctx)
        step = ctx * (2 if pipelined else 4)   # 6 / 3 chunks
        chunks = [x[i:i + step] for i in range(0, x.shape[0], step)]
        res = sdist.sharded_scan(chunks, ops, k, ctx, N, exact=exact, pipelined=pipelined, lb_width=lb_width)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), vals=res.top_vals.numpy(), win=res.top_win.numpy())
    finally:
        dist.destroy_process_group()


def _worker_token_parallel(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "multimodal-sae_b200"))
    sys.path.insert(0, os.path.join(root, "oracle"))
    from saeb200 import dist as sdist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N, d, k, ctx, n_top = 96, 32, 6, 8, 3
        p = O.init_params(d, N, k, seed=77)
        x = torch.randn(ctx * 12, d, generator=torch.Generator().manual_seed(78)).to(torch.bfloat16)
        x[ctx * 7:ctx * 8] = x[ctx * 2:ctx * 3]   # a window repeated on the other rank: equal scores, ids decide
        ops = OracleOps(p, 0, N, n_top, ctx)
        lo, hi = sdist.token_slice(x.shape[0] // ctx, world, rank)   # whole windows per rank
        mine = x[lo * ctx:hi * ctx]
        chunks = [mine[i:i + 2 * ctx] for i in range(0, mine.shape[0], 2 * ctx)]
        res = sdist.token_parallel_scan(chunks, ops, k, ctx, N, lo, n_top=n_top)
        np.savez(os.path.join(out_dir, f"tp_rank{rank}.npz"), vals=res.top_vals.numpy(), win=res.top_win.numpy())
    finally:
        dist.destroy_process_group()


def test_token_parallel_scan_two_ranks(tmp_path):
    """token-parallel form of the scan: no per-chunk exchange, one all-gather + per-feature merge at the end; must
    equal the single-process scan, including the window order of equal scores that come from different ranks"""
    world = 2
    mp.spawn(_worker_token_parallel, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (np.load(tmp_path / f"tp_rank{r}.npz") for r in range(world))
    assert np.array_equal(r0["vals"], r1["vals"]) and np.array_equal(r0["win"], r1["win"])
    N, d, k, ctx, n_top = 96, 32, 6, 8, 3
    p = O.init_params(d, N, k, seed=77)
    x = torch.randn(ctx * 12, d, generator=torch.Generator().manual_seed(78)).to(torch.bfloat16)
    x[ctx * 7:ctx * 8] = x[ctx * 2:ctx * 3]
    enc = O.encode(p, x.float())
    ref_s, ref_w = O.scan_top_windows(enc.top_acts, enc.top_indices, N, ctx, n_top)
    np.testing.assert_array_equal(r0["vals"], ref_s)
    np.testing.assert_array_equal(r0["win"], ref_w)
    both = (ref_w == 2).any(1) & (ref_w == 7).any(1)   # features whose list holds both copies of the repeated window
    assert both.any()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("exact,pipelined,lb_width,packed", [
    (True, False, None, False), (False, False, None, False), (True, True, None, False), (False, True, None, False),
    (True, False, 3, False), (True, True, 4, False), (True, False, None, True), (True, False, 4, True)])
def test_feature_sharded_scan_two_ranks(tmp_path, exact, pipelined, lb_width, packed):
    """sequential schedule and the one-chunk-lookahead schedule with asynchronous all-gathers; `lb_width` = columns of
    the lower-bound lists in exchange 1 (3 = k / world, the narrowest legal width: the result must stay exact);
    `packed`: the bounds arrive as one UNSORTED payload per shard, as from the fused CUDA kernel"""
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), exact, pipelined, str(tmp_path), lb_width, packed), nprocs=world,
             join=True)
    r0, r1 = (np.load(tmp_path / f"rank{r}.npz") for r in range(world))
    assert np.array_equal(r0["vals"], r1["vals"]) and np.array_equal(r0["win"], r1["win"])  # all ranks agree
    assert r0["vals"].shape == (96, 3)
    # single-process ground truth
    N, d, k, ctx, n_top = 96, 32, 6, 8, 3
    p = O.init_params(d, N, k, seed=77)
    x = torch.randn(ctx * 12, d, generator=torch.Generator().manual_seed(78)).to(torch.bfloat16)
    enc = O.encode(p, x.float())
    ref_s, ref_w = O.scan_top_windows(enc.top_acts, enc.top_indices, N, ctx, n_top)
    if exact:  # the threshold exchange reproduces the global TopK mask exactly
        np.testing.assert_array_equal(r0["vals"], ref_s)
        np.testing.assert_array_equal(r0["win"], ref_w)
    else:  # shard-local TopK keeps a superset of each token's latents: scores can only grow
        assert (r0["vals"] >= ref_s - 1e-7).all() and not np.array_equal(r0["vals"], ref_s)


def test_bounds_width():
    from saeb200.dist import bounds_width

    assert bounds_width(64, 64, 8) == 24 and bounds_width(64, 64, 4) == 40 and bounds_width(64, 64, 2) == 64
    assert bounds_width(64, 10, 8) == 10           # never wider than the shard's own list
    for k, w in ((64, 8), (6, 2), (256, 8), (5, 3)):
        assert w * bounds_width(k, k, w) >= k      # the union of the shards' lists always holds k bounds


def test_shard_and_token_ranges():
    from saeb200.dist import shard_range, token_slice

    for n, w in ((131072, 8), (100, 3), (7, 8)):
        ranges = [shard_range(n, w, r) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        sizes = [hi - lo for lo, hi in ranges]
        assert max(sizes) - min(sizes) <= 1
        assert [token_slice(n, w, r) for r in range(w)] == ranges
