"""Peer-memory all-gather of the feature-sharded scan's per-token lists (`saeb_push_gather`, csrc/exchange.cu).

PyTorch's symmetric memory provides the plumbing -- one allocation of the same size on every rank of the box, mapped
into every peer's address space (and, where NVSwitch multicast is available, one multicast mapping) -- and nothing else:
the data movement, the cross-rank signalling and the wait are one kernel of this library.  `PushExchange.gather` is a
drop-in for `dist.all_gather_into_tensor` on [T, m] fp32 lists, ordered on the current CUDA stream.

Layout of the symmetric buffer (identical on every rank):
    [0, 4096)                 flags: uint32 [channels][world], sequence numbers indexed by source rank
    region(channel, slot)     world slabs of max_rows * width(channel) * 4 bytes, 1024-byte aligned
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.distributed as dist

from . import _capi
from ._capi import SaebError, check

FLAGS_BYTES = 4096


class PushExchange:
    def __init__(self, group, device, max_rows: int, widths: Sequence[int], slots: int = 2, multicast: bool = True):
        """Collective over `group`: every rank must construct it with the same arguments.  `widths[c]` = list width
        (fp32 columns) of channel c; `slots` regions per channel alternate between consecutive chunks."""
        import torch.distributed._symmetric_memory as symm

        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise SaebError("PushExchange needs CUDA devices with peer access (one NVLink / NVSwitch box)")
        if self.world * len(widths) * 4 > FLAGS_BYTES:
            raise SaebError("PushExchange: too many channels x ranks for the flag block")
        self.max_rows, self.widths, self.slots = int(max_rows), [int(w) for w in widths], int(slots)
        self.region_off, off = {}, FLAGS_BYTES
        for c, w in enumerate(self.widths):
            nbytes = (self.world * self.max_rows * w * 4 + 1023) // 1024 * 1024
            for s in range(self.slots):
                self.region_off[(c, s)] = off
                off += nbytes
        self.total_bytes = off
        name = self.group.group_name
        enable = getattr(symm, "enable_symm_mem_for_group", None)
        enabled = getattr(symm, "is_symm_mem_enabled_for_group", None)
        if enable is not None and not (enabled is not None and enabled(name)):
            try:
                enable(name)
            except Exception:
                pass   # newer PyTorch enables groups implicitly and only warns here
        with torch.cuda.device(self.device):
            self.buf = symm.empty(self.total_bytes, dtype=torch.uint8, device=self.device)
            self.buf.zero_()
            self.hdl = symm.rendezvous(self.buf, self.group)
            torch.cuda.synchronize(self.device)
        dist.barrier(self.group)   # every rank's flags are zero before anyone publishes a sequence number
        self.peer_ptrs_dev = int(self.hdl.buffer_ptrs_dev)
        mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
        self.multicast_ptr = mc if (multicast and mc != 0) else 0
        self.seq = [0] * len(self.widths)
        self.counters = torch.zeros(len(self.widths), dtype=torch.int32, device=self.device)

    @property
    def transport(self) -> str:
        return "multimem.st (NVSwitch multicast)" if self.multicast_ptr else "peer stores (unicast)"

    def gather(self, t: torch.Tensor, channel: int, slot: int = 0) -> torch.Tensor:
        """t [T, width(channel)] fp32 on this rank -> [world, T, width] view of the local gathered region, valid for
        work enqueued after this call on the current stream (until the region's next use: two chunks later)."""
        w = self.widths[channel]
        if t.dim() != 2 or t.shape[1] != w or t.dtype != torch.float32 or not t.is_cuda:
            raise SaebError(f"PushExchange.gather: need a CUDA float32 [T, {w}] tensor, got {tuple(t.shape)} {t.dtype}")
        T = t.shape[0]
        if T > self.max_rows:
            raise SaebError(f"PushExchange.gather: {T} rows exceed max_rows={self.max_rows}")
        nbytes = T * w * 4
        if nbytes % 16 != 0:
            raise SaebError("PushExchange.gather: T * width must be a multiple of 4 (16-byte vectors)")
        src = t.contiguous()
        off = self.region_off[(channel, slot % self.slots)]
        self.seq[channel] += 1
        L = _capi.lib()
        with torch.cuda.device(self.device):
            check(L.saeb_push_gather(src.data_ptr(), nbytes, self.peer_ptrs_dev, self.world, self.rank, off,
                                     self.multicast_ptr or None, 0, channel, self.seq[channel] & 0xFFFFFFFF,
                                     self.counters[channel:].data_ptr(), torch.cuda.current_stream().cuda_stream),
                  "saeb_push_gather")
        # `src` may be a temporary: keep it alive until the kernel has run on this stream
        src.record_stream(torch.cuda.current_stream())
        return self.buf[off:off + self.world * nbytes].view(torch.float32).view(self.world, T, w)
