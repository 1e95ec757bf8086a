"""Multi-GPU render: samples-per-pixel split, one process per GPU.

Rank r of N traces global sample indices ``sample_range(spp, r, N)`` of EVERY pixel into its own fp32 RGBA sum buffer
(alpha carries the sample count).  The Philox counter is (pixel, global sample index, bounce), so the union of the
samples is the same set for any N.  One exchange step follows:

* ``reduce="nccl"`` — ``torch.distributed.reduce(SUM)`` of the W*H*4 fp32 buffer onto rank 0 over NCCL
  (NVLink 5 / NVSwitch), then rank 0 runs the resolve kernel (÷spp → tone map → gamma → RGBA8);
* ``reduce="p2p"``  — the fused collective+consumer kernel ``brt_reduce_resolve_peers``: the accumulation buffers are
  exchanged as CUDA IPC handles once, every rank pulls its row stripe from all peers with plain 128-bit loads over
  NVLink, sums in fixed rank order (deterministic), resolves and stores RGBA8 (4x fewer bytes than the sums) straight
  into rank 0's output buffer.

The reference (js/ray-tracer.js) is single-threaded; this replaces nothing of it beyond the `for s` loop (:202).
"""
from __future__ import annotations

from typing import Optional, Tuple


def sample_range(spp: int, rank: int, world: int) -> Tuple[int, int]:
    """(first global sample index, count) of `rank`: contiguous ranges, remainder to the low ranks."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    spp = max(0, int(spp))
    base, rem = divmod(spp, world)
    begin = rank * base + min(rank, rem)
    return begin, base + (1 if rank < rem else 0)


def row_stripe(height: int, rank: int, world: int) -> Tuple[int, int]:
    """[row_begin, row_end) resolved by `rank` in the fused p2p reduce."""
    b, c = sample_range(height, rank, world)
    return b, b + c


def reduce_sums(accum, dst: int = 0, group=None):
    """Sum the per-rank accumulation buffers onto `dst` (in place).  Works on any torch.distributed backend
    (NCCL on GPUs; gloo in the CPU tests)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return accum


class SppSplitRenderer:
    """Drives one ``RayTracer`` per rank.  ``step()`` = zero the sums, trace this rank's samples, exchange, resolve."""

    def __init__(self, rt, reduce: str = "nccl", group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rt, self.group = rt, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.reduce = reduce if self.world > 1 else "none"
        if self.reduce not in ("none", "nccl", "p2p"):
            raise ValueError("reduce must be 'nccl' or 'p2p'")
        self.dev = torch.device("cuda", torch.cuda.current_device())
        W, H = rt.width, rt.height
        self.nbytes = W * H * 16
        rt.setStream(torch.cuda.current_stream().cuda_stream)
        rt._push_params()
        self.rgba = torch.zeros((H, W, 4), dtype=torch.uint8, device=self.dev) if self.rank == 0 else None
        self._opened = []
        if self.reduce == "p2p":
            # library-owned (cudaMalloc) buffers so a CUDA IPC handle names them; exchanged once
            self.accum_ptr, h_acc = rt.sharedAlloc(self.nbytes)
            self._rgba_ptr, h_rgba = rt.sharedAlloc(W * H * 4) if self.rank == 0 else (None, None)
            handles = [None] * self.world
            dist.all_gather_object(handles, (h_acc, h_rgba), group=group)
            self._peers = []
            for r, (ha, hr) in enumerate(handles):
                if r == self.rank:
                    self._peers.append(self.accum_ptr)
                else:
                    self._peers.append(rt.sharedOpen(ha)); self._opened.append(self._peers[-1])
                if r == 0:
                    if self.rank == 0:
                        self._root_rgba = self._rgba_ptr
                    else:
                        self._root_rgba = rt.sharedOpen(hr); self._opened.append(self._root_rgba)
            self.accum = None
        else:
            self.accum = torch.zeros((H, W, 4), dtype=torch.float32, device=self.dev)
            self.accum_ptr = self.accum.data_ptr()

    def close(self):
        if self.reduce == "p2p" and self.accum_ptr:
            self.rt.synchronize()
            for p in self._opened:
                self.rt.sharedClose(p)
            self._opened = []
            if self.dist.is_initialized():
                self.dist.barrier(group=self.group)          # nobody frees while a peer still maps it
            self.rt.sharedFree(self.accum_ptr)
            if self._rgba_ptr:
                self.rt.sharedFree(self._rgba_ptr)
            self.accum_ptr = None

    def spp(self) -> int:
        return 1 if self.rt.antiAliasing == "none" else int(self.rt.samples)

    def step(self):
        """One render of the whole image across all ranks, asynchronous on torch's current stream.  Afterwards rank 0
        holds the RGBA8 image (``self.rgba``; in p2p mode call ``image()`` to fetch it from the shared buffer)."""
        rt = self.rt
        begin, count = sample_range(self.spp(), self.rank, self.world)
        rt.deviceMemset(self.accum_ptr, 0, self.nbytes)
        rt.renderAccumulate(self.accum_ptr, begin, count)
        if self.reduce == "p2p":
            # all ranks must finish tracing before anyone pulls peer sums, and finish pulling before the next
            # step zeroes them: a stream-ordered barrier on either side of the fused kernel
            self._stream_barrier()
            r0, r1 = row_stripe(rt.height, self.rank, self.world)
            rt.reduceResolvePeers(self._peers, r0, r1, self._root_rgba)
            self._stream_barrier()
        else:
            if self.reduce == "nccl":
                reduce_sums(self.accum, 0, self.group)
            if self.rank == 0:
                rt.resolveDevice(self.accum_ptr, self.rgba.data_ptr())

    def _stream_barrier(self):
        if not hasattr(self, "_flag"):
            self._flag = self.torch.zeros(1, dtype=self.torch.float32, device=self.dev)
        self.dist.all_reduce(self._flag, group=self.group)   # NCCL: enqueued on the current stream, no host sync

    def image(self):
        """Rank 0: the (H, W, 4) uint8 image as a host numpy array; other ranks: None."""
        if self.rank != 0:
            return None
        import numpy as np
        self.torch.cuda.current_stream().synchronize()
        if self.reduce == "p2p":
            import ctypes as C
            out = np.empty((self.rt.height, self.rt.width, 4), np.uint8)
            self.rt.copyToHost(out.ctypes.data, self._rgba_ptr, out.nbytes)
            return out
        return self.rgba.cpu().numpy()

    def launches_per_step(self) -> int:
        """Kernels of OURS launched per step on this rank (the memset and NCCL's kernels are not counted)."""
        return 1 + (1 if (self.reduce == "p2p" or self.rank == 0) else 0)
