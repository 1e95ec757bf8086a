"""Multi-GPU render: samples-per-pixel split, one process per GPU.

Rank r of N traces global sample indices ``sample_range(spp, r, N)`` of EVERY pixel into its own fp32 RGBA sum buffer
(alpha carries the sample count).  The Philox counter is (pixel, global sample index, bounce), so the union of the
samples is the same set for any N.  One exchange step follows:

* ``reduce="fused"`` (default) — libbrt's peer group (``brt_peer_*``): the exchange blocks are mapped once through CUDA IPC
  handles; per step every rank launches ONE kernel after its path tracer that publishes an epoch flag, waits for the peers'
  flags with acquire loads over NVLink (no host barrier, no NCCL call in the step), pulls its row stripe of all N sum buffers
  with 128-bit loads, sums in fixed rank order (deterministic), tone-maps and stores RGBA8 (4x fewer bytes than the sums)
  straight into rank 0's image.  ``torch.distributed`` is used only to pass the 64-byte handles around at start-up;
* ``reduce="nccl"`` — the library-collective fallback: ``reduce_scatter`` of the fp32 sums (every rank receives the sum of
  its row stripe), each rank resolves its stripe (``brt_reduce_resolve_peers`` on one buffer), ``all_gather`` of the RGBA8
  stripes.

The reference (js/ray-tracer.js) is single-threaded; this replaces nothing of it beyond the `for s` loop (:202).
"""
from __future__ import annotations

from typing import Tuple


def sample_range(spp: int, rank: int, world: int) -> Tuple[int, int]:
    """(first global sample index, count) of `rank`: contiguous ranges, remainder to the low ranks."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    spp = max(0, int(spp))
    base, rem = divmod(spp, world)
    begin = rank * base + min(rank, rem)
    return begin, base + (1 if rank < rem else 0)


def row_stripe(height: int, rank: int, world: int) -> Tuple[int, int]:
    """[row_begin, row_end) resolved by `rank` in the fused exchange (same split as the library's)."""
    b, c = sample_range(height, rank, world)
    return b, b + c


def equal_stripe(height: int, rank: int, world: int) -> Tuple[int, int, int]:
    """(rows per rank S, row_begin, row_end) of the NCCL path: reduce_scatter needs equal parts, so the image is padded to
    S * world rows and the last stripes are clipped to the real height."""
    s = -(-int(height) // world)
    return s, min(height, rank * s), min(height, (rank + 1) * s)


def reduce_sums(accum, dst: int = 0, group=None):
    """Sum the per-rank accumulation buffers onto `dst` (in place).  Works on any torch.distributed backend
    (NCCL on GPUs; gloo in the CPU tests)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return accum


def reduce_scatter_sums(accum_padded, out_stripe, group=None):
    """out_stripe = sum over ranks of this rank's equal row stripe of accum_padded ((S*world, W, 4) fp32)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        if dist.get_backend(group) == "gloo":                        # gloo has no reduce_scatter: all_reduce + slice (CPU tests)
            dist.all_reduce(accum_padded, op=dist.ReduceOp.SUM, group=group)
            s = out_stripe.shape[0]
            r = dist.get_rank(group)
            out_stripe.copy_(accum_padded[r * s:(r + 1) * s])
        else:
            dist.reduce_scatter_tensor(out_stripe, accum_padded, op=dist.ReduceOp.SUM, group=group)
    else:
        out_stripe.copy_(accum_padded[:out_stripe.shape[0]])
    return out_stripe


class SppSplitRenderer:
    """Drives one ``RayTracer`` per rank.  ``step()`` = zero the sums, trace this rank's samples, exchange, resolve;
    afterwards rank 0 holds the RGBA8 image on its device (``image()`` / ``fetch_into()`` bring it to the host)."""

    def __init__(self, rt, reduce: str = "fused", group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rt, self.group = rt, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if reduce == "p2p":                                          # round-1 name of the fused path
            reduce = "fused"
        if reduce not in ("fused", "nccl"):
            raise ValueError("reduce must be 'fused' or 'nccl'")
        self.reduce = reduce
        self.dev = torch.device("cuda", torch.cuda.current_device())
        W, H = rt.width, rt.height
        self.nbytes = W * H * 16
        rt.setStream(torch.cuda.current_stream().cuda_stream)
        rt._push_params()
        self._open = False
        if self.reduce == "fused":
            handle = rt.peerAlloc(self.rank, self.world)
            handles = [None] * self.world
            if self.world > 1:
                dist.all_gather_object(handles, handle, group=group)
            else:
                handles = [handle]
            rt.peerConnect(handles)
            self._open = True
            self.accum = self.rgba = None
        else:
            self.S, self.r0, self.r1 = equal_stripe(H, self.rank, self.world)
            self.accum = torch.zeros((self.S * self.world, W, 4), dtype=torch.float32, device=self.dev)
            self.accum_ptr = self.accum.data_ptr()
            self.stripe = torch.zeros((self.S, W, 4), dtype=torch.float32, device=self.dev)
            self.rgba_stripe = torch.zeros((self.S, W, 4), dtype=torch.uint8, device=self.dev)
            self.rgba = torch.zeros((self.S * self.world, W, 4), dtype=torch.uint8, device=self.dev)

    def close(self):
        if self._open:
            self.rt.synchronize()
            if self.dist.is_initialized() and self.world > 1:
                self.dist.barrier(group=self.group)                  # nobody frees while a peer still maps it
            self.rt.peerFree()
            self._open = False

    def spp(self) -> int:
        return 1 if self.rt.antiAliasing == "none" else int(self.rt.samples)

    def step(self, spp_total=None):
        """One render of the whole image across all ranks, asynchronous on torch's current stream."""
        rt = self.rt
        rt.setStream(self.torch.cuda.current_stream().cuda_stream)   # NCCL and libbrt must share the stream step() runs under
        begin, count = sample_range(self.spp() if spp_total is None else spp_total, self.rank, self.world)
        if self.reduce == "fused":
            rt.peerRender(begin, count)
            return
        W, H = rt.width, rt.height
        rt.deviceMemset(self.accum_ptr, 0, self.nbytes)
        rt.renderAccumulate(self.accum_ptr, begin, count)
        reduce_scatter_sums(self.accum, self.stripe, self.group)
        if self.r1 > self.r0:
            # resolve rows [r0, r1) of the image from the stripe buffer: the kernel indexes whole-image pixels, so hand it the
            # stripe's base moved back by r0 rows (only rows r0..r1 are touched)
            rt.reduceResolvePeers([self.stripe.data_ptr() - self.r0 * W * 16], self.r0, self.r1,
                                  self.rgba_stripe.data_ptr() - self.r0 * W * 4)
        if self.world > 1:
            self.dist.all_gather_into_tensor(self.rgba, self.rgba_stripe, group=self.group)
        else:
            self.rgba[:self.S].copy_(self.rgba_stripe)

    def fetch_into(self, host_ptr):
        """Rank 0: copy the finished RGBA8 image into host memory at `host_ptr` (blocking); other ranks: synchronise."""
        rt = self.rt
        if self.reduce == "fused":
            rt.peerFetch(host_ptr if self.rank == 0 else None)
        else:
            if self.rank == 0:
                rt.copyToHost(host_ptr, self.rgba.data_ptr(), rt.width * rt.height * 4)
            else:
                rt.synchronize()

    def image(self):
        """Rank 0: the (H, W, 4) uint8 image as a host numpy array; other ranks: None (after synchronising)."""
        import numpy as np
        out = np.empty((self.rt.height, self.rt.width, 4), np.uint8) if self.rank == 0 else None
        self.fetch_into(out.ctypes.data if out is not None else None)
        return out

    def launches_per_step(self) -> int:
        """Kernels of OURS launched per step on this rank (memsets and NCCL's kernels are not counted): the path tracer, the
        fused exchange + resolve, and on rank 0 the wait for the peers' stripes."""
        if self.reduce == "fused":
            return 2 + (1 if self.rank == 0 else 0)
        return 1 + (1 if self.r1 > self.r0 else 0)
