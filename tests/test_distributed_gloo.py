"""The N > 1 path on CPU: world_size-2 `gloo` processes exercise the host side of the spp split — sample partition, the
sum-reduce of per-rank accumulation buffers through blenderraytracer_b200.distributed.reduce_sums, and the ÷spp resolve
identity — with the oracle standing in for the device accumulation (libbrt has no CPU renderer)."""
import os
import socket
import sys

import numpy as np
import pytest

from blenderraytracer_b200.distributed import equal_stripe, reduce_scatter_sums, reduce_sums, row_stripe, sample_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sample_range_partitions_exactly():
    for spp in (0, 1, 2, 7, 16, 255, 256, 4096):
        for world in (1, 2, 3, 4, 8):
            parts = [sample_range(spp, r, world) for r in range(world)]
            assert sum(c for _, c in parts) == spp
            pos = 0
            for b, c in parts:                                   # contiguous, ordered, remainder to the low ranks
                assert b == pos and c in (spp // world, spp // world + 1)
                pos += c
            counts = [c for _, c in parts]
            assert counts == sorted(counts, reverse=True)
    with pytest.raises(ValueError):
        sample_range(8, 2, 2)


def test_row_stripes_cover_the_image():
    for h in (1, 7, 1080, 2160):
        for world in (1, 2, 4, 8):
            rows = [row_stripe(h, r, world) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == h
            assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))


def test_equal_stripes_cover_the_padded_image():
    for h in (1, 7, 1080, 2160, 401):
        for world in (1, 2, 3, 4, 8):
            parts = [equal_stripe(h, r, world) for r in range(world)]
            s = parts[0][0]
            assert s * world >= h and all(p[0] == s for p in parts)
            assert parts[0][1] == 0 and max(p[2] for p in parts) == h
            assert all(a[2] == b[1] or b[1] == h for a, b in zip(parts, parts[1:]))     # contiguous until the image ends
            assert sum(p[2] - p[1] for p in parts) == h


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, spp, W, H, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import json
    import torch
    import torch.distributed as dist
    from oracle.oracle import OracleRayTracer
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        scene = json.load(open(os.path.join(ROOT, "tests", "golden", "sample_scene.json")))
        begin, count = sample_range(spp, rank, world)
        o = OracleRayTracer(W, H, seed=17, threads=2)
        assert o.loadFromJSON(scene)
        o.updateRenderSettings(dict(samples=max(count, 1), maxBounces=4))
        o.sampleBegin = begin
        accum = torch.zeros((H, W, 4), dtype=torch.float64)
        if count > 0:
            o.render()
            accum[..., :3] = torch.from_numpy(o.linear[..., :3] * count)     # the device kernel accumulates SUMS, alpha = count
            accum[..., 3] = count
        # the NCCL fallback's exchange: reduce_scatter of the padded sums -> every rank owns the sum of its equal row stripe
        S, r0, r1 = equal_stripe(H, rank, world)
        padded = torch.zeros((S * world, W, 4), dtype=torch.float64)
        padded[:H] = accum
        stripe = torch.zeros((S, W, 4), dtype=torch.float64)
        reduce_scatter_sums(padded, stripe)
        stripes = [torch.zeros_like(stripe) for _ in range(world)]
        dist.all_gather(stripes, stripe)
        reduce_sums(accum, dst=0)
        if rank == 0:
            assert torch.equal(torch.cat(stripes)[:H], accum), "reduce_scatter + all_gather of stripes differs from reduce-to-root"
            q.put(accum.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("spp", [5, 8])
def test_spp_split_reduce_equals_single_rank(spp):
    """Two gloo ranks each trace their sample range; the reduced sums ÷ spp equal one rank tracing all samples
    (same Philox-keyed sample set, float64 sums: equal to rounding)."""
    import json
    import torch.multiprocessing as mp
    from oracle.oracle import OracleRayTracer
    W, H, world = 48, 32, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, spp, W, H, q)) for r in range(world)]
    for p in procs:
        p.start()
    total = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.all(total[..., 3] == spp)
    scene = json.load(open(os.path.join(ROOT, "tests", "golden", "sample_scene.json")))
    o = OracleRayTracer(W, H, seed=17, threads=2)
    assert o.loadFromJSON(scene)
    o.updateRenderSettings(dict(samples=spp, maxBounces=4))
    o.render()
    np.testing.assert_allclose(total[..., :3] / spp, o.linear[..., :3], rtol=1e-12, atol=1e-14)
