"""Multi-GPU behind the C ABI (-m gpu): the peer group (brt_peer_*) and the multi-device context (brt_create_multi).
With one visible GPU the degenerate one-rank forms must reproduce brt_render exactly; with two or more the N-device images
are compared with the single-device render of the same sample set (tools/mgpu_check.py, also run under torchrun)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def brt():
    import blenderraytracer_b200 as b
    return b


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def test_peer_group_of_one_rank_equals_brt_render(brt, sample_scene):
    W, H, spp = 200, 120, 12
    rt = brt.RayTracer(W, H, seed=4)
    assert rt.loadFromJSON(sample_scene)
    for denoise in (False, True):
        rt.updateRenderSettings(dict(samples=spp, maxBounces=6, denoising=denoise))
        ref = rt.render(want_linear=True).copy()
        fref, lref = rt.floatData.copy(), rt.linearMean.copy()
        rt.peerConnect([rt.peerAlloc(0, 1)])
        img, fd, lin = np.zeros((H, W, 4), np.uint8), np.zeros((H, W, 4), np.float32), np.zeros((H, W, 4), np.float32)
        for _ in range(3):                                            # epochs alternate the two sum buffers
            rt.peerRender(0, spp, want_float=True, want_linear=True)
            rt.peerFetch(img.ctypes.data, fd.ctypes.data, lin.ctypes.data)
            assert np.array_equal(img, ref)
            assert np.array_equal(lin, lref)
            if not denoise:                                           # with denoise floatData stays the un-filtered image (ray-tracer.js:267-275)
                assert np.array_equal(fd, fref)
        assert rt.stats()["kernel_ms"] > 0
        rt.peerFree()
    # the group is sized from the image: a resize without a new brt_peer_alloc is refused
    rt.peerConnect([rt.peerAlloc(0, 1)])
    rt.resizeCanvas(W + 8, H)
    rt._push_params()
    with pytest.raises(brt.BrtError):
        rt.peerRender(0, spp)
    rt.peerFree()


def test_multi_device_context_with_one_device(brt, sample_scene):
    W, H = 160, 96
    a = brt.RayTracer(W, H, device=0, seed=2)
    b = brt.RayTracer(W, H, devices=[0], seed=2)
    assert b.deviceCount() == 1
    for r in (a, b):
        assert r.loadFromJSON(sample_scene)
        r.updateRenderSettings(dict(samples=6, maxBounces=5))
    assert np.array_equal(a.render(), b.render())
    with pytest.raises(brt.BrtError):
        brt.RayTracer(W, H, devices=[0, 0])                          # duplicate device ids
    with pytest.raises(brt.BrtError):
        brt.RayTracer(W, H, devices=[0, 99])


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs (run under gpurun --gpus 2)")
def test_multi_device_context_matches_single_device():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "mgpu_check.py"), "--inprocess", str(min(_n_gpus(), 8))],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0 and "MGPU_CHECK OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs (run under gpurun --gpus 2)")
def test_one_process_per_gpu_matches_single_device():
    n = min(_n_gpus(), 8)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                          "--master-port", "29631", os.path.join(ROOT, "tools", "mgpu_check.py")],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0 and "MGPU_CHECK OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
