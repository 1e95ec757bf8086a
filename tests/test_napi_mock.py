"""The Node N-API addon (napi/brt_addon.c), loaded and driven by a mock N-API host (napi/mock_node_host.c) because this
image has no Node.js.  The host implements the napi_* functions the addon imports, dlopen()s brt_addon.node and calls
create -> loadSceneJSON -> setRenderParams -> render(Uint8ClampedArray, onProgress) -> Promise, as napi/raytracer_gpu.mjs does."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAPI = os.path.join(ROOT, "napi")


@pytest.fixture(scope="module")
def built():
    host, addon = os.path.join(NAPI, "mock_node_host"), os.path.join(NAPI, "brt_addon.node")
    subprocess.check_call(["gcc", "-std=c11", "-O1", "-w", "-rdynamic", "-I", NAPI, "-I", os.path.join(ROOT, "include"),
                           os.path.join(NAPI, "mock_node_host.c"), "-ldl", "-o", host])
    # the real link: brt_* resolved against libbrt.so (rpath), napi_* left for the host to provide
    subprocess.check_call(["gcc", "-std=c11", "-O1", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), os.path.join(NAPI, "brt_addon.c"),
                           "-L", os.path.join(ROOT, "blenderraytracer_b200"), "-lbrt", "-Wl,-rpath," + os.path.join(ROOT, "blenderraytracer_b200"),
                           "-o", addon])
    return host, addon


def test_addon_loads_registers_and_fails_loudly_without_a_gpu(built, tmp_path):
    import torch
    host, addon = built
    out = subprocess.run([host, addon, os.path.join(ROOT, "tests", "golden", "sample_scene.json"), "32", "24", "2", "3", "1", str(tmp_path / "o.rgba")],
                         capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert out.returncode == 0, out.stderr
    else:
        # module registration and the exported `create` ran; without a device it throws BRT_E_CUDA (no CPU fallback)
        assert out.returncode == 4 and "create threw BRT_E_CUDA" in out.stderr and "no CPU fallback" in out.stderr


@pytest.mark.gpu
def test_addon_render_equals_python_binding(built, tmp_path):
    """Same scene, seed, settings and spp batching through both bindings of the same C ABI: identical bytes."""
    import blenderraytracer_b200 as brt
    host, addon = built
    W, H, spp, depth, seed = 160, 100, 8, 6, 7
    scene = os.path.join(ROOT, "tests", "golden", "sample_mesh.json")
    raw = tmp_path / "o.rgba"
    out = subprocess.run([host, addon, scene, str(W), str(H), str(spp), str(depth), str(seed), str(raw)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "hasCamera=1" in out.stdout and "render resolved" in out.stdout
    calls = int(out.stdout.split("onProgress calls=")[1].split()[0])
    assert calls == 4 and "last=1.000" in out.stdout                       # 3 batch callbacks + the final onProgress(1.0)
    got = np.fromfile(raw, dtype=np.uint8).reshape(H, W, 4)
    rt = brt.RayTracer(W, H, seed=seed)
    assert rt.loadFromJSON(open(scene).read())
    rt.updateRenderSettings(dict(samples=spp, maxBounces=depth))
    rt.sppBatch = spp // 4
    want = rt.render(onProgress=lambda f: None)
    assert np.array_equal(got, want)
    assert got[..., 3].min() == 255 and got[..., :3].std() > 10
    # create([0]) / create([0, 1]): one context spanning the listed GPUs (brt_create_multi) behind the same render() call
    import torch
    for devs in (["0"], ["0", "1"]):
        if len(devs) > torch.cuda.device_count():
            continue
        raw2 = tmp_path / f"o{len(devs)}.rgba"
        out = subprocess.run([host, addon, scene, str(W), str(H), str(spp), str(depth), str(seed), str(raw2)], capture_output=True, text=True,
                             timeout=300, env=dict(os.environ, BRT_MOCK_DEVICES=",".join(devs)))
        assert out.returncode == 0, out.stderr
        assert f"devices={len(devs)}" in out.stdout and "onProgress calls=4" in out.stdout
        got2 = np.fromfile(raw2, dtype=np.uint8).reshape(H, W, 4)
        assert np.abs(got2.astype(int) - want.astype(int)).max() <= (0 if len(devs) == 1 else 1)
