"""Multi-GPU parity check on real peers (gate G6).  Two forms:

  torchrun --nproc-per-node N tools/mgpu_check.py      one process per GPU: the fused peer exchange and the NCCL fallback
  python tools/mgpu_check.py --inprocess N              one process, N GPUs behind ONE ctx (brt_create_multi)

Each image must equal the single-GPU render of the same sample set within 1 LSB (the ranks' fp32 partial sums are added in
rank order instead of sample order: fp32 summation order only).  Prints MGPU_CHECK OK / FAIL; the log is kept under profiles/."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import blenderraytracer_b200 as brt

scene = json.load(open(os.path.join(ROOT, "tests", "golden", "sample_mesh.json")))
W, H, spp = 640, 360, 37


def report(tag, img, single):
    d = np.abs(img.astype(int) - single.astype(int))
    print(f"[{tag}] max LSB diff vs single GPU: {d.max()}, differing bytes: {(d > 0).sum()} of {d.size}", flush=True)
    return d.max() <= 1


def inprocess(n):
    ok = True
    one = brt.RayTracer(W, H, device=0, seed=9)
    assert one.loadFromJSON(scene)
    for denoise in (False, True):
        for batch in (0, 5):
            settings = dict(samples=spp, maxBounces=8, denoising=denoise)
            one.updateRenderSettings(settings)
            one.sppBatch = batch
            single = one.render(want_linear=True).copy()
            lin1 = one.linearMean.copy()
            multi = brt.RayTracer(W, H, devices=list(range(n)), seed=9)
            assert multi.deviceCount() == n
            assert multi.loadFromJSON(scene)
            multi.updateRenderSettings(settings)
            multi.sppBatch = batch
            multi.preview = batch > 0
            calls = []
            for rep in range(2):                        # repeated renders must not leak state between epochs
                img = multi.render(onProgress=(lambda f, im=None: calls.append(f)) if batch else None, want_linear=True)
            ok = report(f"in-process n={n} denoise={denoise} batch={batch}", img, single) and ok
            rel = np.abs(multi.linearMean[..., :3] - lin1[..., :3]).max() / max(1e-9, np.abs(lin1[..., :3]).max())
            print(f"    linear mean max rel diff {rel:.2e}; progress calls {len(calls)}", flush=True)
            ok = ok and rel < 1e-5 and (batch == 0 or (len(calls) >= 2 and calls[-1] == 1.0))
            multi.close()
    print("MGPU_CHECK", "OK" if ok else "FAIL", flush=True)
    return ok


def per_process():
    import torch.distributed as dist
    from blenderraytracer_b200.distributed import SppSplitRenderer
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for mode in ("fused", "nccl"):
        for denoise in (False, True):
            rt = brt.RayTracer(W, H, device=local, seed=9)
            assert rt.loadFromJSON(scene)
            rt.updateRenderSettings(dict(samples=spp, maxBounces=8, denoising=denoise))
            if mode == "nccl" and denoise:
                rt.close()
                continue                                # the NCCL fallback resolves stripes: the 3x3 denoise is a fused-path feature
            sr = SppSplitRenderer(rt, reduce=mode)
            for _ in range(3):                          # repeated steps must not leak state between epochs
                sr.step()
            img = sr.image()
            sr.close()
            if rank == 0:
                rt.setStream(None)
                single = rt.render()
                ok = report(f"{mode} world={world} denoise={denoise}", img, single) and ok
            dist.barrier()
            rt.close()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_CHECK", "OK" if ok else "FAIL", flush=True)
    return ok


if __name__ == "__main__":
    if "--inprocess" in sys.argv:
        sys.exit(0 if inprocess(int(sys.argv[sys.argv.index("--inprocess") + 1])) else 1)
    sys.exit(0 if per_process() else 1)
