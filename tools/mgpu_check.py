"""Multi-GPU parity check (run under torchrun, one rank per GPU): the spp-split image (NCCL reduce and fused P2P
reduce+resolve) equals the single-GPU image of the same sample set within 1 LSB (fp32 summation order only)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import blenderraytracer_b200 as brt
from blenderraytracer_b200.distributed import SppSplitRenderer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
scene = json.load(open(os.path.join(ROOT, "tests", "golden", "sample_mesh.json")))
W, H, spp = 640, 360, 37
ok = True
for mode in ("nccl", "p2p"):
    rt = brt.RayTracer(W, H, device=local, seed=9)
    assert rt.loadFromJSON(scene)
    rt.updateRenderSettings(dict(samples=spp, maxBounces=8))
    sr = SppSplitRenderer(rt, reduce=mode)
    for _ in range(3):                      # repeated steps must not leak state between renders
        sr.step()
    img = sr.image()
    sr.close()
    if rank == 0:
        single = rt.render()
        d = np.abs(img.astype(int) - single.astype(int))
        print(f"[{mode}] world={world} max LSB diff vs single GPU: {d.max()}, differing bytes: {(d>0).mean():.2e}", flush=True)
        ok = ok and d.max() <= 1
    dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("MGPU_CHECK", "OK" if ok else "FAIL", flush=True)
sys.exit(0 if ok else 1)
