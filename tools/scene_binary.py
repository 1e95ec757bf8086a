"""Scene JSON (docs/scene_format.md) <-> BRTSCN01 binary container: every mesh's `vertices` / `indices` move into a binary
blob (float64 vertex triples, uint32 indices) referenced by `vertices_bin` / `indices_bin` = {offset, count}; everything
else stays JSON.  SURVEY §8(f) row 4 (exporter-side format: only matters at the 1 M-triangle scale).

    python tools/scene_binary.py scene.json scene.brtscn
"""
from __future__ import annotations

import json
import struct
import sys

import numpy as np


def pack(scene: dict) -> bytes:
    head = json.loads(json.dumps({k: v for k, v in scene.items() if k != "objects"}))
    blob = bytearray()
    objs = []
    for o in scene.get("objects", []):
        if isinstance(o, dict) and isinstance(o.get("type"), str) and o["type"].lower() == "mesh" and "vertices" in o and "indices" in o:
            v = np.asarray(o["vertices"], dtype=np.float64).reshape(-1)
            idx = np.asarray(o["indices"])
            # only meshes whose indices are plain non-negative integers can be stored as uint32; others stay JSON
            if v.size % 3 == 0 and idx.ndim == 1 and idx.size and np.all(idx >= 0) and np.all(idx == np.floor(idx)) and idx.max() < 2 ** 32:
                q = {k: val for k, val in o.items() if k not in ("vertices", "indices")}
                while len(blob) % 8:
                    blob.append(0)
                q["vertices_bin"] = {"offset": len(blob), "count": int(v.size)}
                blob += v.tobytes()
                q["indices_bin"] = {"offset": len(blob), "count": int(idx.size)}
                blob += idx.astype(np.uint32).tobytes()
                objs.append(q)
                continue
        objs.append(o)
    head["objects"] = objs
    js = json.dumps(head).encode()
    out = bytearray(b"BRTSCN01") + struct.pack("<Q", len(js)) + js
    while len(out) % 8:
        out.append(0)
    return bytes(out + blob)


if __name__ == "__main__":
    data = pack(json.load(open(sys.argv[1])))
    open(sys.argv[2], "wb").write(data)
    print(f"{sys.argv[2]}: {len(data)} bytes")
