#!/usr/bin/env python
"""Runs the BASELINE.json configs 3-5 (and 2) once on one GPU and prints counts + device stage times.
Not a bench line: a scale check of the path (memory, 64-bit indexing, large launches) with the
size-independent invariants of SURVEY.md §4 checked on the result.

    python tools/scale_run.py c3 c4 c5 [--ref]     # --ref: also run the reference CPU binary (slow)
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from ooc_svo_builder_b200 import SvoBuilder, meshgen  # noqa: E402

CFG = {
    "c2": ("c2_displaced_sphere_1024", 1024, 2048),
    "c3": ("c3_terrain_2048_payload", 2048, 2048),
    "c4": ("c4_sphere_4096", 4096, 2048),
    "c5": ("c5_shell_8192", 8192, 2048),
}


def main():
    names = [a for a in sys.argv[1:] if not a.startswith("-")]
    out = {}
    sb = SvoBuilder(0)
    for n in names:
        cfg, g, lim = CFG[n]
        t = time.time()
        mesh = meshgen.make(cfg)
        tgen = time.time() - t
        payload = mesh.payload
        prm = sb.make_params(mesh.length, g, payload, memory_limit_mb=lim)
        res = {}
        for rep in range(2):                      # second run = steady state (buffers allocated, pyramid sparse-cleared)
            t = time.time()
            sb.set_triangles(mesh.tris)
            sb.partition(prm, want_counts=False)
            sb.voxelize()
            nv, nn, nd = sb.build()
            wall = time.time() - t
            st = sb.stats()
            res = dict(st, wall_s=wall, gen_s=tgen, n_triangles=mesh.n_triangles)
        # invariants on the head and the tail of the node file
        k = min(nn, 2_000_000)
        head = np.frombuffer(sb.fetch_nodes(0, k).tobytes(), dtype=np.uint64).reshape(-1, 3)
        tail = np.frombuffer(sb.fetch_nodes(nn - k, k).tobytes(), dtype=np.uint64).reshape(-1, 3)
        leaf = head[:, 2] == np.uint64(0xFFFFFFFFFFFFFFFF)
        assert (head[leaf, 1] == 0).all()
        if not payload:
            assert (head[leaf, 0] == 1).all() and (head[~leaf, 0] == 0).all()
        off = tail[-1:].view(np.int8).reshape(-1, 24)[:, 16:]
        cnt_root = int((off != -1).sum())
        assert int(tail[-1, 1]) == nn - 1 - cnt_root, "root must be last with its children right before it"
        inner = head[~leaf]
        assert (inner[:, 1] < np.uint64(nn)).all()
        dev_ms = res["ms_partition"] + res["ms_voxelize"] + res["ms_build"] + res["ms_clear"]
        res["device_ms"] = dev_ms
        res["triangles_per_s_device"] = mesh.n_triangles / (dev_ms * 1e-3)
        res["voxels_per_s_device"] = nv / (dev_ms * 1e-3)
        res["octree_build_GBs"] = (8 * nv + 24 * nn) / max(res["ms_emit_leaf"], 1e-9) / 1e6
        out[n] = res
        print(n, json.dumps(res), flush=True)
        del mesh
    sb.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "scale_run.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
