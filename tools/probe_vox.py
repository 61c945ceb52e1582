#!/usr/bin/env python
"""Voxelizer variants on ONE GPU, same 2M-triangle sphere: P=1 (no partition enumeration), P=8 (inline enumeration),
and the remote-staging kernel with all ranks as contexts of this process. Prints ms_vox_small (CUDA events)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from ooc_svo_builder_b200 import SvoBuilder, meshgen, sharded  # noqa: E402

which = sys.argv[1:] or ["p1", "p8", "w2", "w8"]
base = meshgen.displaced_sphere(1000, 1000, seed=1, length=1.0)


def single(g, length):
    sb = SvoBuilder(0)
    d = torch.from_numpy(base.tris).cuda()
    sb.set_triangles(d)
    prm = sb.make_params(length, g, False)
    best = 1e9
    for _ in range(5):
        sb.partition(prm, want_counts=False); sb.voxelize(); sb.build()
        best = min(best, sb.stats()["ms_vox_small"])
    st = sb.stats()
    sb.close()
    return best, st["n_voxels"]


def remote(world):
    octants = {2: [0, 4], 4: [0, 2, 4, 6], 8: list(range(8))}[world]
    parts = []
    for i in range(world):
        o = octants[(i + 1) % world]
        off = np.array([(o & 1), (o >> 1) & 1, (o >> 2) & 1], dtype=np.float32)
        parts.append((base.tris.reshape(-1, 3, 3) + off).reshape(-1, 9))
    tris = np.ascontiguousarray(np.concatenate(parts), dtype=np.float32)
    T = tris.shape[0]
    ctxs = [SvoBuilder(0) for _ in range(world)]
    for r, sb in enumerate(ctxs):
        sb.shard_configure(r, world)
    per = (T + world - 1) // world
    wins = [sb.slice_create(per, 9) for sb in ctxs]
    for sb in ctxs:
        sb.slice_attach(wins)
    for r, sb in enumerate(ctxs):
        sb.slice_upload(tris[r * per:(r + 1) * per])
    prm = SvoBuilder.make_params(2.0, 2048, False)
    best = [1e9] * world
    disp = [1e9] * world
    for _ in range(4):
        for sb in ctxs:
            sb.slice_publish(prm, T)
        tables = []
        for sb in ctxs:
            sb.partition(prm, want_counts=False); sb.voxelize()
            t = torch.zeros(sb.shard_table_size(), dtype=torch.int64, device="cuda")
            sb.shard_count(t.data_ptr()); sb.synchronize()
            tables.append(t)
        for sb, t in zip(ctxs, tables):
            sb.shard_exchange(t.data_ptr())
        for r, sb in enumerate(ctxs):
            sb.shard_emit(tables[r].data_ptr())
            st = sb.stats()
            best[r] = min(best[r], st["ms_vox_small"]); disp[r] = min(disp[r], st["ms_dispatch"])
    for sb in ctxs:
        sb.close()
    return best, disp


for w in which:
    if w == "p1":
        print("P=1  g=1024 k_vox_small<enum=0>           ms", single(1024, 1.0), flush=True)
    elif w == "p8":
        print("P=8  g=2048 k_vox_small<enum=1>           ms", single(2048, 2.0), flush=True)
    else:
        print("world=%s remote staging (contexts on one GPU) ms_vox_small / ms_publish per rank" % w[1:], remote(int(w[1:])), flush=True)
