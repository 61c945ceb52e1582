#!/usr/bin/env python
"""Strong-scaling efficiency of the remote-staging voxelizer WITHOUT NVLink effects: the same mesh voxelized by one
context, and by `world` contexts of this process (all on one GPU, sequentially). If the per-rank times add up to the
single-context time, list order / block granularity / persistence cost nothing."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from ooc_svo_builder_b200 import SvoBuilder, meshgen, sharded  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2500
g = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
world = int(sys.argv[3]) if len(sys.argv) > 3 else 8
mesh = meshgen.displaced_sphere(n, n, seed=3)
T = mesh.n_triangles
prm = SvoBuilder.make_params(mesh.length, g, False)

sb = SvoBuilder(0)
sb.set_triangles(torch.from_numpy(mesh.tris).cuda())
best = 1e9
for _ in range(3):
    sb.partition(prm, want_counts=False); sb.voxelize(); sb.build()
    best = min(best, sb.stats()["ms_vox_small"])
print("single context: T=%d g=%d ms_vox_small %.3f" % (T, g, best), flush=True)
sb.close()

# the same work through the remote-staging kernel with ONE rank: isolates persistent-kernel efficiency from sharding
sb = SvoBuilder(0)
sb.shard_configure(0, 1)
w = sb.slice_create(T, 9); sb.slice_attach([w]); sb.slice_upload(mesh.tris)
best = 1e9
for _ in range(3):
    sb.slice_publish(prm, T); sb.partition(prm, want_counts=False); sb.voxelize(); sb.build()
    best = min(best, sb.stats()["ms_vox_small"])
print("one rank through the remote-staging kernel: ms_vox_small %.3f, ms_publish %.3f" % (best, sb.stats()["ms_dispatch"]), flush=True)
sb.close()

ctxs = [SvoBuilder(0) for _ in range(world)]
for r, c in enumerate(ctxs):
    c.shard_configure(r, world)
per = (T + world - 1) // world
wins = [c.slice_create(per, 9) for c in ctxs]
for c in ctxs:
    c.slice_attach(wins)
for r, c in enumerate(ctxs):
    c.slice_upload(mesh.tris[r * per:(r + 1) * per])
vs = [1e9] * world; ds = [1e9] * world
for _ in range(3):
    for c in ctxs:
        c.slice_publish(prm, T)
    tables = []
    for c in ctxs:
        c.partition(prm, want_counts=False); c.voxelize()
        t = torch.zeros(c.shard_table_size(), dtype=torch.int64, device="cuda")
        c.shard_count(t.data_ptr()); c.synchronize()
        tables.append(t)
    for c, t in zip(ctxs, tables):
        c.shard_exchange(t.data_ptr())
    for r, c in enumerate(ctxs):
        c.shard_emit(tables[r].data_ptr())
        st = c.stats()
        vs[r] = min(vs[r], st["ms_vox_small"]); ds[r] = min(ds[r], st["ms_dispatch"])
print("world %d: ms_vox_small per rank %s  sum %.3f" % (world, [round(v, 3) for v in vs], sum(vs)))
print("          ms_publish per rank %s" % [round(v, 3) for v in ds])
for c in ctxs:
    c.close()
