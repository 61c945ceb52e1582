#!/usr/bin/env python
"""One step of each flavour, for ncu: C2 binary (P=1), a payload multi-partition step, a list-mode step and a
sharded (world=2, both ranks on this GPU) step -- so that every kernel of the library appears in the capture."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from ooc_svo_builder_b200 import SvoBuilder, meshgen, sharded  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
sb = SvoBuilder(0)
if which == "c2":
    mesh = meshgen.displaced_sphere(1000, 1000, seed=1)
    d = torch.from_numpy(mesh.tris).cuda()
    sb.set_triangles(d)
    prm = sb.make_params(mesh.length, 1024, False)
    for _ in range(3):
        sb.partition(prm, want_counts=False); sb.voxelize(); sb.build()
else:
    mesh = meshgen.terrain(700, seed=2)                    # ~1 M payload triangles
    for _ in range(2):
        sb.run(mesh.tris, mesh.length, 1024, memory_limit_mb=200, fetch=False)            # 8 partitions, payload
    soup = meshgen.random_soup(200000, seed=3, large_frac=0.01)
    os.environ["SVO_PARTITION_LISTS"] = "1"
    sb.run(soup.tris, soup.length, 1024, memory_limit_mb=200, fetch=False)                # list-based partitioner, medium / large queues
    os.environ.pop("SVO_PARTITION_LISTS")
    sb.run(mesh.tris, mesh.length, 512, levels=True, fetch=False)                         # -levels
    sharded.run_single_process(soup.tris, soup.length, 1024, 2, memory_limit_mb=200, fetch=False)
sb.close()
