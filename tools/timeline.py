#!/usr/bin/env python
"""Host-side timeline of a C2 step (SVO_TIMELINE=1 makes the library print wall-clock stamps, us)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["SVO_TIMELINE"] = "1"
import torch
from ooc_svo_builder_b200 import SvoBuilder, meshgen
mesh = meshgen.displaced_sphere(1000, 1000, seed=1)
sb = SvoBuilder(0)
sb.set_triangles(torch.from_numpy(mesh.tris).cuda())
prm = sb.make_params(mesh.length, 1024, False)
for _ in range(6):
    sb.partition(prm, want_counts=False); sb.voxelize(); sb.build()
print({k: round(v, 4) for k, v in sb.stats().items() if k.startswith("ms_")})
sb.close()
