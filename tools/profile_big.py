#!/usr/bin/env python
"""Two builds of a full-size BASELINE config on one GPU (the second one is the steady state), for an ncu launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c5.csv python tools/profile_big.py c5
Prints the library's own stage times of the second build as well."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ooc_svo_builder_b200 import SvoBuilder, meshgen  # noqa: E402

CFG = {"c2": ("c2_displaced_sphere_1024", 1024), "c3": ("c3_terrain_2048_payload", 2048), "c4": ("c4_sphere_4096", 4096), "c5": ("c5_shell_8192", 8192)}
name = sys.argv[1] if len(sys.argv) > 1 else "c5"
cfg, g = CFG[name]
mesh = meshgen.make(cfg)
sb = SvoBuilder(0)
prm = sb.make_params(mesh.length, g, mesh.payload)
sb.set_triangles(mesh.tris)
for i in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    sb.partition(prm, want_counts=False)
    sb.voxelize()
    nv, nn, nd = sb.build()
    st = sb.stats()
print(name, json.dumps({k: st[k] for k in ("n_voxels", "n_nodes", "n_bricks", "n_tiles1", "n_brick_records", "ms_voxelize", "ms_vox_small", "ms_compact", "ms_build", "ms_emit", "ms_emit_leaf", "ms_clear", "speculative", "kernel_launches")}))
sb.close()
