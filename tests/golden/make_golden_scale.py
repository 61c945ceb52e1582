#!/usr/bin/env python
"""Generates tests/golden/golden_scale.json: digests of the UNMODIFIED reference's output files
(oracle/_ref, built from /root/reference by oracle/Makefile) on the full-size BASELINE.json configs.

    python tests/golden/make_golden_scale.py c2 c3 c4 c5      (dev container, CPU; c5 takes ~15 min)

Per config: sha256 of the mesh bytes (so a consumer can tell whether it regenerated the same input), the
.octree header text, sha256 of .octreenodes / .octreedata, the voxel count, and a POSITION-AWARE additive
checksum of the node file (`tests/filesum.py`) which ranks of a sharded build can each compute over their
own [node_lo, node_hi) range and add up -- so the multi-GPU path is checked against the reference's bytes
without ever assembling the 9.5 GB file in one place.
Existing entries of other configs are kept (the file is updated, not rewritten)."""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

import numpy as np  # noqa: E402
from ooc_svo_builder_b200 import meshgen  # noqa: E402
from oracle import oracle as O  # noqa: E402
from filesum import filesum_file  # noqa: E402

CFG = {"c1": ("c1_icosphere_256", 256), "c2": ("c2_displaced_sphere_1024", 1024), "c3": ("c3_terrain_2048_payload", 2048),
       "c4": ("c4_sphere_4096", 4096), "c5": ("c5_shell_8192", 8192)}
OUT = os.path.join(HERE, "golden_scale.json")


def sha_file(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        while True:
            b = f.read(64 << 20)
            if not b:
                break
            h.update(b)
    return h.hexdigest()


def sha_array(a):
    h = hashlib.sha256()
    flat = a.reshape(-1).view(np.uint8)
    step = 256 << 20
    for lo in range(0, flat.size, step):
        h.update(flat[lo:lo + step].tobytes())
    return h.hexdigest()


def main():
    assert O.ref_available(), "build oracle/_ref first: make -C oracle ref"
    out = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for n in sys.argv[1:]:
        cfg, g = CFG[n]
        t = time.time()
        mesh = meshgen.make(cfg)
        payload = mesh.payload
        P = O.lib().svo_oracle_estimate_partitions(g, 2048)
        d = tempfile.mkdtemp(prefix="svo_golden_%s_" % n, dir=os.environ.get("SVO_GOLDEN_TMP", "/tmp"))
        try:
            res = {"config": cfg, "gridsize": g, "n_triangles": mesh.n_triangles, "payload": bool(payload), "n_partitions": int(P),
                   "mesh_sha256": sha_array(mesh.tris), "gen_s": time.time() - t}
            hdr = meshgen.write_tri(os.path.join(d, "m"), mesh)
            del mesh
            exe = O.ref_exe(payload)
            t = time.time()
            p = subprocess.run([exe, "-f", hdr, "-s", str(g)], capture_output=True, text=True)
            res["ref_wall_s"] = time.time() - t
            for line in p.stdout.splitlines():
                if line.startswith("Total amount of voxels:"):
                    res["n_voxels"] = int(line.split(":")[1])
            base = os.path.join(d, "m%d_%d" % (g, P))
            res["header"] = open(base + ".octree").read()
            res["nodes_sha256"] = sha_file(base + ".octreenodes")
            res["data_sha256"] = sha_file(base + ".octreedata")
            res["n_nodes"] = os.path.getsize(base + ".octreenodes") // 24
            res["n_data"] = os.path.getsize(base + ".octreedata") // 32
            res["nodes_filesum"] = [int(x) for x in filesum_file(base + ".octreenodes")]
            res["data_filesum"] = [int(x) for x in filesum_file(base + ".octreedata")]
        finally:
            shutil.rmtree(d, ignore_errors=True)
        out[n] = res
        with open(OUT, "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)
        print(n, json.dumps({k: v for k, v in res.items() if k != "header"}), flush=True)


if __name__ == "__main__":
    main()
