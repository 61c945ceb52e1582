#!/usr/bin/env python
"""End-to-end drop-in check at BASELINE.json scale: writes the config's .tri/.tridata, runs the UNMODIFIED
reference CLI (oracle/_ref, CPU, 1 thread) and our CLI (ooc_svo_builder_b200/bin, B200) on the same files and
compares the three output files byte for byte. Prints wall-clock times of both processes (file IO included).

    python tests/scale_parity.py c3 c4
    python tests/scale_parity.py --golden [--gpus N] c4 c5     our CLI only, against the digests of the reference's files in
                                                               tests/golden/golden_scale.json (the reference run takes minutes)
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ooc_svo_builder_b200 import meshgen, estimate_partitions  # noqa: E402

CFG = {"c1": ("c1_icosphere_256", 256), "c2": ("c2_displaced_sphere_1024", 1024), "c3": ("c3_terrain_2048_payload", 2048),
       "c4": ("c4_sphere_4096", 4096), "c5": ("c5_shell_8192", 8192)}


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        while True:
            b = f.read(64 << 20)
            if not b:
                break
            h.update(b)
    return h.hexdigest()


def main():
    out = {}
    args = sys.argv[1:]
    golden = None
    gpus = 1
    if "--golden" in args:
        args.remove("--golden")
        golden = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_scale.json")))
    if "--gpus" in args:
        i = args.index("--gpus")
        gpus = int(args[i + 1])
        del args[i:i + 2]
    for n in args:
        cfg, g = CFG[n]
        mesh = meshgen.make(cfg)
        payload = mesh.payload
        P = estimate_partitions(g, 2048)
        res = {"n_triangles": mesh.n_triangles, "gridsize": g, "partitions": P}
        dirs = {}
        for who in (("ours",) if golden else ("ref", "ours")):
            d = tempfile.mkdtemp(prefix="svo_%s_%s_" % (n, who), dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
            dirs[who] = d
            hdr = meshgen.write_tri(os.path.join(d, "m"), mesh)
            exe = os.path.join(ROOT, "oracle", "_ref") if who == "ref" else os.path.join(ROOT, "ooc_svo_builder_b200", "bin")
            exe = os.path.join(exe, "svo_builder" if payload else "svo_builder_binary")
            t = time.perf_counter()
            cmd = [exe, "-f", hdr, "-s", str(g)] + (["-gpus", str(gpus)] if who == "ours" and gpus > 1 else [])
            p = subprocess.run(cmd, capture_output=True, text=True)
            res[who + "_wall_s"] = time.perf_counter() - t
            for line in p.stdout.splitlines():
                if line.startswith("Total amount of voxels:"):
                    res[who + "_voxels"] = int(line.split(":")[1])
            base = os.path.join(d, "m%d_%d" % (g, P))
            res[who + "_sha"] = {ext: sha(base + ext) for ext in (".octree", ".octreenodes", ".octreedata")}
            res[who + "_bytes"] = sum(os.path.getsize(base + ext) for ext in (".octree", ".octreenodes", ".octreedata"))
            if who == "ours":
                res["ours_tail"] = p.stdout[-600:]
        if golden:
            gd = golden[n]
            res["gpus"] = gpus
            res["ref_wall_s"] = gd["ref_wall_s"]                      # measured in the dev container when the golden was made
            res["identical"] = (res["ours_sha"][".octreenodes"] == gd["nodes_sha256"] and res["ours_sha"][".octreedata"] == gd["data_sha256"]
                                and open(os.path.join(dirs["ours"], "m%d_%d.octree" % (g, P))).read() == gd["header"])
            res["against"] = "tests/golden/golden_scale.json (sha256 of the unmodified reference's files)"
        else:
            res["identical"] = res["ref_sha"] == res["ours_sha"]
        res["speedup_wall"] = res["ref_wall_s"] / res["ours_wall_s"]
        for d in dirs.values():
            shutil.rmtree(d, ignore_errors=True)
        out[n] = res
        print(n, json.dumps({k: v for k, v in res.items() if not k.endswith("_sha") and k != "ours_tail"}), flush=True)
        del mesh
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "scale_parity.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
