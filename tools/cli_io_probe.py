#!/usr/bin/env python
"""Where does the CLI's wall clock go at 4096^3 (1.8 GB in, 2.26 GB out)? Writes the C4 mesh to /dev/shm once, then runs
ooc_svo_builder_b200/bin/svo_builder_binary a few times (IO thread counts, 1 and N GPUs) and prints its own timing block
next to the process wall clock; plus the raw speed of reading the same file with plain read() calls.
    python tools/cli_io_probe.py [c4] [gpus]"""
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ooc_svo_builder_b200 import meshgen  # noqa: E402

CFG = {"c2": ("c2_displaced_sphere_1024", 1024), "c4": ("c4_sphere_4096", 4096), "c5": ("c5_shell_8192", 8192)}
name = sys.argv[1] if len(sys.argv) > 1 else "c4"
gpus = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg, g = CFG[name]
d = tempfile.mkdtemp(prefix="svo_io_", dir="/dev/shm")
mesh = meshgen.make(cfg)
hdr = meshgen.write_tri(os.path.join(d, "m"), mesh)
del mesh
data = hdr + "data"
size = os.path.getsize(data)
buf = bytearray(16 << 20)
t = time.perf_counter()
with open(data, "rb", buffering=0) as f:
    while f.readinto(buf):
        pass
print("plain read() of %s: %.2f GB in %.3f s = %.2f GB/s" % (data, size / 1e9, time.perf_counter() - t, size / 1e9 / (time.perf_counter() - t)), flush=True)
exe = os.path.join(ROOT, "ooc_svo_builder_b200", "bin", "svo_builder_binary")
runs = [({}, 1), ({}, 1), ({"SVO_IO_THREADS": "1"}, 1), ({"SVO_IO_THREADS": "16"}, 1)]
if gpus > 1:
    runs += [({}, gpus), ({}, gpus)]
for env, ng in runs:
    e = dict(os.environ)
    e.update(env)
    e["SVO_TIMELINE"] = "1"
    cmd = [exe, "-f", hdr, "-s", str(g)] + (["-gpus", str(ng)] if ng > 1 else [])
    t = time.perf_counter()
    p = subprocess.run(cmd, capture_output=True, text=True, env=e)
    wall = time.perf_counter() - t
    keep = [l.strip() for l in p.stdout.splitlines() if re.search(r"MAIN time|IO IN|IO OUT|upload time|Total time|algorithm time", l)]
    print("gpus %d env %s: wall %.3f s | %s" % (ng, env, wall, " | ".join(keep)), flush=True)
    tl = [l for l in p.stderr.splitlines() if "timeline" in l.lower()][:3]
    for l in tl:
        print("   ", l[:300])
    for f in os.listdir(d):
        if f.endswith((".octree", ".octreenodes", ".octreedata")):
            os.remove(os.path.join(d, f))
import shutil
shutil.rmtree(d, ignore_errors=True)
