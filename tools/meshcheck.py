import sys, os, json, hashlib, subprocess
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from cases import CASES
from ooc_svo_builder_b200 import meshgen
G = json.load(open("tests/golden/golden.json"))
print(subprocess.run("lscpu | grep -E 'Model name|Flags' | cut -c1-400", shell=True, capture_output=True, text=True).stdout)
import numpy; print(numpy.__version__)
bad = 0
for name, f, g, kw in CASES:
    m = f()
    ok = hashlib.sha256(m.tris.tobytes()).hexdigest() == G[name]["mesh_sha256"]
    bad += (not ok)
    print(name, "same" if ok else "DIFFERENT")
print("different:", bad)
for c in ("c2_displaced_sphere_1024", "c3_terrain_2048_payload", "c4_sphere_4096"):
    m = meshgen.make(c)
    print(c, hashlib.sha256(m.tris.tobytes()).hexdigest())
