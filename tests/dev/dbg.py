import sys, subprocess, os, tempfile
sys.path.insert(0,'.')
from ooc_svo_builder_b200 import meshgen as mg, SvoBuilder
from oracle import oracle as O
import numpy as np
d=tempfile.mkdtemp()
m=mg.icosphere(4)
h=mg.write_tri(os.path.join(d,"mesh"),m)
p=subprocess.run(["ooc_svo_builder_b200/bin/svo_builder_binary","-f",h,"-s","128","-v"],capture_output=True,text=True)
print(p.stdout[-1500:], p.stderr[-500:])
sb=SvoBuilder(0)
for mesh,g,kw in ((mg.icosphere(4),128,{}),(mg.random_soup(1500,seed=7),256,{"memory_limit_mb":2}),(mg.icosphere(4),128,{})):
    try:
        got=sb.run(mesh.tris,mesh.length,g,**kw)
        want=O.build(mesh.tris,mesh.length,g,**kw)
        gn=np.frombuffer(got.nodes.tobytes(),dtype=np.uint64).reshape(-1,3); wn=np.frombuffer(want.nodes,dtype=np.uint64).reshape(-1,3)
        print("hdr", got.header==want.header, got.n_voxels, want.n_voxels, gn.shape, wn.shape)
        if gn.shape==wn.shape:
            bad=np.flatnonzero((gn!=wn).any(axis=1)); print("bad", bad.size, bad[:10])
            if bad.size: print(gn[bad[:5]], wn[bad[:5]])
        print(got.stats)
    except Exception as e:
        print("ERR", e)
