#!/usr/bin/env python
"""Generates tests/golden/golden.json (+ a few complete output files) by running the
UNMODIFIED reference (oracle/_ref, built from /root/reference by oracle/Makefile) on
the shared parity cases.  Run in the dev container:  python tests/golden/make_golden.py

The reference itself ships no tests or golden vectors; these digests
pin the oracle restatement and the CUDA path to the reference's actual output."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from oracle import oracle as O  # noqa: E402
from cases import CASES, FULL_FILE_CASES  # noqa: E402


def sha(b):
    return hashlib.sha256(b).hexdigest()


def main():
    assert O.ref_available(), "build oracle/_ref first: make -C oracle ref"
    out = {}
    for name, factory, g, kw in CASES:
        mesh = factory()
        r = O.ref_build(mesh, g, memory_limit_mb=kw.get("memory_limit_mb"), levels=kw.get("levels", False), color=kw.get("color"))
        out[name] = {
            "gridsize": g, "kwargs": kw, "n_triangles": mesh.n_triangles,
            "mesh_sha256": sha(mesh.tris.tobytes()),
            "n_partitions": r.n_partitions, "n_voxels": r.n_voxels, "n_nodes": r.n_nodes, "n_data": r.n_data,
            "header": r.header.decode(), "nodes_sha256": sha(r.nodes), "data_sha256": sha(r.data),
        }
        if name in FULL_FILE_CASES:
            with open(os.path.join(HERE, name + ".octreenodes"), "wb") as f:
                f.write(r.nodes)
            with open(os.path.join(HERE, name + ".octreedata"), "wb") as f:
                f.write(r.data)
        print(name, out[name]["n_voxels"], out[name]["n_nodes"], out[name]["n_data"], flush=True)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
