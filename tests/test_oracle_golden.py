"""CPU: the oracle restatement (oracle/svo_oracle.c) against the golden digests that
tests/golden/make_golden.py took from the UNMODIFIED reference, and -- when the
reference binaries are present (dev container, or shipped prebuilt in oracle/_ref) --
against the reference itself."""
import hashlib
import json
import os

import numpy as np
import pytest

from cases import CASES, FULL_FILE_CASES

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "golden.json")))


def sha(b):
    return hashlib.sha256(b).hexdigest()


@pytest.mark.parametrize("name,factory,g,kw", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_reference_digest(oracle, name, factory, g, kw):
    gold = GOLDEN[name]
    mesh = factory()
    # the generators are deterministic across this image's hosts (tools/meshcheck.py): a different mesh is an error
    assert sha(mesh.tris.tobytes()) == gold["mesh_sha256"], "the generated mesh differs from the golden input"
    r = oracle.build(mesh.tris, mesh.length, g, memory_limit_mb=kw.get("memory_limit_mb", 2048),
                     levels=kw.get("levels", False), color=kw.get("color", "model"))
    assert r.n_partitions == gold["n_partitions"]
    assert r.n_voxels == gold["n_voxels"]
    assert r.header.decode() == gold["header"]
    assert (r.n_nodes, r.n_data) == (gold["n_nodes"], gold["n_data"])
    assert sha(r.nodes) == gold["nodes_sha256"]
    assert sha(r.data) == gold["data_sha256"]
    if name in FULL_FILE_CASES:
        assert r.nodes == open(os.path.join(HERE, "golden", name + ".octreenodes"), "rb").read()
        assert r.data == open(os.path.join(HERE, "golden", name + ".octreedata"), "rb").read()


def test_known_answers(oracle):
    # SURVEY.md §8c / F5 known answers
    assert GOLDEN["c1_icosphere_256"]["n_voxels"] == 308581 and GOLDEN["c1_icosphere_256"]["n_nodes"] == 411166
    assert GOLDEN["f5_plane_p1"]["n_voxels"] == 378 and GOLDEN["f5_plane_p8"]["n_voxels"] == 756
    assert GOLDEN["c1_icosphere_256_p8"]["nodes_sha256"] == GOLDEN["c1_icosphere_256"]["nodes_sha256"]


@pytest.mark.parametrize("name", ["f5_plane_p8", "degenerate_256_p8", "payload_ico4_levels", "payload_terrain_linear"])
def test_oracle_vs_live_reference(oracle, name):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    _, factory, g, kw = next(c for c in CASES if c[0] == name)
    mesh = factory()
    ref = oracle.ref_build(mesh, g, memory_limit_mb=kw.get("memory_limit_mb"), levels=kw.get("levels", False), color=kw.get("color"))
    got = oracle.build(mesh.tris, mesh.length, g, memory_limit_mb=kw.get("memory_limit_mb", 2048),
                       levels=kw.get("levels", False), color=kw.get("color", "model"))
    assert (ref.header, ref.nodes, ref.data, ref.n_voxels) == (got.header, got.nodes, got.data, got.n_voxels)


def test_estimate_partitions(oracle):
    # partitioner.cpp:12-28: P is the smallest power of 8 with g^3/2^20/P <= limit (integer MB)
    assert oracle.estimate_partitions(256, 2048) == 1
    assert oracle.estimate_partitions(1024, 2048) == 1
    assert oracle.estimate_partitions(2048, 2048) == 8
    assert oracle.estimate_partitions(4096, 2048) == 64
    assert oracle.estimate_partitions(8192, 2048) == 512
    assert oracle.estimate_partitions(256, 3) == 8
    assert oracle.estimate_partitions(512, 2) == 64


def test_partition_counts_sum(oracle):
    from ooc_svo_builder_b200 import meshgen as mg
    m = mg.icosphere(4)
    c = oracle.partition_counts(m.tris, m.length, 256, 8)
    assert c.sum() >= m.n_triangles and (c > 0).all()


def test_structural_invariants(oracle):
    # invariants that need no reference
    from ooc_svo_builder_b200 import meshgen as mg
    m = mg.icosphere(4)
    r = oracle.build(m.tris, m.length, 64)
    n = np.frombuffer(r.nodes, dtype=np.uint64).reshape(-1, 3)
    off = np.frombuffer(r.nodes, dtype=np.int8).reshape(-1, 24)[:, 16:]
    leaf = (off == -1).all(axis=1)
    assert (n[leaf, 1] == 0).all() and (n[leaf, 0] == 1).all()         # leaves: base 0, data 1 (binary)
    assert (n[~leaf, 0] == 0).all()                                     # internal: data 0 without -levels
    cnt_root = (off[-1] != -1).sum()
    assert n[-1, 1] == len(n) - 1 - cnt_root                            # root last, its children right before it
    assert leaf.sum() == r.n_voxels
