"""GPU parity tests: the CUDA path through the C ABI vs the CPU oracle, byte for byte.

Bar (BASELINE.json north_star): binary mode bit-exact (same voxels, same node
bytes, same file bytes); payload mode bit-exact index structure and -- because
the kernels replicate the reference's float operation order without FMA --
bit-exact floats too (tolerance 0 ulp, asserted on the raw bytes).
"""
import hashlib
import json
import os

import numpy as np
import pytest

from cases import CASES
from ooc_svo_builder_b200 import meshgen as mg

pytestmark = pytest.mark.gpu


def _check(builder, oracle, mesh, gridsize, memory_limit_mb=2048, color="model", levels=False):
    got = builder.run(mesh.tris, mesh.length, gridsize, memory_limit_mb=memory_limit_mb, color=color, levels=levels)
    want = oracle.build(mesh.tris, mesh.length, gridsize, memory_limit_mb=memory_limit_mb, color=color, levels=levels)
    assert got.n_partitions == want.n_partitions
    assert got.n_voxels == want.n_voxels
    assert got.header == want.header
    gn = np.frombuffer(got.nodes.tobytes(), dtype=np.uint64).reshape(-1, 3)
    wn = np.frombuffer(want.nodes, dtype=np.uint64).reshape(-1, 3)
    assert gn.shape == wn.shape
    bad = np.flatnonzero((gn != wn).any(axis=1))
    assert bad.size == 0, "first differing node %d: got %s want %s" % (bad[0], gn[bad[0]], wn[bad[0]])
    gd = np.frombuffer(got.data.tobytes(), dtype=np.uint32).reshape(-1, 8)
    wd = np.frombuffer(want.data, dtype=np.uint32).reshape(-1, 8)
    assert gd.shape == wd.shape
    bad = np.flatnonzero((gd != wd).any(axis=1))
    assert bad.size == 0, "first differing data record %d: got %s want %s (%d differ)" % (bad[0], gd[bad[0]], wd[bad[0]], bad.size)
    return got


GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.json")))


@pytest.mark.parametrize("name,factory,g,kw", CASES, ids=[c[0] for c in CASES])
def test_case_vs_oracle_and_golden(builder, oracle, name, factory, g, kw):
    mesh = factory()
    got = _check(builder, oracle, mesh, g, memory_limit_mb=kw.get("memory_limit_mb", 2048),
                 color=kw.get("color", "model"), levels=kw.get("levels", False))
    gold = GOLDEN[name]
    # The generators are deterministic across this image's hosts (tools/meshcheck.py, run on the GPU box): a mesh that
    # differs from the one the golden digests were made from is an error, not a reason to skip the comparison.
    assert hashlib.sha256(mesh.tris.tobytes()).hexdigest() == gold["mesh_sha256"], "the generated mesh differs from the golden input"
    # digests of the UNMODIFIED reference's output files (tests/golden/make_golden.py)
    assert got.header.decode() == gold["header"]
    assert hashlib.sha256(got.nodes.tobytes()).hexdigest() == gold["nodes_sha256"]
    assert hashlib.sha256(got.data.tobytes()).hexdigest() == gold["data_sha256"]
    assert got.n_voxels == gold["n_voxels"]


def test_known_answers(builder, oracle):
    got = _check(builder, oracle, mg.icosphere(6), 256)
    assert got.n_voxels == 308581 and got.n_nodes == 411166
    m = mg.single_triangle_on_partition_plane()
    assert _check(builder, oracle, m, 256).n_voxels == 378
    assert _check(builder, oracle, m, 256, memory_limit_mb=3).n_voxels == 756


@pytest.mark.parametrize("g", [2, 4, 8, 16, 64, 128])
def test_gridsizes_payload_levels(builder, oracle, g):
    m = mg.icosphere(2)
    _check(builder, oracle, mg.Mesh(mg.with_payload(m.tris), m.length), g, levels=True)
    _check(builder, oracle, m, g, levels=True)
    _check(builder, oracle, mg.empty_mesh(payload=True), g, levels=True)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_soup_all_classes(builder, oracle, seed):
    # small + medium + large bounding boxes, non power-of-two unit length
    got = _check(builder, oracle, mg.random_soup(2500, seed=seed, large_frac=0.03), 256, memory_limit_mb=2)
    assert got.stats["n_medium"] > 0 and got.stats["n_large"] > 0


def test_huge_triangles(builder, oracle):
    # two triangles spanning the whole cube diagonal + a quad on a face: exercises the large class and the exact box pruning
    L = 1.9
    t = np.array([[0, 0, 0, L, L, 0.3, 0.2, L, L], [L, 0, 0.1, 0, L, 0.7, L, L, L],
                  [0, 0, L, L, 0, L, L, L, L], [0.05, 0, 0.05, L, 0.02, L, 0.01, L, 0.5]], dtype=np.float32)
    _check(builder, oracle, mg.Mesh(t, L), 256)
    _check(builder, oracle, mg.Mesh(t, L), 512, memory_limit_mb=2)


def test_c2_full_size_bit_exact(builder, oracle):
    """BASELINE.json configs[1] at full size: svo_builder_binary -s 1024 on the 2 M-triangle displaced
    sphere, bit-exact .octree / .octreenodes / .octreedata against the CPU oracle."""
    mesh = mg.displaced_sphere(1000, 1000, seed=1)
    got = builder.run(mesh.tris, mesh.length, 1024)
    want = oracle.build(mesh.tris, mesh.length, 1024)
    assert got.header == want.header and got.n_voxels == want.n_voxels
    assert hashlib.sha256(got.nodes.tobytes()).digest() == hashlib.sha256(want.nodes).digest()
    assert got.data.tobytes() == want.data
    # size-independent invariants
    n = np.frombuffer(got.nodes.tobytes(), dtype=np.uint64).reshape(-1, 3)
    leaf = n[:, 2] == np.uint64(0xFFFFFFFFFFFFFFFF)
    assert leaf.sum() == got.n_voxels and (n[leaf, 0] == 1).all() and (n[leaf, 1] == 0).all()
    codes = builder.voxel_codes()
    assert codes.size == got.n_voxels and (np.diff(codes.astype(np.int64)) > 0).all()


@pytest.mark.parametrize("name", ["c1_icosphere_256_p8", "soup0_256_p8", "soup11_512_p64", "f5_plane_p8", "payload_terrain_256_p8", "box_256_p8"])
def test_list_based_partitioner_mode(builder, oracle, name, monkeypatch):
    """SVO_PARTITION_LISTS=1: per-partition index lists (count / scan / fill with warp-aggregated atomics)
    feeding the voxelizer, instead of the default inline enumeration. Same bytes either way."""
    monkeypatch.setenv("SVO_PARTITION_LISTS", "1")
    _, factory, g, kw = next(c for c in CASES if c[0] == name)
    _check(builder, oracle, factory(), g, memory_limit_mb=kw.get("memory_limit_mb", 2048), color=kw.get("color", "model"))


def test_partition_counts_match_oracle(builder, oracle):
    # the values the reference writes to the .trip header (trip_tools.h:118-120)
    m = mg.random_soup(3000, seed=12, large_frac=0.03)
    for g, lim in ((256, 2), (512, 2)):
        prm = builder.make_params(m.length, g, False, memory_limit_mb=lim)
        builder.set_triangles(m.tris)
        got = builder.partition(prm)
        want = oracle.partition_counts(m.tris, m.length, g, len(got))
        assert np.array_equal(got, want)
        builder.voxelize(); builder.build()
        assert builder.stats()["n_pairs"] == int(want.sum())


def test_voxel_codes_match_oracle(builder, oracle):
    m = mg.random_soup(800, seed=9)
    builder.run(m.tris, m.length, 128)
    codes = builder.voxel_codes()
    want = oracle.voxelize(m.tris, m.length, 128)
    assert np.array_equal(codes, want)


def test_rerun_is_idempotent(builder, oracle):
    # the sparse clear must leave a clean pyramid: same context, different meshes, back to back
    a = mg.random_soup(500, seed=21)
    b = mg.icosphere(4)
    for m in (a, b, a, b):
        _check(builder, oracle, m, 128)


def test_svo_run_host_buffers(builder, oracle):
    m = mg.icosphere(5)
    want = oracle.build(m.tris, m.length, 128)
    prm = builder.make_params(m.length, 128, False)
    nodes = np.zeros(want.n_nodes * 24, dtype=np.uint8)
    data = np.zeros(want.n_data * 32, dtype=np.uint8)
    st = builder.run_host(prm, m.tris, nodes, data)
    assert st["n_nodes"] == want.n_nodes and nodes.tobytes() == want.nodes and data.tobytes() == want.data


def test_speculative_emission_regrows_the_node_buffer(oracle):
    # a context that built a SMALL tree first emits the next, larger tree speculatively into the old buffer (guarded),
    # notices the overflow with the final sync and repeats the emission into a bigger buffer: files still byte-exact
    from ooc_svo_builder_b200 import SvoBuilder
    sb = SvoBuilder(0)
    try:
        for mesh, g in ((mg.icosphere(2), 32), (mg.icosphere(5), 256), (mg.icosphere(3), 64), (mg.random_soup(2000, seed=12), 256)):
            got = sb.run(mesh.tris, mesh.length, g)
            want = oracle.build(mesh.tris, mesh.length, g)
            assert got.header == want.header
            assert got.nodes.tobytes() == want.nodes and got.data.tobytes() == want.data
    finally:
        sb.close()


@pytest.mark.parametrize("chunk", [1, 1000, 4096, 10 ** 6])
def test_streamed_triangle_upload(oracle, builder, chunk):
    # svo_triangles_begin / _append (what the CLI does while it reads the file) == svo_set_triangles
    from ooc_svo_builder_b200 import SvoBuilder
    m = mg.random_soup(5000, seed=21, large_frac=0.01) if chunk > 1 else mg.icosphere(1)
    builder.set_triangles_streamed(m.tris, chunk)
    prm = SvoBuilder.make_params(m.length, 128, False)
    builder.partition(prm, want_counts=False); builder.voxelize()
    nv, nn, nd = builder.build()
    want = oracle.build(m.tris, m.length, 128)
    assert builder.fetch_nodes(0, nn).tobytes() == want.nodes and nv == want.n_voxels


def test_steady_state_build_needs_no_readback_and_is_exact(oracle):
    """Second and later builds of a context run speculatively (capacities of the previous build, counts on the device,
    no host read-back before the final one): same bytes; a different mesh in between forces the sized path again."""
    from ooc_svo_builder_b200 import SvoBuilder
    sb = SvoBuilder(0)
    try:
        a, b = mg.displaced_sphere(150, 150, seed=3), mg.random_soup(3000, seed=17, large_frac=0.02)
        wa, wb = oracle.build(a.tris, a.length, 256), oracle.build(b.tris, b.length, 512, memory_limit_mb=2)
        for mesh, g, lim, want in ((a, 256, 2048, wa), (a, 256, 2048, wa), (a, 256, 2048, wa), (b, 512, 2, wb), (b, 512, 2, wb), (a, 256, 2048, wa)):
            got = sb.run(mesh.tris, mesh.length, g, memory_limit_mb=lim)
            assert got.header == want.header and got.nodes.tobytes() == want.nodes and got.data.tobytes() == want.data
    finally:
        sb.close()


@pytest.mark.parametrize("name", ["c1_icosphere_256", "soup0_256_p8", "soup11_512_p64", "payload_ico5_256_p8", "icosphere3_g32", "sphere200_512"])
def test_classic_build_path(builder, oracle, name, monkeypatch):
    """SVO_BUILD_PATH=classic: the host-driven build (the only one for -levels and grids below 16^3) on cases the
    device-driven build normally takes. Same bytes either way."""
    monkeypatch.setenv("SVO_BUILD_PATH", "classic")
    _, factory, g, kw = next(c for c in CASES if c[0] == name)
    _check(builder, oracle, factory(), g, memory_limit_mb=kw.get("memory_limit_mb", 2048), color=kw.get("color", "model"))


def test_second_build_without_voxelize_is_an_error(builder):
    from ooc_svo_builder_b200 import SvoError
    m = mg.icosphere(3)
    builder.run(m.tris, m.length, 64)
    with pytest.raises(SvoError):
        builder.build()
