"""GPU parity tests: the CUDA path through the C ABI vs the CPU oracle, byte for byte.

Bar (BASELINE.json north_star): binary mode bit-exact (same voxels, same node
bytes, same file bytes); payload mode bit-exact index structure and -- because
the kernels replicate the reference's float operation order without FMA --
bit-exact floats too (tolerance 0 ulp, asserted on the raw bytes).
"""
import numpy as np
import pytest

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


def test_c1_icosphere_256(builder, oracle):
    got = _check(builder, oracle, mg.icosphere(6), 256)
    assert got.n_voxels == 308581 and got.n_nodes == 411166      # SURVEY.md §8c known answer


def test_icosphere_8_partitions(builder, oracle):
    _check(builder, oracle, mg.icosphere(6), 256, memory_limit_mb=3)


@pytest.mark.parametrize("g", [2, 4, 8, 16, 32, 64, 128, 512])
def test_gridsizes(builder, oracle, g):
    _check(builder, oracle, mg.icosphere(3), g)


def test_f5_partition_plane(builder, oracle):
    m = mg.single_triangle_on_partition_plane()
    assert _check(builder, oracle, m, 256).n_voxels == 378
    assert _check(builder, oracle, m, 256, memory_limit_mb=3).n_voxels == 756


@pytest.mark.parametrize("limit", [2048, 2, 1])
def test_axis_aligned_box(builder, oracle, limit):
    _check(builder, oracle, mg.axis_aligned_box(), 128, memory_limit_mb=max(limit, 2) if limit != 1 else 2)
    _check(builder, oracle, mg.axis_aligned_box(2.0, 0.5, 1.5), 256, memory_limit_mb=3)


def test_degenerate(builder, oracle):
    _check(builder, oracle, mg.degenerate_mix(), 64)
    _check(builder, oracle, mg.degenerate_mix(), 256, memory_limit_mb=2)


def test_empty(builder, oracle):
    for g in (2, 4, 64, 128):
        got = _check(builder, oracle, mg.empty_mesh(), g)
        assert got.n_nodes == 1
    _check(builder, oracle, mg.empty_mesh(payload=True), 64)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_soup_all_classes(builder, oracle, seed):
    # small + medium + large bounding boxes, non power-of-two unit length
    got = _check(builder, oracle, mg.random_soup(2500, seed=seed, large_frac=0.03), 256, memory_limit_mb=2)
    assert got.stats["n_medium"] > 0 and got.stats["n_large"] > 0


def test_random_soup_64_partitions(builder, oracle):
    _check(builder, oracle, mg.random_soup(1500, seed=11), 512, memory_limit_mb=2)


def test_displaced_sphere(builder, oracle):
    _check(builder, oracle, mg.displaced_sphere(200, 200, seed=1), 512)


def test_payload_icosphere(builder, oracle):
    m = mg.icosphere(5)
    _check(builder, oracle, mg.Mesh(mg.with_payload(m.tris), m.length), 128)


def test_payload_partitions(builder, oracle):
    m = mg.icosphere(5)
    _check(builder, oracle, mg.Mesh(mg.with_payload(m.tris), m.length), 256, memory_limit_mb=3)


@pytest.mark.parametrize("color", ["fixed", "linear", "normal"])
def test_payload_color_modes(builder, oracle, color):
    _check(builder, oracle, mg.terrain(60, seed=2), 128, color=color)


def test_payload_soup(builder, oracle):
    _check(builder, oracle, mg.random_soup(1500, seed=4, payload=True), 128, memory_limit_mb=2)
    _check(builder, oracle, mg.terrain(120, seed=2), 256, memory_limit_mb=3)


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
