"""GPU: randomized parity against the oracle -- random soups with random grid sizes, memory limits (partition
counts), bbox lengths (unit lengths that are not powers of two), payload / colour modes / -levels, both
partitioner modes and the sharded path. Seeds are fixed: failures are reproducible."""
import numpy as np
import pytest

from ooc_svo_builder_b200 import meshgen as mg
from ooc_svo_builder_b200 import sharded

pytestmark = pytest.mark.gpu


def _random_case(seed):
    rng = np.random.default_rng(1000 + seed)
    g = int(rng.choice([16, 32, 64, 128, 256]))
    length = float(rng.choice([1.0, 1.9, 2.0, 3.3333333, 0.7, 117.25]))
    n = int(rng.integers(1, 1500))
    payload = bool(rng.integers(0, 2))
    m = mg.random_soup(n, seed=seed, length=length, small=float(rng.choice([0.005, 0.02, 0.08])),
                       large_frac=float(rng.choice([0.0, 0.01, 0.05])), payload=payload)
    # a few vertices exactly on voxel / partition planes and on the cube boundary
    t = m.tris.copy()
    k = min(len(t), 8)
    u = np.float32(length) / np.float32(g)
    t[:k, 0] = np.round(t[:k, 0] / u) * u
    t[:k, 4] = np.float32(length) / 2
    t[k:2 * k, 8] = np.float32(length)
    t[:, :9] = np.clip(t[:, :9], 0, np.float32(length))
    limit = int(rng.choice([2048, max(2, g ** 3 // (1 << 20) // 8), max(2, g ** 3 // (1 << 20) // 64)]))
    color = str(rng.choice(["model", "fixed", "linear", "normal"])) if payload else "model"
    levels = bool(rng.integers(0, 4) == 0)
    return mg.Mesh(t, length), g, limit, color, levels


def _compare(got_header, got_nodes, got_data, want):
    assert got_header == want.header
    assert got_nodes == want.nodes
    assert got_data == want.data


@pytest.mark.parametrize("seed", range(24))
def test_random_parity(builder, oracle, seed, monkeypatch):
    mesh, g, limit, color, levels = _random_case(seed)
    if seed % 3 == 0:
        monkeypatch.setenv("SVO_PARTITION_LISTS", "1")
    got = builder.run(mesh.tris, mesh.length, g, memory_limit_mb=limit, color=color, levels=levels)
    want = oracle.build(mesh.tris, mesh.length, g, memory_limit_mb=limit, color=color, levels=levels)
    assert got.n_voxels == want.n_voxels, (g, limit, mesh.n_triangles)
    _compare(got.header, got.nodes.tobytes(), got.data.tobytes(), want)


@pytest.mark.parametrize("seed", range(8))
def test_random_parity_sharded(oracle, seed):
    mesh, g, limit, color, _ = _random_case(100 + seed)
    g = max(g, 64)
    world = int(np.random.default_rng(seed).choice([2, 4, 8]))
    res = sharded.run_single_process(mesh.tris, mesh.length, g, world, memory_limit_mb=limit, color=color)
    hdr, nodes, data = sharded.assemble(res, g)
    want = oracle.build(mesh.tris, mesh.length, g, memory_limit_mb=limit, color=color)
    _compare(hdr, nodes.tobytes(), data.tobytes(), want)


def test_tiny_coordinates_and_huge_length(builder, oracle):
    # unit lengths far from 1: 2^-20 and 1e6
    for length in (2.0 ** -14, 1.0e6):
        m = mg.random_soup(400, seed=77, length=length)
        got = builder.run(m.tris, m.length, 64)
        want = oracle.build(m.tris, m.length, 64)
        _compare(got.header, got.nodes.tobytes(), got.data.tobytes(), want)
