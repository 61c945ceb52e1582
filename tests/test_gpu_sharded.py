"""GPU: the sharded (multi-GPU) build path, all ranks as contexts of one process on cuda:0 -- the C ABI
does not care whether ranks share a device -- with the table exchange done as a tensor sum. The
assembled per-rank file ranges must equal the oracle's files byte for byte."""
import numpy as np
import pytest

from ooc_svo_builder_b200 import meshgen as mg
from ooc_svo_builder_b200 import sharded

pytestmark = pytest.mark.gpu


def _check(oracle, mesh, g, world, limit=2048, color="model"):
    res = sharded.run_single_process(mesh.tris, mesh.length, g, world, memory_limit_mb=limit, color=color)
    hdr, nodes, data = sharded.assemble(res, g)
    want = oracle.build(mesh.tris, mesh.length, g, memory_limit_mb=limit, color=color)
    assert hdr == want.header
    gn = np.frombuffer(nodes.tobytes(), dtype=np.uint64).reshape(-1, 3)
    wn = np.frombuffer(want.nodes, dtype=np.uint64).reshape(-1, 3)
    bad = np.flatnonzero((gn != wn).any(axis=1))
    assert bad.size == 0, "first differing node %d of %d: got %s want %s; ranges %s" % (
        bad[0], len(wn), gn[bad[0]], wn[bad[0]], [r.node_range for r in res])
    assert data.tobytes() == want.data
    assert all(r.n_voxels == want.n_voxels for r in res)
    return res


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_single_partition_grid(oracle, world):
    # P == 1: the grid is cut into octants for building only; voxelizer semantics stay those of one partition
    _check(oracle, mg.icosphere(5), 128, world)
    _check(oracle, mg.random_soup(1200, seed=3, large_frac=0.03), 256, world)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_eight_partitions(oracle, world):
    _check(oracle, mg.icosphere(6), 256, world, limit=3)
    _check(oracle, mg.random_soup(1500, seed=5, large_frac=0.03), 256, world, limit=2)


@pytest.mark.parametrize("world", [2, 8])
def test_sharded_64_partitions_odd_depth(oracle, world):
    _check(oracle, mg.random_soup(1500, seed=11), 512, world, limit=2)


def test_sharded_empty_ranks(oracle):
    # geometry confined to one octant: most ranks own nothing
    m = mg.icosphere(4)
    t = (m.tris * np.float32(0.45)).astype(np.float32)
    _check(oracle, mg.Mesh(t, 2.0), 256, 8, limit=3)
    _check(oracle, mg.empty_mesh(), 256, 4, limit=3)
    _check(oracle, mg.single_triangle_on_partition_plane(), 256, 8, limit=3)


@pytest.mark.parametrize("world", [2, 8])
def test_sharded_payload(oracle, world):
    m = mg.icosphere(5)
    _check(oracle, mg.Mesh(mg.with_payload(m.tris), m.length), 256, world, limit=3)
    _check(oracle, mg.terrain(100, seed=2), 128, world, color="linear")


def test_sharded_levels_is_rejected():
    from ooc_svo_builder_b200 import SvoBuilder, SvoError
    sb = SvoBuilder(0)
    try:
        sb.shard_configure(0, 2)
        m = mg.icosphere(2)
        sb.set_triangles(m.tris)
        with pytest.raises(SvoError):
            sb.partition(SvoBuilder.make_params(m.length, 64, False, levels=True))
    finally:
        sb.close()
