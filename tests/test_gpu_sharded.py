"""GPU: the sharded (multi-GPU) build path, all ranks as contexts of one process on cuda:0 -- the C ABI
does not care whether ranks share a device -- with the table exchange done as a tensor sum. The
assembled per-rank file ranges must equal the oracle's files byte for byte."""
import numpy as np
import pytest

from ooc_svo_builder_b200 import meshgen as mg
from ooc_svo_builder_b200 import sharded

pytestmark = pytest.mark.gpu


def _check(oracle, mesh, g, world, limit=2048, color="model", levels=False, **kw):
    res = sharded.run_single_process(mesh.tris, mesh.length, g, world, memory_limit_mb=limit, color=color, levels=levels, **kw)
    hdr, nodes, data = sharded.assemble(res, g)
    want = oracle.build(mesh.tris, mesh.length, g, memory_limit_mb=limit, color=color, levels=levels)
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


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_levels(oracle, world):
    """-levels across ranks (OctreeBuilder.cpp:82-99): the data caches of the top-of-shard tiles ride in the exchange table,
    the shared upper levels are averaged from it, internal data records interleave with the leaf records in post-order
    across the ranks' ranges of the data file. Payload (float averaging, bit-exact) and binary (index structure)."""
    ico = mg.icosphere(4)
    _check(oracle, mg.Mesh(mg.with_payload(ico.tris), ico.length), 128, world, levels=True)
    _check(oracle, mg.terrain(90, seed=2), 256, world, limit=3, levels=True)                  # 8 partitions, odd depth
    _check(oracle, mg.random_soup(900, seed=6, payload=True), 128, world, levels=True, color="normal")
    _check(oracle, mg.random_soup(400, seed=3), 64, world, levels=True)                       # binary
    _check(oracle, mg.empty_mesh(payload=True), 64, world, levels=True)


def test_sharded_levels_remote_staging(oracle):
    _check(oracle, mg.terrain(90, seed=2), 256, 4, limit=3, levels=True, remote=True)
    _check(oracle, mg.icosphere(4), 128, 2, levels=True, remote=True)


# ---- triangle dispatch over peer memory: every rank starts with a slice of the file only ----------------------

@pytest.mark.parametrize("world", [2, 4, 8])
def test_dispatch_single_partition_grid(oracle, world):
    _check(oracle, mg.icosphere(5), 128, world, dispatch=True)
    _check(oracle, mg.random_soup(1200, seed=3, large_frac=0.03), 256, world, dispatch=True)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_dispatch_partitions(oracle, world):
    _check(oracle, mg.icosphere(6), 256, world, limit=3, dispatch=True)
    _check(oracle, mg.random_soup(1500, seed=5, large_frac=0.03), 256, world, limit=2, dispatch=True)
    _check(oracle, mg.random_soup(1500, seed=11), 512, world, limit=2, dispatch=True)


@pytest.mark.parametrize("world", [2, 8])
def test_dispatch_payload_keeps_file_order(oracle, world):
    # first-triangle-wins (voxelizer.cpp:263) must survive the trip through the inboxes
    m = mg.icosphere(5)
    _check(oracle, mg.Mesh(mg.with_payload(m.tris), m.length), 256, world, limit=3, dispatch=True)
    _check(oracle, mg.terrain(100, seed=2), 128, world, color="linear", dispatch=True)
    _check(oracle, mg.random_soup(1500, seed=4, payload=True), 128, world, limit=2, dispatch=True)


def test_dispatch_ragged_slices(oracle):
    m = mg.random_soup(1000, seed=9, large_frac=0.02)
    T = m.tris.shape[0]
    _check(oracle, m, 256, 4, limit=3, dispatch=True, slices=[(0, T), (T, T), (T, T), (T, T)])          # all on rank 0
    _check(oracle, m, 256, 4, limit=3, dispatch=True, slices=[(0, 0), (0, 1), (1, 130), (130, T)])      # empty / tiny slices
    _check(oracle, mg.empty_mesh(), 256, 4, limit=3, dispatch=True)
    _check(oracle, mg.single_triangle_on_partition_plane(), 256, 8, limit=3, dispatch=True)


def test_dispatch_repeated_epochs(oracle):
    # the same contexts run two jobs back to back: flags are epochs, the inbox is reused
    import torch
    from ooc_svo_builder_b200 import SvoBuilder
    world = 4
    ctxs = [SvoBuilder(0) for _ in range(world)]
    try:
        for r, sb in enumerate(ctxs):
            sb.shard_configure(r, world)
        ptrs = [sb.dispatch_create(4000, 9) for sb in ctxs]
        for sb in ctxs:
            sb.dispatch_attach([p[0] for p in ptrs], [p[1] for p in ptrs])
        for seed, g in ((1, 128), (2, 256)):
            m = mg.random_soup(3000, seed=seed)
            prm = SvoBuilder.make_params(m.length, g, False, 2)
            T = m.tris.shape[0]
            local = [torch.from_numpy(m.tris[lo:hi].copy()).cuda() for lo, hi in (sharded.slice_bounds(T, world, r) for r in range(world))]
            for sb, l in zip(ctxs, local):
                sb.dispatch_count(prm, l)
            for sb in ctxs:
                sb.dispatch_send()
            tables = []
            for sb in ctxs:
                sb.dispatch_finish()
                sb.partition(prm, want_counts=False)
                sb.voxelize()
                t = torch.zeros(sb.shard_table_size(), dtype=torch.int64, device="cuda")
                sb.shard_count(t.data_ptr())
                sb.synchronize()
                tables.append(t)
            merged = torch.stack(tables).sum(dim=0)
            res = []
            for r, sb in enumerate(ctxs):
                nv, nn, nd = sb.shard_emit(merged.data_ptr())
                nlo, nhi, dlo, dhi = sb.shard_ranges()
                res.append(sharded.ShardResult(r, nv, nn, nd, (nlo, nhi), (dlo, dhi), sb.fetch_nodes(nlo, nhi - nlo), sb.fetch_data(dlo, dhi - dlo), sb.stats()))
            hdr, nodes, data = sharded.assemble(res, g)
            want = oracle.build(m.tris, m.length, g, memory_limit_mb=2)
            assert hdr == want.header and nodes.tobytes() == want.nodes and data.tobytes() == want.data
    finally:
        for sb in ctxs:
            sb.close()


def test_dispatch_inbox_overflow_is_reported_by_every_rank():
    import torch
    from ooc_svo_builder_b200 import SvoBuilder, SvoError
    world = 2
    ctxs = [SvoBuilder(0) for _ in range(world)]
    try:
        for r, sb in enumerate(ctxs):
            sb.shard_configure(r, world)
        m = mg.icosphere(4)                      # 5120 triangles, about half per rank
        T = m.tris.shape[0]
        ptrs = [sb.dispatch_create(T // 2, 9) for sb in ctxs]      # too small: boundary triangles go to both ranks
        for sb in ctxs:
            sb.dispatch_attach([p[0] for p in ptrs], [p[1] for p in ptrs])
        prm = SvoBuilder.make_params(m.length, 128, False)
        # file order of an icosphere is not spatial: each half routes ~half of its triangles to either rank
        local = [torch.from_numpy(m.tris[lo:hi].copy()).cuda() for lo, hi in ((0, T // 2), (T // 2, T))]
        for sb, l in zip(ctxs, local):
            sb.dispatch_count(prm, l)
        for sb in ctxs:
            sb.dispatch_send()
        for sb in ctxs:
            with pytest.raises(SvoError) as e:
                sb.dispatch_finish()
            assert e.value.code == 4
    finally:
        for sb in ctxs:
            sb.close()


# ---- remote staging: every rank keeps its slice, voxelizers read blocks from the owners' buffers ------------------

@pytest.mark.parametrize("world", [2, 4, 8])
def test_remote_slices_single_partition_grid(oracle, world):
    _check(oracle, mg.icosphere(5), 128, world, remote=True)
    _check(oracle, mg.random_soup(1200, seed=3, large_frac=0.03), 256, world, remote=True)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_remote_slices_partitions(oracle, world):
    _check(oracle, mg.icosphere(6), 256, world, limit=3, remote=True)
    _check(oracle, mg.random_soup(1500, seed=5, large_frac=0.03), 256, world, limit=2, remote=True)
    _check(oracle, mg.random_soup(1500, seed=11), 512, world, limit=2, remote=True)


@pytest.mark.parametrize("world", [2, 8])
def test_remote_slices_payload(oracle, world):
    m = mg.icosphere(5)
    _check(oracle, mg.Mesh(mg.with_payload(m.tris), m.length), 256, world, limit=3, remote=True)
    _check(oracle, mg.terrain(100, seed=2), 128, world, color="linear", remote=True)
    _check(oracle, mg.random_soup(1500, seed=4, payload=True, large_frac=0.03), 128, world, limit=2, remote=True)


def test_remote_slices_ragged(oracle):
    m = mg.random_soup(1000, seed=9, large_frac=0.02)
    T = m.tris.shape[0]
    _check(oracle, m, 256, 4, limit=3, remote=True, slices=[(0, T), (T, T), (T, T), (T, T)])
    _check(oracle, m, 256, 4, limit=3, remote=True, slices=[(0, 0), (0, 1), (1, 130), (130, T)])
    _check(oracle, mg.empty_mesh(), 256, 4, limit=3, remote=True)
    _check(oracle, mg.single_triangle_on_partition_plane(), 256, 8, limit=3, remote=True)


def test_remote_slices_repeated_jobs(oracle):
    # same contexts, three jobs back to back with new slice contents: the fences order reuse of the buffers
    import torch
    from ooc_svo_builder_b200 import SvoBuilder
    world = 4
    ctxs = [SvoBuilder(0) for _ in range(world)]
    try:
        for r, sb in enumerate(ctxs):
            sb.shard_configure(r, world)
        wins = [sb.slice_create(1000, 9) for sb in ctxs]
        for sb in ctxs:
            sb.slice_attach(wins)
        for seed, g in ((1, 128), (2, 256), (3, 64)):
            m = mg.random_soup(3000 + 100 * seed, seed=seed)
            prm = SvoBuilder.make_params(m.length, g, False, 2)
            T = m.tris.shape[0]
            for r, sb in enumerate(ctxs):
                lo, hi = sharded.slice_bounds(T, world, r)
                sb.slice_upload(m.tris[lo:hi])
            for sb in ctxs:
                sb.slice_publish(prm, T)
            tables = []
            for sb in ctxs:
                sb.partition(prm, want_counts=False)
                sb.voxelize()
                t = torch.zeros(sb.shard_table_size(), dtype=torch.int64, device="cuda")
                sb.shard_count(t.data_ptr())
                sb.synchronize()
                tables.append(t)
            merged = torch.stack(tables).sum(dim=0)
            res = []
            for r, sb in enumerate(ctxs):
                nv, nn, nd = sb.shard_emit(merged.data_ptr())
                nlo, nhi, dlo, dhi = sb.shard_ranges()
                res.append(sharded.ShardResult(r, nv, nn, nd, (nlo, nhi), (dlo, dhi), sb.fetch_nodes(nlo, nhi - nlo), sb.fetch_data(dlo, dhi - dlo), sb.stats()))
            hdr, nodes, data = sharded.assemble(res, g)
            want = oracle.build(m.tris, m.length, g, memory_limit_mb=2)
            assert hdr == want.header and nodes.tobytes() == want.nodes and data.tobytes() == want.data
    finally:
        for sb in ctxs:
            sb.close()


@pytest.mark.parametrize("world", [2, 8])
def test_remote_slices_degenerate_and_box(oracle, world):
    # zero-area / axis-aligned triangles on partition planes: the unit-level routing test must stay a superset
    _check(oracle, mg.degenerate_mix(), 256, world, limit=2, remote=True)
    _check(oracle, mg.degenerate_mix(), 64, world, remote=True)
    _check(oracle, mg.axis_aligned_box(), 256, world, limit=3, remote=True)


def test_remote_speculative_builds_and_collective_retry(oracle):
    """The same sharded contexts run three jobs over the peer-memory exchange: the second repeats the first (speculative
    builds, no host read-back), the third is larger: some rank's lists overflow, it poisons the exchange flag and ALL
    ranks answer SVO_E_RETRY in the same step; the repeated build is sized and byte-exact."""
    import torch
    from ooc_svo_builder_b200 import SvoBuilder
    world = 4
    small, big = mg.icosphere(4), mg.random_soup(6000, seed=31, large_frac=0.02)
    cap = 6000
    ctxs = [SvoBuilder(0) for _ in range(world)]
    try:
        for r, sb in enumerate(ctxs):
            sb.shard_configure(r, world)
        wins = [sb.slice_create(cap, 9) for sb in ctxs]
        for sb in ctxs:
            sb.slice_attach(wins)
        for mesh, g in ((small, 128), (small, 128), (big, 256), (big, 256), (small, 128)):
            prm = SvoBuilder.make_params(mesh.length, g, False, 2)
            T = mesh.tris.shape[0]
            for r, sb in enumerate(ctxs):
                lo, hi = sharded.slice_bounds(T, world, r)
                sb.slice_upload(mesh.tris[lo:hi])
            for sb in ctxs:
                sb.slice_publish(prm, T)
            tables = []
            for sb in ctxs:
                sb.partition(prm, want_counts=False)
                sb.voxelize()
                tables.append(torch.zeros(sb.shard_table_size(), dtype=torch.int64, device="cuda"))
            res = sharded.shard_build_single_process(ctxs, tables, True)
            hdr, nodes, data = sharded.assemble(res, g)
            want = oracle.build(mesh.tris, mesh.length, g, memory_limit_mb=2)
            assert hdr == want.header and nodes.tobytes() == want.nodes and data.tobytes() == want.data
    finally:
        for sb in ctxs:
            sb.close()
