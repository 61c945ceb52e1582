"""CPU tests of the multi-GPU host logic: the shard plan, and the table exchange run for real over
torch.distributed (gloo, world_size 2) with per-rank tables derived from the oracle's voxels."""
import os
import subprocess
import sys

import numpy as np
import pytest

from ooc_svo_builder_b200 import meshgen as mg
from ooc_svo_builder_b200 import sharded

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("g,P,world", [(2048, 8, 2), (2048, 8, 4), (2048, 8, 8), (4096, 64, 8), (8192, 512, 8), (1024, 1, 8), (256, 8, 2), (64, 1, 2)])
def test_plan_tiles_the_grid(g, P, world):
    plans = [sharded.plan(g, P, world, r) for r in range(world)]
    assert plans[0].morton_range[0] == 0 and plans[-1].morton_range[1] == g ** 3
    for a, b in zip(plans, plans[1:]):
        assert a.morton_range[1] == b.morton_range[0]                 # contiguous Morton ranges
        assert a.own_entries[1] <= b.own_entries[0] or a.own_entries == b.own_entries
    p = plans[0]
    assert p.dc >= p.k                                                # chunks never straddle a logical partition
    assert 64 ** (p.top_local_level + 1) <= 8 ** (p.depth - p.dc)    # a top-local word stays inside one chunk
    if P >= world:                                                    # whole partitions per rank
        owned = [set(range(q.partition_range[0], q.partition_range[1] + 1)) for q in plans]
        assert sum(len(o) for o in owned) == P and set().union(*owned) == set(range(P))


def test_plan_rejects_tiny_grids():
    with pytest.raises(ValueError):
        sharded.plan(4, 1, 8, 0)


def test_merge_tables_requires_disjoint_entries():
    a = np.zeros(16, dtype=np.int64); b = np.zeros(16, dtype=np.int64)
    a[0:4] = [5, 9, 3, 0]; b[8:12] = [1, 2, 1, 0]
    m = sharded.merge_tables([a, b])
    assert m[1] == 9 and m[9] == 2
    with pytest.raises(AssertionError):
        sharded.merge_tables([a, a])


WORKER = r"""
import os, sys, json
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from ooc_svo_builder_b200 import meshgen as mg, sharded
from oracle import oracle as O
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
g = 64
mesh = mg.icosphere(3)
codes = O.voxelize(mesh.tris, mesh.length, g)            # P == 1: one partition, all triangles
p = sharded.plan(g, 1, world, rank)
mine = sharded.subtree_table_from_codes(codes, g, p)
t = torch.from_numpy(mine.copy())
dist.all_reduce(t)                                       # the one exchange of the sharded build
merged = t.numpy()
# every rank must now be able to derive the global counts
D, J = p.depth, p.top_local_level
dr = D - 2 * (J + 1)
keys = np.flatnonzero(merged[0::4])
n_nodes = int(merged[1::4].sum())
for d in range(0, dr + 1):
    n_nodes += np.unique(keys >> (3 * (dr - d))).size
want = O.build(mesh.tris, mesh.length, g)
ok = (n_nodes == want.n_nodes) and (int(merged[2::4].sum()) == want.n_voxels)
own = (np.flatnonzero(mine[0::4]) >= p.own_entries[0]).all() and (np.flatnonzero(mine[0::4]) < p.own_entries[1]).all()
gathered = [None] * world
dist.all_gather_object(gathered, mine)
# ... and its own range of the node file with the library's host-side merge (no GPU involved)
from ooc_svo_builder_b200 import SvoBuilder
from ooc_svo_builder_b200.api import shard_layout_from_table
lay, rpos, rwords = shard_layout_from_table(SvoBuilder.make_params(mesh.length, g, False), rank, world, merged)
want_nodes = np.frombuffer(want.nodes, dtype=np.uint64).reshape(-1, 3)
rec_ok = bool((want_nodes[rpos.astype(np.int64)] == rwords).all())
ranges = [None] * world
dist.all_gather_object(ranges, (lay["node_lo"], lay["node_hi"], lay["n_nodes"], rec_ok))
if rank == 0:
    ref = sharded.merge_tables(gathered)
    tiled = ranges[0][0] == 0 and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:])) and ranges[-1][1] == want.n_nodes
    print(json.dumps({{"ok": bool(ok and own and (ref == merged).all() and tiled and all(r[3] and r[2] == want.n_nodes for r in ranges)),
                      "n_nodes": n_nodes, "want": want.n_nodes, "ranges": ranges}}))
dist.destroy_process_group()
"""


def test_table_exchange_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", str(script)], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    import json
    r = json.loads(line)
    assert r["ok"], r


def test_node_count_closed_form(oracle):
    # n_nodes = 1 + number of distinct Morton prefixes over all depths
    m = mg.random_soup(400, seed=8)
    codes = oracle.voxelize(m.tris, m.length, 128)
    assert sharded.node_count_from_codes(codes, 128) == oracle.build(m.tris, m.length, 128).n_nodes


@pytest.mark.parametrize("T,world", [(0, 2), (1, 8), (127, 4), (128, 4), (1000, 8), (2 ** 20 + 3, 8)])
def test_slices_tile_the_triangle_file(T, world):
    # rank r holds the r-th slice in file order: contiguous, disjoint, complete, equal capacity
    b = [sharded.slice_bounds(T, world, r) for r in range(world)]
    assert b[0][0] == 0 and b[-1][1] == T
    assert all(x[1] == y[0] for x, y in zip(b, b[1:]))
    cap = (T + world - 1) // world
    assert all(0 <= hi - lo <= cap for lo, hi in b)


def _codes_from_octree(nodes: bytes, gridsize: int) -> np.ndarray:
    """Ascending Morton codes of the leaves of an .octreenodes image (root = last record, octree_io.h:62-66, Node.h:13-65)."""
    rec = np.frombuffer(nodes, dtype=np.uint8).reshape(-1, 24)
    base = rec[:, 8:16].copy().view(np.uint64).reshape(-1)
    off = rec[:, 16:24].copy().view(np.int8)
    D = int(np.log2(gridsize))
    out = []
    stack = [(len(rec) - 1, 0, 0)]
    while stack:
        i, depth, prefix = stack.pop()
        if depth == D:
            out.append(prefix)
            continue
        for ch in range(8):
            if off[i, ch] >= 0:
                stack.append((int(base[i]) + int(off[i, ch]), depth + 1, prefix * 8 + ch))
    return np.array(sorted(out), dtype=np.uint64)


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("case", ["ico_g64_p1", "soup_g128_p1", "soup_g256_p8", "ico_g32_odd"])
def test_host_merge_against_the_oracle(oracle, world, case):
    """The merge of the shared upper levels (svo_shard_emit's host part, exposed as svo_shard_layout_from_table) needs no
    GPU: feed it the table built from the oracle's voxels and compare the counts, the per-rank file ranges and every
    upper-level record it produces with the oracle's .octreenodes image, byte for byte."""
    from ooc_svo_builder_b200 import SvoBuilder, estimate_partitions
    from ooc_svo_builder_b200.api import shard_layout_from_table
    mesh, g, limit = {"ico_g64_p1": (mg.icosphere(3), 64, 2048), "soup_g128_p1": (mg.random_soup(500, seed=2, large_frac=0.02), 128, 2048),
                      "soup_g256_p8": (mg.random_soup(600, seed=5, large_frac=0.02), 256, 2), "ico_g32_odd": (mg.icosphere(2), 32, 2048)}[case]
    P = estimate_partitions(g, limit)
    want = oracle.build(mesh.tris, mesh.length, g, memory_limit_mb=limit)
    codes = _codes_from_octree(want.nodes, g)          # the voxels of THIS build (with P > 1 they depend on the partitioning)
    assert codes.size == want.n_voxels
    if P == 1:
        assert (codes == oracle.voxelize(mesh.tris, mesh.length, g)).all()
    try:
        plans = [sharded.plan(g, P, world, r) for r in range(world)]
    except ValueError:
        pytest.skip("grid too small for this many shards")
    merged = sharded.merge_tables([sharded.subtree_table_from_codes(codes, g, p) for p in plans])
    prm = SvoBuilder.make_params(mesh.length, g, False, limit)
    want_nodes = np.frombuffer(want.nodes, dtype=np.uint64).reshape(-1, 3)
    pos_end, seen = 0, 0
    for r in range(world):
        lay, rpos, rwords = shard_layout_from_table(prm, r, world, merged)
        assert lay["n_voxels"] == want.n_voxels and lay["n_nodes"] == want.n_nodes
        assert lay["node_lo"] == pos_end or lay["node_lo"] == lay["node_hi"]
        pos_end = max(pos_end, lay["node_hi"])
        assert ((rpos >= lay["node_lo"]) & (rpos < lay["node_hi"])).all()
        assert (want_nodes[rpos.astype(np.int64)] == rwords).all(), "an upper-level record differs from the oracle's node file"
        seen += len(rpos)
    assert pos_end == want.n_nodes
    assert seen > 0


def test_host_merge_rejects_bad_arguments():
    from ooc_svo_builder_b200 import SvoBuilder, SvoError
    from ooc_svo_builder_b200.api import shard_layout_from_table
    prm = SvoBuilder.make_params(2.0, 64, False)
    p = sharded.plan(64, 1, 2, 0)
    good = np.zeros(p.table_entries * 4, dtype=np.int64)
    lay, rpos, _ = shard_layout_from_table(prm, 0, 2, good)                   # empty grid: one null root, on rank 0
    assert lay["n_voxels"] == 0 and lay["n_nodes"] == 1 and (lay["node_lo"], lay["node_hi"]) == (0, 1) and len(rpos) == 0
    lay1, _, _ = shard_layout_from_table(prm, 1, 2, good)
    assert (lay1["node_lo"], lay1["node_hi"]) == (1, 1)
    for bad in (lambda: shard_layout_from_table(prm, 0, 3, good),             # world must be a power of two
                lambda: shard_layout_from_table(prm, 2, 2, good),             # rank out of range
                lambda: shard_layout_from_table(prm, 0, 2, good[:-4]),        # table of the wrong size
                lambda: shard_layout_from_table(SvoBuilder.make_params(2.0, 4, False), 0, 8, good),          # grid too small for 8 shards
                lambda: shard_layout_from_table(SvoBuilder.make_params(2.0, 64, False, levels=True), 0, 2, good)):   # -levels is not sharded
        with pytest.raises(SvoError):
            bad()
