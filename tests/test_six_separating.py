"""The opt-in 6-separating ("thin") Schwarz-Seidel variant (svo_params.separability = 6, CLI -sep 6).

The reference only implements the conservative 26-separating test: there is no reference output to pin
this mode to. The definition is restated in oracle/svo_oracle.c (header comment there); the CPU tests below check the
restatement's DEFINING properties, the GPU tests hold the CUDA path to the restatement bit for bit."""
import numpy as np
import pytest

from ooc_svo_builder_b200 import meshgen as mg


def _grid(codes, g):
    from oracle import oracle as O
    c = np.asarray(codes, dtype=np.uint64)
    x = np.zeros(c.size, dtype=np.int64); y = np.zeros_like(x); z = np.zeros_like(x)
    for b in range(21):
        x |= ((c >> np.uint64(3 * b)) & np.uint64(1)).astype(np.int64) << b
        y |= ((c >> np.uint64(3 * b + 1)) & np.uint64(1)).astype(np.int64) << b
        z |= ((c >> np.uint64(3 * b + 2)) & np.uint64(1)).astype(np.int64) << b
    vol = np.zeros((g, g, g), dtype=bool)
    vol[x, y, z] = True
    return vol


def test_thin_is_a_subset_of_conservative_and_smaller(oracle):
    for m, g in ((mg.icosphere(4), 128), (mg.random_soup(800, seed=4, large_frac=0.05), 64), (mg.terrain(40, seed=2, payload=False), 64)):
        cons = oracle.voxelize(m.tris, m.length, g)
        thin = oracle.voxelize(m.tris, m.length, g, separability=6)
        assert np.isin(thin, cons).all()
        assert 0 < thin.size < cons.size


def test_thin_single_triangle_covers_exactly_the_columns_under_it(oracle):
    """The defining property, per triangle: along the dominant normal axis the thin voxelization holds the voxel(s) whose
    centre segment the plane crosses, in exactly the columns whose centre projects into the triangle -- one voxel per
    column (two only where the plane passes through a voxel face)."""
    g, L = 64, 2.0
    t = np.array([[0.1, 0.1, 0.9, 1.9, 0.2, 1.1, 0.3, 1.8, 1.3]], dtype=np.float32)      # z-dominant normal
    vol = _grid(oracle.voxelize(t, L, g, separability=6), g)
    per_column = vol.sum(axis=2)
    cx = (np.arange(g) + 0.5) * (L / g)
    px, py = np.meshgrid(cx, cx, indexing="ij")

    def edge(p, q):
        return (q[0] - p[0]) * (py - p[1]) - (q[1] - p[1]) * (px - p[0])
    a, b, c = t[0, 0:2].astype(np.float64), t[0, 3:5].astype(np.float64), t[0, 6:8].astype(np.float64)
    s1, s2, s3 = edge(a, b), edge(b, c), edge(c, a)
    inside = ((s1 >= 0) & (s2 >= 0) & (s3 >= 0)) | ((s1 <= 0) & (s2 <= 0) & (s3 <= 0))
    assert ((per_column > 0) == inside).all()
    assert per_column.max() <= 2 and (per_column[inside] >= 1).all()


def test_thin_heightfield_covers_its_columns(oracle):
    """A z-dominant heightfield over the whole grid: every column holds a voxel, except where a column centre lies
    within float rounding of an edge shared by two triangles -- the test is per triangle (as in the paper's kernel),
    without a tie-breaking fill rule, so both neighbours may reject such a centre. Watertightness of a whole MESH is
    therefore not promised by this mode (nor across creases where the dominant axis changes)."""
    g, L, n = 64, 2.0, 40
    xs = np.linspace(0, L, n + 1)
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    Z = 1.0 + 0.25 * np.sin(2.1 * X + 0.3) * np.cos(1.7 * Y + 0.2) + 0.11 * X
    v = np.stack([X, Y, Z], -1).reshape(-1, 3)
    idx = np.arange((n + 1) ** 2).reshape(n + 1, n + 1)
    a, b, c, d = idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]
    f = np.concatenate([np.stack([a, b, c], -1).reshape(-1, 3), np.stack([a, c, d], -1).reshape(-1, 3)])
    tris = v[f].reshape(-1, 9).astype(np.float32)
    vol = _grid(oracle.voxelize(tris, L, g, separability=6), g)
    covered = (vol.sum(axis=2) > 0).mean()
    assert covered > 0.995, covered
    assert vol.sum(axis=2).max() <= 3


def test_thin_axis_aligned_triangle_is_one_voxel_thick(oracle):
    t = np.array([[0.3, 0.3, 1.01, 1.7, 0.3, 1.01, 0.3, 1.7, 1.01]], dtype=np.float32)
    vol = _grid(oracle.voxelize(t, 2.0, 64, separability=6), 64)
    assert vol.any(axis=(0, 1)).sum() == 1                      # one z layer
    assert vol.sum() < _grid(oracle.voxelize(t, 2.0, 64), 64).sum()


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["ico_128", "soup_256_p8", "soup_large_256", "payload_terrain_128", "payload_ico_256_p8", "degenerate_64", "sphere_512"])
def test_thin_cuda_path_matches_the_restatement(builder, oracle, case):
    cfg = {
        "ico_128": (mg.icosphere(5), 128, 2048),
        "soup_256_p8": (mg.random_soup(2500, seed=0, large_frac=0.03), 256, 2),
        "soup_large_256": (mg.random_soup(600, seed=8, large_frac=0.3), 256, 2048),
        "payload_terrain_128": (mg.terrain(60, seed=2), 128, 2048),
        "payload_ico_256_p8": (mg.Mesh(mg.with_payload(mg.icosphere(5).tris), 2.0), 256, 3),
        "degenerate_64": (mg.degenerate_mix(), 64, 2048),
        "sphere_512": (mg.displaced_sphere(200, 200, seed=1), 512, 2048),
    }[case]
    mesh, g, lim = cfg
    got = builder.run(mesh.tris, mesh.length, g, memory_limit_mb=lim, separability=6)
    want = oracle.build(mesh.tris, mesh.length, g, memory_limit_mb=lim, separability=6)
    assert got.n_voxels == want.n_voxels and got.header == want.header
    assert got.nodes.tobytes() == want.nodes and got.data.tobytes() == want.data
    cons = oracle.build(mesh.tris, mesh.length, g, memory_limit_mb=lim)
    assert want.n_voxels <= cons.n_voxels
    # the default is untouched: separability 0 and 26 both mean the reference's conservative test
    assert builder.run(mesh.tris, mesh.length, g, memory_limit_mb=lim, separability=0).nodes.tobytes() == cons.nodes


@pytest.mark.gpu
def test_thin_sharded_and_cli(oracle, tmp_path):
    import os
    import subprocess
    from ooc_svo_builder_b200 import sharded, SvoBuilder
    m = mg.icosphere(5)
    want = oracle.build(m.tris, m.length, 256, memory_limit_mb=3, separability=6)
    # CLI: -sep 6
    hdr = mg.write_tri(str(tmp_path / "mesh"), m)
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ooc_svo_builder_b200", "bin", "svo_builder_binary")
    p = subprocess.run([exe, "-f", hdr, "-s", "256", "-l", "3", "-sep", "6"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout
    got = oracle.read_outputs(str(tmp_path / "mesh") + "256_%d" % want.n_partitions)
    assert (got.header, got.nodes, got.data) == (want.header, want.nodes, want.data), p.stdout[-1000:]
    # sharded over 4 ranks (contexts of one process)
    ctxs = [SvoBuilder(0) for _ in range(4)]
    try:
        import torch
        prm = SvoBuilder.make_params(m.length, 256, False, 3, separability=6)
        tables = []
        for r, sb in enumerate(ctxs):
            sb.shard_configure(r, 4)
            sb.set_triangles(m.tris)
            sb.partition(prm)
            sb.voxelize()
            tables.append(torch.zeros(sb.shard_table_size(), dtype=torch.int64, device="cuda"))
        res = sharded.shard_build_single_process(ctxs, tables, False)
        hdr_b, nodes, data = sharded.assemble(res, 256)
        assert hdr_b == want.header and nodes.tobytes() == want.nodes and data.tobytes() == want.data
    finally:
        for sb in ctxs:
            sb.close()
