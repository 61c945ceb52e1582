import numpy as np

from ooc_svo_builder_b200 import meshgen as mg


def test_tri_roundtrip(tmp_path):
    m = mg.random_soup(50, seed=3, payload=True)
    hdr = mg.write_tri(str(tmp_path / "m"), m)
    txt = open(hdr).read().split()
    assert txt[0] == "#tri" and "geo_only" in txt and txt[-1] == "END"
    r = mg.read_tri(hdr)
    assert r.tris.shape == (50, 21) and np.array_equal(r.tris, m.tris) and abs(r.length - 1.9) < 1e-6
    assert (tmp_path / "m.tridata").stat().st_size == 50 * 84


def test_generators_stay_in_cube():
    for m in (mg.icosphere(3), mg.displaced_sphere(50, 60, seed=1), mg.terrain(20, seed=2), mg.random_soup(100), mg.thin_shell(40),
              mg.axis_aligned_box(), mg.degenerate_mix(), mg.single_triangle_on_partition_plane()):
        v = m.tris[:, :9]
        assert v.dtype == np.float32 and v.min() >= 0 and v.max() <= m.length
    assert mg.icosphere(6).n_triangles == 81920
    assert mg.displaced_sphere(10, 20).n_triangles == 400
    assert mg.empty_mesh().n_triangles == 0
