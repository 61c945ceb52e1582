"""Shared parity cases: (name, mesh factory, gridsize, kwargs). Used by the golden
generator (reference run in the dev container), the CPU oracle tests and the GPU tests."""
from ooc_svo_builder_b200 import meshgen as mg


def _payload(m):
    return mg.Mesh(mg.with_payload(m.tris), m.length)


CASES = [
    # name, factory, gridsize, kwargs (memory_limit_mb, levels, color)
    ("c1_icosphere_256", lambda: mg.icosphere(6), 256, {}),
    ("c1_icosphere_256_p8", lambda: mg.icosphere(6), 256, {"memory_limit_mb": 3}),
    ("icosphere3_g2", lambda: mg.icosphere(3), 2, {}),
    ("icosphere3_g4", lambda: mg.icosphere(3), 4, {}),
    ("icosphere3_g8", lambda: mg.icosphere(3), 8, {}),
    ("icosphere3_g32", lambda: mg.icosphere(3), 32, {}),
    ("icosphere3_g512", lambda: mg.icosphere(3), 512, {}),
    ("f5_plane_p1", mg.single_triangle_on_partition_plane, 256, {}),
    ("f5_plane_p8", mg.single_triangle_on_partition_plane, 256, {"memory_limit_mb": 3}),
    ("box_128", mg.axis_aligned_box, 128, {"memory_limit_mb": 2}),
    ("box_256_p8", mg.axis_aligned_box, 256, {"memory_limit_mb": 3}),
    ("degenerate_64", mg.degenerate_mix, 64, {}),
    ("degenerate_256_p8", mg.degenerate_mix, 256, {"memory_limit_mb": 2}),
    ("empty_64", mg.empty_mesh, 64, {}),
    ("empty_2", mg.empty_mesh, 2, {}),
    ("soup0_256_p8", lambda: mg.random_soup(2500, seed=0, large_frac=0.03), 256, {"memory_limit_mb": 2}),
    ("soup1_256_p8", lambda: mg.random_soup(2500, seed=1, large_frac=0.03), 256, {"memory_limit_mb": 2}),
    ("soup11_512_p64", lambda: mg.random_soup(1500, seed=11), 512, {"memory_limit_mb": 2}),
    ("sphere200_512", lambda: mg.displaced_sphere(200, 200, seed=1), 512, {}),
    ("payload_ico5_128", lambda: _payload(mg.icosphere(5)), 128, {}),
    ("payload_ico5_256_p8", lambda: _payload(mg.icosphere(5)), 256, {"memory_limit_mb": 3}),
    ("payload_terrain_fixed", lambda: mg.terrain(60, seed=2), 128, {"color": "fixed"}),
    ("payload_terrain_linear", lambda: mg.terrain(60, seed=2), 128, {"color": "linear"}),
    ("payload_terrain_normal", lambda: mg.terrain(60, seed=2), 128, {"color": "normal"}),
    ("payload_terrain_256_p8", lambda: mg.terrain(120, seed=2), 256, {"memory_limit_mb": 3}),
    ("payload_soup4_128", lambda: mg.random_soup(1500, seed=4, payload=True), 128, {"memory_limit_mb": 2}),
    ("payload_ico4_levels", lambda: _payload(mg.icosphere(4)), 64, {"levels": True}),
    ("payload_terrain_levels_p8", lambda: mg.terrain(120, seed=2), 256, {"memory_limit_mb": 3, "levels": True}),
    ("payload_soup_levels_normal", lambda: mg.random_soup(800, seed=6, payload=True), 128, {"levels": True, "color": "normal"}),
    ("binary_levels_64", lambda: mg.random_soup(300, seed=3), 64, {"levels": True}),
]

# cases whose complete reference output files are committed (small)
FULL_FILE_CASES = ("f5_plane_p1", "f5_plane_p8", "empty_64", "icosphere3_g8")
