"""Synthetic `.tri` / `.tridata` generators (numpy, fixed seeds).

The reference's input contract (reference
src/libs/libtri/include/tri_tools.h:75-128, tri_util.h:7-55):

* ``<base>.tri``      text header ``#tri 1 / ntriangles / geo_only / bbox / END``
* ``<base>.tridata``  packed little-endian float32 records, 9 floats (binary /
  ``geo_only 1``: v0 v1 v2) or 21 floats (payload / ``geo_only 0``:
  v0 v1 v2 normal c0 c1 c2)

Vertices live in ``[0, L]^3`` with ``L = bbox.max[0] - bbox.min[0]`` — the
origin shift is tri_convert's job in the reference
(src/tri_convert/tri_convert.cpp:92-100) and an input invariant here.

These meshes are the BASELINE.json workloads (C1..C5) plus small edge-case
meshes for the parity tests.  Nothing here touches the GPU.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np


@dataclass
class Mesh:
    """Triangle soup. ``tris`` is (T, 9) float32 (binary) or (T, 21) float32 (payload)."""

    tris: np.ndarray
    length: float  # L: cube side; header bbox is 0 0 0 L L L

    @property
    def n_triangles(self) -> int:
        return int(self.tris.shape[0])

    @property
    def payload(self) -> bool:
        return self.tris.shape[1] == 21

    def geometry_only(self) -> "Mesh":
        return Mesh(np.ascontiguousarray(self.tris[:, :9]), self.length)


# ----------------------------------------------------------------------------
# file IO
# ----------------------------------------------------------------------------

def _fmt_float(x: float) -> str:
    # C++ `ostream << float` default formatting = %g with 6 significant digits.
    return "%g" % float(np.float32(x))


def write_tri(base: str, mesh: Mesh) -> str:
    """Write ``<base>.tri`` + ``<base>.tridata``; returns the header path."""
    tris = np.ascontiguousarray(mesh.tris, dtype="<f4")
    assert tris.ndim == 2 and tris.shape[1] in (9, 21)
    os.makedirs(os.path.dirname(os.path.abspath(base)), exist_ok=True)
    with open(base + ".tridata", "wb") as f:
        f.write(tris.tobytes())
    L = _fmt_float(mesh.length)
    with open(base + ".tri", "w") as f:
        f.write("#tri 1\n")
        f.write("ntriangles %d\n" % tris.shape[0])
        f.write("geo_only %d\n" % (1 if tris.shape[1] == 9 else 0))
        f.write("bbox  0 0 0 %s %s %s\n" % (L, L, L))
        f.write("END\n")
    return base + ".tri"


def read_tri(header_path: str) -> Mesh:
    info = {}
    with open(header_path) as f:
        toks = f.read().split()
    assert toks[0] == "#tri"
    i = 2
    while i < len(toks) and toks[i] != "END":
        if toks[i] == "ntriangles":
            info["n"] = int(toks[i + 1]); i += 2
        elif toks[i] == "geo_only":
            info["geo"] = int(toks[i + 1]); i += 2
        elif toks[i] == "bbox":
            info["bbox"] = [float(t) for t in toks[i + 1:i + 7]]; i += 7
        else:
            i += 1
    w = 9 if info.get("geo", 0) else 21
    data = np.fromfile(header_path[:-4] + ".tridata", dtype="<f4").reshape(info["n"], w)
    return Mesh(data, np.float32(info["bbox"][3]) - np.float32(info["bbox"][0]))


# ----------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------

def _soup(verts: np.ndarray, faces: np.ndarray) -> np.ndarray:
    v = verts.astype(np.float32)
    return np.ascontiguousarray(v[faces].reshape(-1, 9))


def _shift_into_cube(verts: np.ndarray, length: float | None, margin: float = 0.0):
    """Translate so min is `margin`, return (verts, L)."""
    verts = verts - verts.min(axis=0) + margin
    ext = float(verts.max()) + margin
    if length is None:
        length = ext
    assert ext <= length * (1 + 1e-6), (ext, length)
    return verts, float(length)


def with_payload(tris9: np.ndarray, colour_fn=None) -> np.ndarray:
    """Attach face normal + per-vertex colours -> (T, 21) float32 records."""
    t = tris9.reshape(-1, 3, 3).astype(np.float32)
    e0 = t[:, 1] - t[:, 0]
    e1 = t[:, 2] - t[:, 1]
    n = np.cross(e0.astype(np.float64), e1.astype(np.float64))
    ln = np.linalg.norm(n, axis=1, keepdims=True)
    ln[ln == 0] = 1.0
    n = (n / ln).astype(np.float32)
    if colour_fn is None:
        def colour_fn(p):  # smooth, position-derived colour in [0,1]
            q = p.astype(np.float64)
            return np.stack([0.5 + 0.5 * np.sin(3.1 * q[..., 0] + 0.3),
                             0.5 + 0.5 * np.sin(2.3 * q[..., 1] + 1.1),
                             0.5 + 0.5 * np.sin(1.7 * q[..., 2] + 2.2)], axis=-1)
    cols = colour_fn(t).astype(np.float32)  # (T,3,3)
    out = np.concatenate([t.reshape(-1, 9), n, cols.reshape(-1, 9)], axis=1)
    return np.ascontiguousarray(out, dtype=np.float32)


# ----------------------------------------------------------------------------
# C1: icosphere (SURVEY.md §8c known answer: 6 subdivisions, 81 920 triangles)
# ----------------------------------------------------------------------------

def icosphere(subdivisions: int = 6) -> Mesh:
    """12-vertex icosahedron on the unit sphere, midpoint subdivision in
    float64, translated by -min, cast to float32; header bbox 0 0 0 2 2 2."""
    p = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, p, 0], [1, p, 0], [-1, -p, 0], [1, -p, 0],
                  [0, -1, p], [0, 1, p], [0, -1, -p], [0, 1, -p],
                  [p, 0, -1], [p, 0, 1], [-p, 0, -1], [-p, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
                  [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
                  [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
                  [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    for _ in range(subdivisions):
        a, b, c = f[:, 0], f[:, 1], f[:, 2]
        edges = np.concatenate([np.stack([a, b], 1), np.stack([b, c], 1), np.stack([c, a], 1)])
        key = np.sort(edges, axis=1)
        uniq, inv = np.unique(key, axis=0, return_inverse=True)
        inv = inv.reshape(-1)
        mid = v[uniq[:, 0]] + v[uniq[:, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        base = len(v)
        v = np.concatenate([v, mid])
        nf = len(f)
        ab, bc, ca = base + inv[:nf], base + inv[nf:2 * nf], base + inv[2 * nf:]
        f = np.concatenate([np.stack([a, ab, ca], 1), np.stack([b, bc, ab], 1),
                            np.stack([c, ca, bc], 1), np.stack([ab, bc, ca], 1)])
    v = v - v.min(axis=0)
    return Mesh(_soup(v, f), 2.0)


# ----------------------------------------------------------------------------
# C2 / C4 / C5: lat-long sphere with seeded radial displacement
# ----------------------------------------------------------------------------

def _value_noise_sphere(theta, phi, rng, octaves=4):
    """Smooth seeded displacement: a few random low-order harmonics."""
    out = np.zeros_like(theta)
    for o in range(octaves):
        k1, k2 = rng.integers(1, 6 + 6 * o, size=2)
        ph1, ph2 = rng.uniform(0, 2 * np.pi, size=2)
        out += np.sin(k1 * theta + ph1) * np.sin(k2 * phi + ph2) / (1 + o)
    return out / np.abs(out).max()


def displaced_sphere(nu: int = 1000, nv: int = 1000, seed: int = 1, amp: float = 0.05,
                     length: float = 2.0, radius: float | None = None) -> Mesh:
    """nu x nv lat-long quads split in two -> 2*nu*nv triangles; radius
    r*(1+amp*noise(seed)). nu=nv=1000 is BASELINE.json config 2 (2 M triangles)."""
    rng = np.random.default_rng(seed)
    if radius is None:
        radius = 0.5 * length / (1.0 + amp) * 0.98
    th = np.linspace(0.0, np.pi, nu + 1)            # polar
    ph = np.linspace(0.0, 2 * np.pi, nv + 1)        # azimuth (seam duplicated)
    T, Pp = np.meshgrid(th, ph, indexing="ij")
    disp = _value_noise_sphere(T, Pp, rng) if amp != 0 else np.zeros_like(T)
    # make the seam and the poles watertight
    disp[:, -1] = disp[:, 0]
    disp[0, :] = disp[0, 0]
    disp[-1, :] = disp[-1, 0]
    r = radius * (1.0 + amp * disp)
    x = r * np.sin(T) * np.cos(Pp)
    y = r * np.sin(T) * np.sin(Pp)
    z = r * np.cos(T)
    v = np.stack([x, y, z], -1).reshape(-1, 3) + 0.5 * length
    idx = np.arange((nu + 1) * (nv + 1)).reshape(nu + 1, nv + 1)
    a, b, c, d = idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]
    f = np.concatenate([np.stack([a, b, c], -1).reshape(-1, 3),
                        np.stack([a, c, d], -1).reshape(-1, 3)])
    assert v.min() >= 0 and v.max() <= length
    return Mesh(_soup(v, f), float(length))


def thin_shell(n: int = 7071, seed: int = 4, length: float = 2.0) -> Mesh:
    """C5: undisplaced-ish thin shell, 2*n*n triangles (n=7071 -> 100 M)."""
    return displaced_sphere(n, n, seed=seed, amp=0.01, length=length)


# ----------------------------------------------------------------------------
# C3: terrain heightfield with payload (face normal + vertex colours)
# ----------------------------------------------------------------------------

def terrain(n: int = 2237, seed: int = 2, length: float = 1.9, payload: bool = True) -> Mesh:
    """n x n quads -> 2*n*n triangles (n=2237 -> 10.0 M). Height = 4 seeded
    octaves; colour = f(height). L=1.9 so the unit length is not a power of two."""
    rng = np.random.default_rng(seed)
    g = np.linspace(0.0, 1.0, n + 1)
    X, Y = np.meshgrid(g, g, indexing="ij")
    h = np.zeros_like(X)
    for o in range(4):
        fx, fy = rng.uniform(1.0, 3.0, size=2) * (2 ** o)
        px, py = rng.uniform(0, 2 * np.pi, size=2)
        h += np.sin(2 * np.pi * fx * X + px) * np.cos(2 * np.pi * fy * Y + py) / (2 ** o)
    h = (h - h.min()) / (h.max() - h.min())
    v = np.stack([X * length, Y * length, (0.25 + 0.5 * h) * length], -1).reshape(-1, 3)
    idx = np.arange((n + 1) * (n + 1)).reshape(n + 1, n + 1)
    a, b, c, d = idx[:-1, :-1], idx[1:, :-1], idx[1:, 1:], idx[:-1, 1:]
    f = np.concatenate([np.stack([a, b, c], -1).reshape(-1, 3),
                        np.stack([a, c, d], -1).reshape(-1, 3)])
    tris = _soup(v, f)
    if payload:
        def col(p):
            hh = (p[..., 2].astype(np.float64) / length - 0.25) / 0.5
            return np.stack([hh, 1.0 - np.abs(2 * hh - 1.0), 1.0 - hh], -1)
        tris = with_payload(tris, col)
    return Mesh(tris, float(length))


# ----------------------------------------------------------------------------
# small meshes for the parity tests
# ----------------------------------------------------------------------------

def random_soup(n: int = 2000, seed: int = 0, length: float = 1.9,
                small: float = 0.02, large_frac: float = 0.02, payload: bool = False) -> Mesh:
    """Random triangles: mostly small (edge ~ `small`*L), a few spanning a large
    part of the cube (exercise the medium / large bbox work classes)."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(0.0, length, size=(n, 1, 3))
    scale = np.where(rng.uniform(size=(n, 1, 1)) < large_frac, 0.6, small) * length
    v = c + rng.uniform(-1.0, 1.0, size=(n, 3, 3)) * scale
    v = np.clip(v, 0.0, length)
    tris = np.ascontiguousarray(v.reshape(n, 9).astype(np.float32))
    if payload:
        cols = rng.uniform(0.0, 1.0, size=(n, 3, 3))
        tris = with_payload(tris, lambda p: cols)
    return Mesh(tris, float(length))


def single_triangle_on_partition_plane() -> Mesh:
    """one triangle in the plane x=1.0 of the [0,2]^3 cube.
    378 voxels at -s 256 with one partition, 756 with eight."""
    t = np.array([[1, .3, .3, 1, .5, .3, 1, .3, .5]], dtype=np.float32)
    return Mesh(t, 2.0)


def axis_aligned_box(length: float = 2.0, lo: float = 0.5, hi: float = 1.5) -> Mesh:
    """12 triangles, faces exactly on voxel / partition planes (touching cases)."""
    c = np.array([[x, y, z] for x in (lo, hi) for y in (lo, hi) for z in (lo, hi)], dtype=np.float64)
    q = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    f = []
    for a, b, cc, d in q:
        f += [[a, b, cc], [a, cc, d]]
    return Mesh(_soup(c, np.array(f)), float(length))


def degenerate_mix(seed: int = 5, length: float = 2.0) -> Mesh:
    """Zero-area triangles (repeated / collinear vertices), vertices on the cube
    boundary (coordinate == L -> grid index g, clamped), plus a few normal ones."""
    rng = np.random.default_rng(seed)
    t = []
    p = rng.uniform(0.2, 1.8, size=(4, 3))
    t.append(np.concatenate([p[0], p[0], p[0]]))                   # point
    t.append(np.concatenate([p[1], p[1], p[2]]))                   # segment
    t.append(np.concatenate([p[1], 0.5 * (p[1] + p[3]), p[3]]))    # collinear
    t.append(np.array([length, length, length, length, 1.5, length, 1.5, length, length]))
    t.append(np.array([0, 0, 0, 0, 0.3, 0, 0.3, 0, 0]))
    t.append(np.array([0.1, 0.1, length, 0.9, 0.1, length, 0.1, 0.9, length]))
    soup = random_soup(40, seed=seed + 1, length=length).tris
    return Mesh(np.concatenate([np.array(t, dtype=np.float32), soup]), float(length))


def empty_mesh(length: float = 2.0, payload: bool = False) -> Mesh:
    return Mesh(np.zeros((0, 21 if payload else 9), dtype=np.float32), float(length))


CONFIGS = {
    # name: (generator kwargs summary) — see BASELINE.json `configs`
    "c1_icosphere_256": dict(gen="icosphere", args=dict(subdivisions=6), gridsize=256, payload=False),
    "c2_displaced_sphere_1024": dict(gen="displaced_sphere", args=dict(nu=1000, nv=1000, seed=1), gridsize=1024, payload=False),
    "c3_terrain_2048_payload": dict(gen="terrain", args=dict(n=2237, seed=2), gridsize=2048, payload=True),
    "c4_sphere_4096": dict(gen="displaced_sphere", args=dict(nu=5000, nv=5000, seed=3), gridsize=4096, payload=False),
    "c5_shell_8192": dict(gen="thin_shell", args=dict(n=7071, seed=4), gridsize=8192, payload=False),
}


def make(name: str) -> Mesh:
    cfg = CONFIGS[name]
    return globals()[cfg["gen"]](**cfg["args"])
