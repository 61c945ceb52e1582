"""TEST INFRASTRUCTURE ONLY — Python access to the CPU oracle.

Two checkers live here:

* ``liboracle.so`` — our plain-C restatement (``svo_oracle.c``), loaded via ctypes;
* ``_ref/svo_builder[_binary]`` — the UNMODIFIED reference compiled by
  ``oracle/Makefile`` (present when built in the dev container; travels to the
  GPU box as prebuilt binaries), driven through its own CLI on temp files.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` leg may import this module.  The product never does.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
import tempfile
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")

COLOR_MODES = {"model": 0, "fixed": 1, "linear": 2, "normal": 3}


class _Result(C.Structure):
    _fields_ = [("n_partitions", C.c_uint64), ("n_voxels", C.c_uint64),
                ("n_nodes", C.c_uint64), ("n_data", C.c_uint64),
                ("nodes", C.POINTER(C.c_uint8)), ("data", C.POINTER(C.c_uint8))]


@dataclass
class OctreeFiles:
    """In-memory image of the three output files."""
    header: bytes
    nodes: bytes
    data: bytes
    n_partitions: int
    n_voxels: int | None = None

    @property
    def n_nodes(self) -> int:
        return len(self.nodes) // 24

    @property
    def n_data(self) -> int:
        return len(self.data) // 32


def header_bytes(gridsize: int, n_nodes: int, n_data: int) -> bytes:
    # octree_io.h:74-83
    return ("#octreeheader 1\ngridlength %d\nn_nodes %d\nn_data %d\nEND\n" % (gridsize, n_nodes, n_data)).encode()


_lib = None


def build_lib(force: bool = False) -> str:
    if force or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, "svo_oracle.c")):
        subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build_lib()
        L = C.CDLL(LIB_PATH)
        L.svo_oracle_estimate_partitions.restype = C.c_uint64
        L.svo_oracle_estimate_partitions.argtypes = [C.c_uint64, C.c_uint64]
        L.svo_oracle_text_roundtrip.restype = C.c_float
        L.svo_oracle_text_roundtrip.argtypes = [C.c_float]
        L.svo_oracle_partition.restype = C.c_int
        L.svo_oracle_partition.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_float, C.c_float,
                                           C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        L.svo_oracle_voxelize.restype = C.c_uint64
        L.svo_oracle_voxelize.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64,
                                          C.c_float, C.c_void_p, C.c_void_p]
        L.svo_oracle_build.restype = C.c_int
        L.svo_oracle_build.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_float, C.c_float, C.c_uint64,
                                       C.c_uint64, C.c_int, C.c_int, C.POINTER(_Result)]
        L.svo_oracle_build_from_codes.restype = C.c_int
        L.svo_oracle_build_from_codes.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(_Result)]
        L.svo_oracle_free.argtypes = [C.POINTER(_Result)]
        L.svo_oracle_set_separability.argtypes = [C.c_int]
        L.svo_oracle_morton_encode.restype = C.c_uint64
        L.svo_oracle_morton_encode.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        _lib = L
    return _lib


def estimate_partitions(gridsize: int, memory_limit_mb: int) -> int:
    return int(lib().svo_oracle_estimate_partitions(gridsize, memory_limit_mb))


def _take(res: _Result, gridsize: int) -> OctreeFiles:
    nodes = C.string_at(res.nodes, res.n_nodes * 24)
    data = C.string_at(res.data, res.n_data * 32)
    out = OctreeFiles(header_bytes(gridsize, res.n_nodes, res.n_data), nodes, data,
                      int(res.n_partitions), int(res.n_voxels))
    lib().svo_oracle_free(C.byref(res))
    return out


def build(tris: np.ndarray, length: float, gridsize: int, memory_limit_mb: int = 2048,
          levels: bool = False, color: str = "model", separability: int = 26) -> OctreeFiles:
    """The restatement's whole pipeline (main.cpp:281-399). bbox = 0..length.
    separability=6: the 6-separating variant (not in the reference; see svo_oracle.c)."""
    tris = np.ascontiguousarray(tris, dtype=np.float32)
    res = _Result()
    lib().svo_oracle_set_separability(separability)
    try:
        rc = lib().svo_oracle_build(tris.ctypes.data, tris.shape[0], tris.shape[1], 0.0, np.float32(length),
                                    gridsize, memory_limit_mb, int(levels), COLOR_MODES[color], C.byref(res))
    finally:
        lib().svo_oracle_set_separability(26)
    assert rc == 0
    return _take(res, gridsize)


def build_from_codes(codes: np.ndarray, gridsize: int, levels: bool = False) -> OctreeFiles:
    codes = np.ascontiguousarray(codes, dtype=np.uint64)
    res = _Result()
    rc = lib().svo_oracle_build_from_codes(codes.ctypes.data, codes.shape[0], gridsize, int(levels), C.byref(res))
    assert rc == 0
    return _take(res, gridsize)


def partition_counts(tris: np.ndarray, length: float, gridsize: int, n_partitions: int) -> np.ndarray:
    tris = np.ascontiguousarray(tris, dtype=np.float32)
    counts = np.zeros(n_partitions, dtype=np.uint64)
    lib().svo_oracle_partition(tris.ctypes.data, tris.shape[0], tris.shape[1], 0.0, np.float32(length),
                               gridsize, n_partitions, counts.ctypes.data, None)
    return counts


def voxelize(tris: np.ndarray, length: float, gridsize: int,
             morton_start: int = 0, morton_end: int | None = None, separability: int = 26) -> np.ndarray:
    """voxelize_schwarz_method over all triangles for one Morton range;
    returns the ascending Morton codes of the filled voxels."""
    tris = np.ascontiguousarray(tris, dtype=np.float32)
    if morton_end is None:
        morton_end = gridsize ** 3
    vox = np.zeros(morton_end - morton_start, dtype=np.uint8)
    rt = lib().svo_oracle_text_roundtrip
    unit = np.float32(np.float32(rt(np.float32(length))) - np.float32(rt(0.0))) / np.float32(gridsize)
    lib().svo_oracle_set_separability(separability)
    try:
        lib().svo_oracle_voxelize(tris.ctypes.data, tris.shape[1], None, tris.shape[0], morton_start, morton_end,
                                  np.float32(unit), vox.ctypes.data, None)
    finally:
        lib().svo_oracle_set_separability(26)
    return np.flatnonzero(vox).astype(np.uint64) + np.uint64(morton_start)


# ----------------------------------------------------------------------------
# the real reference, through its own CLI
# ----------------------------------------------------------------------------

def ref_available() -> bool:
    return os.access(os.path.join(REF_DIR, "svo_builder_binary"), os.X_OK) and \
        os.access(os.path.join(REF_DIR, "svo_builder"), os.X_OK)


def ref_exe(payload: bool) -> str:
    return os.path.join(REF_DIR, "svo_builder" if payload else "svo_builder_binary")


def read_outputs(base: str) -> OctreeFiles:
    """Read <base>.octree/.octreenodes/.octreedata; base includes `<g>_<P>`."""
    with open(base + ".octree", "rb") as f:
        header = f.read()
    with open(base + ".octreenodes", "rb") as f:
        nodes = f.read()
    with open(base + ".octreedata", "rb") as f:
        data = f.read()
    P = int(base.rsplit("_", 1)[1])
    return OctreeFiles(header, nodes, data, P)


def run_cli(exe: str, tri_header: str, gridsize: int, memory_limit_mb: int | None = None,
            levels: bool = False, color: str | None = None, sparse: int | None = None,
            extra: list[str] | None = None, timeout: float = 3600) -> tuple[OctreeFiles, str]:
    """Run a svo_builder-compatible executable; returns (files, stdout)."""
    args = [exe, "-f", tri_header, "-s", str(gridsize)]
    if memory_limit_mb is not None:
        args += ["-l", str(memory_limit_mb)]
    if sparse is not None:
        args += ["-d", str(sparse)]
    if levels:
        args += ["-levels"]
    if color is not None:
        args += ["-c", color]
    if extra:
        args += extra
    p = subprocess.run(args, capture_output=True, text=True, timeout=timeout)
    P = estimate_partitions(gridsize, 2048 if memory_limit_mb is None else memory_limit_mb)
    base = tri_header[:-4] + "%d_%d" % (gridsize, P)
    if not os.path.exists(base + ".octree"):
        raise RuntimeError("no output from %s\n%s\n%s" % (" ".join(args), p.stdout[-2000:], p.stderr[-2000:]))
    return read_outputs(base), p.stdout


def ref_build(mesh, gridsize: int, memory_limit_mb: int | None = None, levels: bool = False,
              color: str | None = None, sparse: int | None = None, workdir: str | None = None) -> OctreeFiles:
    """Write the mesh to a temp dir and run the unmodified reference on it."""
    from ooc_svo_builder_b200 import meshgen
    own = workdir is None
    d = tempfile.mkdtemp(prefix="svo_ref_") if own else workdir
    try:
        hdr = meshgen.write_tri(os.path.join(d, "m"), mesh)
        files, out = run_cli(ref_exe(mesh.payload), hdr, gridsize, memory_limit_mb, levels, color, sparse)
        for line in out.splitlines():
            if line.startswith("Total amount of voxels:"):
                files.n_voxels = int(line.split(":")[1])
        return files
    finally:
        if own:
            shutil.rmtree(d, ignore_errors=True)
