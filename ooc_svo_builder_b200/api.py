"""ctypes binding of the C ABI in ``include/svo_b200.h`` (libsvo_b200.so).

Host-side mirror of the reference's in-process seams for the voxelize-and-build
path (reference src/svo_builder/main.cpp:298-389): ``estimate_partitions`` →
``partition`` → ``voxelize`` → ``build`` → ``fetch``.  Everything that computes
runs in hand-written sm_100a kernels behind the C ABI; there is no CPU
fallback: if the shared library is missing or no B200 is visible the calls
raise.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsvo_b200.so")

COLOR_MODES = {"model": 0, "fixed": 1, "linear": 2, "normal": 3}
NODE_BYTES = 24
DATA_BYTES = 32

# every symbol include/svo_b200.h declares (tests check the .so exports all of them)
ABI_SYMBOLS = [
    "svo_ctx_create", "svo_ctx_destroy", "svo_ctx_set_stream", "svo_last_error", "svo_version",
    "svo_estimate_partitions", "svo_text_roundtrip_float",
    "svo_set_triangles", "svo_set_triangles_device", "svo_triangles_begin", "svo_triangles_append", "svo_partition", "svo_voxelize", "svo_build",
    "svo_fetch_nodes", "svo_fetch_data", "svo_device_nodes", "svo_device_data", "svo_fetch_voxel_codes",
    "svo_shard_configure", "svo_shard_table_size", "svo_shard_count", "svo_shard_emit", "svo_shard_ranges",
    "svo_shard_dispatch_create", "svo_shard_dispatch_attach", "svo_shard_dispatch_count", "svo_shard_dispatch_send",
    "svo_shard_dispatch_finish", "svo_ipc_export", "svo_ipc_open", "svo_ipc_close",
    "svo_shard_slice_create", "svo_shard_slice_attach", "svo_shard_slice_upload", "svo_shard_slice_begin", "svo_shard_slice_append",
    "svo_shard_slice_publish", "svo_shard_slice_fence", "svo_partition_voxel_counts",
    "svo_shard_exchange", "svo_shard_layout_from_table",
    "svo_run", "svo_get_stats", "svo_synchronize", "svo_host_alloc", "svo_host_free",
]


class SvoError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("svo_b200 error %d: %s" % (code, msg))
        self.code = code


class Params(C.Structure):
    _fields_ = [("gridsize", C.c_uint64), ("memory_limit_mb", C.c_uint64),
                ("bbox_min0", C.c_float), ("bbox_max0", C.c_float),
                ("payload", C.c_int32), ("generate_levels", C.c_int32), ("color_mode", C.c_int32),
                ("sparseness_limit", C.c_float), ("separability", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("n_partitions", C.c_uint64), ("n_pairs", C.c_uint64), ("n_voxels", C.c_uint64),
                ("n_nodes", C.c_uint64), ("n_data", C.c_uint64),
                ("n_small", C.c_uint64), ("n_medium", C.c_uint64), ("n_large", C.c_uint64),
                ("ms_upload", C.c_float), ("ms_partition", C.c_float), ("ms_voxelize", C.c_float),
                ("ms_build", C.c_float), ("ms_emit", C.c_float), ("ms_clear", C.c_float), ("ms_download", C.c_float),
                ("ms_vox_small", C.c_float), ("ms_emit_leaf", C.c_float), ("ms_compact", C.c_float),
                ("ms_dispatch", C.c_float), ("ms_peer_wait", C.c_float),
                ("kernel_launches", C.c_uint32), ("speculative", C.c_uint32),
                ("n_bricks", C.c_uint64), ("n_tiles1", C.c_uint64), ("n_brick_records", C.c_uint64)]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "reserved"}


class ShardLayout(C.Structure):
    _fields_ = [("n_voxels", C.c_uint64), ("n_nodes", C.c_uint64), ("node_lo", C.c_uint64), ("node_hi", C.c_uint64),
                ("leaf_offset", C.c_uint64), ("n_voxels_local", C.c_uint64), ("n_upper_records", C.c_uint64)]


_lib = None


def load_library(path: str | None = None):
    """Loads libsvo_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise SvoError(-1, "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a). There is no CPU fallback." % path)
    L = C.CDLL(path)
    vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
    L.svo_ctx_create.restype = i32; L.svo_ctx_create.argtypes = [i32, C.POINTER(vp)]
    L.svo_ctx_destroy.restype = None; L.svo_ctx_destroy.argtypes = [vp]
    L.svo_ctx_set_stream.restype = i32; L.svo_ctx_set_stream.argtypes = [vp, vp]
    L.svo_last_error.restype = C.c_char_p; L.svo_last_error.argtypes = [vp]
    L.svo_version.restype = C.c_char_p; L.svo_version.argtypes = []
    L.svo_estimate_partitions.restype = u64; L.svo_estimate_partitions.argtypes = [u64, u64]
    L.svo_text_roundtrip_float.restype = C.c_float; L.svo_text_roundtrip_float.argtypes = [C.c_float]
    L.svo_set_triangles.restype = i32; L.svo_set_triangles.argtypes = [vp, vp, u64, i32]
    L.svo_set_triangles_device.restype = i32; L.svo_set_triangles_device.argtypes = [vp, vp, u64, i32]
    L.svo_triangles_begin.restype = i32; L.svo_triangles_begin.argtypes = [vp, u64, i32]
    L.svo_triangles_append.restype = i32; L.svo_triangles_append.argtypes = [vp, vp, u64]
    L.svo_partition.restype = i32; L.svo_partition.argtypes = [vp, C.POINTER(Params), C.POINTER(u64), vp, u64]
    L.svo_voxelize.restype = i32; L.svo_voxelize.argtypes = [vp]
    L.svo_build.restype = i32; L.svo_build.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
    L.svo_fetch_nodes.restype = i32; L.svo_fetch_nodes.argtypes = [vp, u64, u64, vp]
    L.svo_fetch_data.restype = i32; L.svo_fetch_data.argtypes = [vp, u64, u64, vp]
    L.svo_device_nodes.restype = i32; L.svo_device_nodes.argtypes = [vp, C.POINTER(vp), C.POINTER(u64)]
    L.svo_device_data.restype = i32; L.svo_device_data.argtypes = [vp, C.POINTER(vp), C.POINTER(u64)]
    L.svo_fetch_voxel_codes.restype = i32; L.svo_fetch_voxel_codes.argtypes = [vp, vp, u64, C.POINTER(u64)]
    L.svo_shard_configure.restype = i32; L.svo_shard_configure.argtypes = [vp, i32, i32]
    L.svo_shard_table_size.restype = i32; L.svo_shard_table_size.argtypes = [vp, C.POINTER(u64)]
    L.svo_shard_count.restype = i32; L.svo_shard_count.argtypes = [vp, vp]
    L.svo_shard_emit.restype = i32; L.svo_shard_emit.argtypes = [vp, vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
    L.svo_shard_ranges.restype = i32; L.svo_shard_ranges.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
    L.svo_shard_dispatch_create.restype = i32; L.svo_shard_dispatch_create.argtypes = [vp, u64, i32, C.POINTER(vp), C.POINTER(vp)]
    L.svo_shard_dispatch_attach.restype = i32; L.svo_shard_dispatch_attach.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.svo_shard_dispatch_count.restype = i32; L.svo_shard_dispatch_count.argtypes = [vp, C.POINTER(Params), vp, u64, i32]
    L.svo_shard_dispatch_send.restype = i32; L.svo_shard_dispatch_send.argtypes = [vp]
    L.svo_shard_dispatch_finish.restype = i32; L.svo_shard_dispatch_finish.argtypes = [vp, C.POINTER(u64)]
    L.svo_shard_slice_create.restype = i32; L.svo_shard_slice_create.argtypes = [vp, u64, i32, C.POINTER(vp)]
    L.svo_shard_slice_attach.restype = i32; L.svo_shard_slice_attach.argtypes = [vp, C.POINTER(vp)]
    L.svo_shard_exchange.restype = i32; L.svo_shard_exchange.argtypes = [vp, vp]
    L.svo_shard_layout_from_table.restype = i32
    L.svo_shard_layout_from_table.argtypes = [C.POINTER(Params), i32, i32, vp, u64, C.POINTER(ShardLayout), vp, vp, u64]
    L.svo_shard_slice_upload.restype = i32; L.svo_shard_slice_upload.argtypes = [vp, vp, u64]
    L.svo_shard_slice_begin.restype = i32; L.svo_shard_slice_begin.argtypes = [vp, u64]
    L.svo_shard_slice_append.restype = i32; L.svo_shard_slice_append.argtypes = [vp, vp, u64]
    L.svo_partition_voxel_counts.restype = i32; L.svo_partition_voxel_counts.argtypes = [vp, vp, u64]
    L.svo_shard_slice_publish.restype = i32; L.svo_shard_slice_publish.argtypes = [vp, C.POINTER(Params), u64]
    L.svo_shard_slice_fence.restype = i32; L.svo_shard_slice_fence.argtypes = [vp]
    L.svo_ipc_export.restype = i32; L.svo_ipc_export.argtypes = [vp, vp]
    L.svo_ipc_open.restype = i32; L.svo_ipc_open.argtypes = [vp, C.POINTER(vp)]
    L.svo_ipc_close.restype = i32; L.svo_ipc_close.argtypes = [vp]
    L.svo_run.restype = i32; L.svo_run.argtypes = [vp, C.POINTER(Params), vp, u64, vp, u64, vp, u64, C.POINTER(Stats)]
    L.svo_get_stats.restype = i32; L.svo_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.svo_synchronize.restype = i32; L.svo_synchronize.argtypes = [vp]
    L.svo_host_alloc.restype = vp; L.svo_host_alloc.argtypes = [C.c_size_t]
    L.svo_host_free.restype = None; L.svo_host_free.argtypes = [vp]
    _lib = L
    return L


def ipc_export(dev_ptr: int) -> bytes:
    """cudaIpcGetMemHandle of a cudaMalloc'd pointer, as 64 bytes."""
    buf = C.create_string_buffer(64)
    rc = load_library().svo_ipc_export(dev_ptr, buf)
    if rc != 0:
        raise SvoError(rc, "cudaIpcGetMemHandle failed")
    return buf.raw


def ipc_open(handle: bytes) -> int:
    """Maps a peer process's allocation into this process (cudaIpcOpenMemHandle, enables peer access)."""
    p = C.c_void_p()
    rc = load_library().svo_ipc_open(C.create_string_buffer(handle, 64), C.byref(p))
    if rc != 0 or not p.value:
        raise SvoError(rc, "cudaIpcOpenMemHandle failed")
    return p.value


def ipc_close(dev_ptr: int) -> None:
    load_library().svo_ipc_close(dev_ptr)


def shard_layout_from_table(params: Params, rank: int, world: int, table: np.ndarray):
    """svo_shard_layout_from_table: the host-side merge of the sharded build as a pure function (no GPU).
    Returns (layout dict, rec_pos uint64[n], rec_words uint64[n, 3])."""
    L = load_library()
    table = np.ascontiguousarray(table).view(np.uint64)
    out = ShardLayout()
    rc = L.svo_shard_layout_from_table(C.byref(params), rank, world, table.ctypes.data, table.size, C.byref(out), None, None, 0)
    if rc != 0:
        raise SvoError(rc, (L.svo_last_error(None) or b"").decode())
    n = out.n_upper_records
    pos = np.zeros(max(n, 1), dtype=np.uint64)
    words = np.zeros((max(n, 1), 3), dtype=np.uint64)
    rc = L.svo_shard_layout_from_table(C.byref(params), rank, world, table.ctypes.data, table.size, C.byref(out), pos.ctypes.data, words.ctypes.data, n)
    if rc != 0:
        raise SvoError(rc, (L.svo_last_error(None) or b"").decode())
    return {k: getattr(out, k) for k, _ in ShardLayout._fields_}, pos[:n], words[:n]


def estimate_partitions(gridsize: int, memory_limit_mb: int) -> int:
    """partitioner.cpp:12-28 (host arithmetic inside the library)."""
    return int(load_library().svo_estimate_partitions(gridsize, memory_limit_mb))


def header_bytes(gridsize: int, n_nodes: int, n_data: int) -> bytes:
    """The `.octree` text header, octree_io.h:74-83."""
    return ("#octreeheader 1\ngridlength %d\nn_nodes %d\nn_data %d\nEND\n" % (gridsize, n_nodes, n_data)).encode()


@dataclass
class Octree:
    """In-memory image of `.octree` / `.octreenodes` / `.octreedata`."""
    header: bytes
    nodes: np.ndarray      # uint8, n_nodes * 24
    data: np.ndarray       # uint8, n_data * 32
    n_partitions: int
    n_voxels: int
    stats: dict

    @property
    def n_nodes(self) -> int:
        return self.nodes.size // NODE_BYTES

    @property
    def n_data(self) -> int:
        return self.data.size // DATA_BYTES

    def write(self, base: str) -> None:
        """Writes <base>.octree/.octreenodes/.octreedata (base includes `<g>_<P>`)."""
        with open(base + ".octree", "wb") as f:
            f.write(self.header)
        self.nodes.tofile(base + ".octreenodes")
        self.data.tofile(base + ".octreedata")


class PinnedBuffer:
    """cudaHostAlloc'd byte buffer exposed as a numpy array."""

    def __init__(self, nbytes: int):
        self._lib = load_library()
        self.nbytes = int(nbytes)
        self.ptr = self._lib.svo_host_alloc(max(self.nbytes, 1))
        if not self.ptr:
            raise SvoError(3, "cudaHostAlloc(%d) failed" % nbytes)
        self.array = np.ctypeslib.as_array((C.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr))[: self.nbytes]

    def free(self):
        if self.ptr:
            self._lib.svo_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class SvoBuilder:
    """One context = one GPU.  Methods follow the reference's stages."""

    def __init__(self, device: int = 0):
        self._lib = load_library()
        h = C.c_void_p()
        rc = self._lib.svo_ctx_create(device, C.byref(h))
        if rc != 0:
            raise SvoError(rc, (self._lib.svo_last_error(None) or b"").decode())
        self._h = h
        self.device = device
        self._keep = None
        self.params: Params | None = None

    # -- plumbing ---------------------------------------------------------
    def _ck(self, rc: int):
        if rc != 0:
            raise SvoError(rc, (self._lib.svo_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.svo_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream: int | None) -> None:
        """Issue all work on the given cudaStream_t handle (e.g. torch.cuda.Stream().cuda_stream)."""
        self._ck(self._lib.svo_ctx_set_stream(self._h, cuda_stream))

    @staticmethod
    def make_params(length: float, gridsize: int, payload: bool, memory_limit_mb: int = 2048,
                    levels: bool = False, color: str = "model", bbox_min0: float = 0.0, separability: int = 26) -> Params:
        return Params(gridsize, memory_limit_mb, np.float32(bbox_min0), np.float32(bbox_min0) + np.float32(length),
                      int(payload), int(levels), COLOR_MODES[color], 0.10, int(separability))

    # -- stages -----------------------------------------------------------
    def set_triangles(self, tris) -> None:
        """tris: numpy (T, 9|21) float32 host array, or a torch CUDA tensor of that shape (borrowed)."""
        if isinstance(tris, np.ndarray):
            tris = np.ascontiguousarray(tris, dtype=np.float32)
            self._keep = tris
            self._ck(self._lib.svo_set_triangles(self._h, tris.ctypes.data, tris.shape[0], tris.shape[1]))
        else:  # torch tensor
            assert tris.is_cuda and tris.is_contiguous() and tris.dtype.is_floating_point and tris.element_size() == 4
            self._keep = tris
            self._ck(self._lib.svo_set_triangles_device(self._h, tris.data_ptr(), tris.shape[0], tris.shape[1]))

    def set_triangles_streamed(self, tris: np.ndarray, chunk_tris: int) -> None:
        """svo_triangles_begin / _append with two alternating pinned chunks (what the CLI does while it reads the file)."""
        tris = np.ascontiguousarray(tris, dtype=np.float32)
        T, fpt = tris.shape
        self._ck(self._lib.svo_triangles_begin(self._h, T, fpt))
        bufs = [PinnedBuffer(max(chunk_tris, 1) * fpt * 4) for _ in range(2)]
        flat = tris.reshape(-1).view(np.uint8)
        for k, lo in enumerate(range(0, T, max(chunk_tris, 1))):
            n = min(chunk_tris, T - lo)
            b = bufs[k & 1]
            b.array[: n * fpt * 4] = flat[lo * fpt * 4: (lo + n) * fpt * 4]
            self._ck(self._lib.svo_triangles_append(self._h, b.ptr, n))
        self.synchronize()
        for b in bufs:
            b.free()

    def partition(self, params: Params, want_counts: bool = True) -> np.ndarray | None:
        """partitioner.cpp:101-149 → per-partition triangle counts (the .trip header values).
        want_counts=False skips the counting pass: the voxelizer enumerates partitions inline."""
        self.params = params
        P = estimate_partitions(params.gridsize, params.memory_limit_mb)
        n = C.c_uint64()
        if not want_counts:
            self._ck(self._lib.svo_partition(self._h, C.byref(params), C.byref(n), None, 0))
            return None
        counts = np.zeros(P, dtype=np.uint64)
        self._ck(self._lib.svo_partition(self._h, C.byref(params), C.byref(n), counts.ctypes.data, P))
        assert n.value == P
        return counts

    def voxelize(self) -> None:
        self._ck(self._lib.svo_voxelize(self._h))

    def build(self) -> tuple[int, int, int]:
        nv, nn, nd = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._ck(self._lib.svo_build(self._h, C.byref(nv), C.byref(nn), C.byref(nd)))
        return nv.value, nn.value, nd.value

    # -- sharded (multi-GPU) build: see include/svo_b200.h -------------------
    def shard_configure(self, rank: int, world: int) -> None:
        self._ck(self._lib.svo_shard_configure(self._h, rank, world))

    def shard_table_size(self) -> int:
        n = C.c_uint64()
        self._ck(self._lib.svo_shard_table_size(self._h, C.byref(n)))
        return n.value

    def shard_count(self, dev_table_ptr: int) -> None:
        self._ck(self._lib.svo_shard_count(self._h, dev_table_ptr))

    def shard_emit(self, dev_table_ptr: int) -> tuple[int, int, int]:
        nv, nn, nd = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._ck(self._lib.svo_shard_emit(self._h, dev_table_ptr, C.byref(nv), C.byref(nn), C.byref(nd)))
        return nv.value, nn.value, nd.value

    def shard_ranges(self) -> tuple[int, int, int, int]:
        a, b, c_, d = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._ck(self._lib.svo_shard_ranges(self._h, C.byref(a), C.byref(b), C.byref(c_), C.byref(d)))
        return a.value, b.value, c_.value, d.value

    # -- remote staging of triangle slices (multi-GPU, default input path) ------
    def slice_create(self, capacity_tris: int, fpt: int) -> int:
        """Allocates this rank's window (control block, exchange table, block lists, slice); returns its device
        pointer (to share with the peers)."""
        a = C.c_void_p()
        self._ck(self._lib.svo_shard_slice_create(self._h, capacity_tris, fpt, C.byref(a)))
        return a.value

    def slice_attach(self, windows: list[int]) -> None:
        n = len(windows)
        self._ck(self._lib.svo_shard_slice_attach(self._h, (C.c_void_p * n)(*windows)))

    def shard_exchange(self, dev_table_ptr: int) -> None:
        """The table exchange over peer memory (instead of the caller's all-reduce); needs the slice windows."""
        self._ck(self._lib.svo_shard_exchange(self._h, dev_table_ptr))

    def slice_upload(self, tris) -> None:
        """This rank's slice of the triangle file: numpy host array (pageable or pinned) or torch CUDA tensor, (n, 9|21) float32."""
        if isinstance(tris, np.ndarray):
            tris = np.ascontiguousarray(tris, dtype=np.float32)
            ptr = tris.ctypes.data
        else:
            assert tris.is_contiguous() and tris.element_size() == 4
            ptr = tris.data_ptr()
        self._keep_local = tris
        self._ck(self._lib.svo_shard_slice_upload(self._h, ptr, tris.shape[0]))

    def slice_publish(self, params: Params, n_total: int) -> None:
        self.params = params
        self._ck(self._lib.svo_shard_slice_publish(self._h, C.byref(params), n_total))

    def slice_fence(self) -> None:
        self._ck(self._lib.svo_shard_slice_fence(self._h))

    # -- triangle dispatch over peer memory (multi-GPU) -------------------------
    def dispatch_create(self, capacity_tris: int, fpt: int) -> tuple[int, int]:
        """Allocates this rank's inbox + control block; returns their device pointers (to share with the peers)."""
        a, b = C.c_void_p(), C.c_void_p()
        self._ck(self._lib.svo_shard_dispatch_create(self._h, capacity_tris, fpt, C.byref(a), C.byref(b)))
        return a.value, b.value

    def dispatch_attach(self, inbox_ptrs: list[int], ctrl_ptrs: list[int]) -> None:
        n = len(inbox_ptrs)
        A = (C.c_void_p * n)(*inbox_ptrs)
        B = (C.c_void_p * n)(*ctrl_ptrs)
        self._ck(self._lib.svo_shard_dispatch_attach(self._h, A, B))

    def dispatch_count(self, params: Params, local_tris) -> None:
        """local_tris: torch CUDA tensor (n_local, 9|21) float32 -- this rank's slice of the file, in file order."""
        assert local_tris.is_cuda and local_tris.is_contiguous() and local_tris.element_size() == 4
        self._keep_local = local_tris
        self.params = params
        self._ck(self._lib.svo_shard_dispatch_count(self._h, C.byref(params), local_tris.data_ptr(), local_tris.shape[0], local_tris.shape[1]))

    def dispatch_send(self) -> None:
        self._ck(self._lib.svo_shard_dispatch_send(self._h))

    def dispatch_finish(self) -> int:
        n = C.c_uint64()
        self._ck(self._lib.svo_shard_dispatch_finish(self._h, C.byref(n)))
        return n.value

    def fetch_nodes(self, first: int, count: int, out=None):
        """Records [first, first + count) of the node file into `out`: a numpy uint8 array (host) or a torch CUDA
        uint8 tensor on this context's device (device-to-device copy)."""
        if out is None:
            out = np.empty(count * NODE_BYTES, dtype=np.uint8)
        ptr = out.ctypes.data if isinstance(out, np.ndarray) else out.data_ptr()
        self._ck(self._lib.svo_fetch_nodes(self._h, first, count, ptr))
        return out

    def fetch_data(self, first: int, count: int, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty(count * DATA_BYTES, dtype=np.uint8)
        self._ck(self._lib.svo_fetch_data(self._h, first, count, out.ctypes.data))
        return out

    def voxel_codes(self) -> np.ndarray:
        """Ascending Morton codes of the filled voxels (what main.cpp:355-368 feeds addVoxel)."""
        st = self.stats()
        out = np.empty(st["n_voxels"], dtype=np.uint64)
        n = C.c_uint64()
        self._ck(self._lib.svo_fetch_voxel_codes(self._h, out.ctypes.data, out.size, C.byref(n)))
        return out[: n.value]

    def partition_voxel_counts(self) -> np.ndarray:
        """Voxels per logical partition (the reference's per-partition `found N new voxels`, main.cpp:348)."""
        P = estimate_partitions(self.params.gridsize, self.params.memory_limit_mb)
        out = np.zeros(P, dtype=np.uint64)
        self._ck(self._lib.svo_partition_voxel_counts(self._h, out.ctypes.data, P))
        return out

    def device_nodes(self) -> tuple[int, int]:
        p, n = C.c_void_p(), C.c_uint64()
        self._ck(self._lib.svo_device_nodes(self._h, C.byref(p), C.byref(n)))
        return p.value or 0, n.value

    def stats(self) -> dict:
        s = Stats()
        self._ck(self._lib.svo_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def synchronize(self) -> None:
        self._ck(self._lib.svo_synchronize(self._h))

    # -- whole path -------------------------------------------------------
    def run(self, tris, length: float, gridsize: int, memory_limit_mb: int = 2048,
            levels: bool = False, color: str = "model", fetch: bool = True, separability: int = 26) -> Octree:
        """main.cpp:298-389: partition → voxelize → build (→ fetch)."""
        payload = tris.shape[1] == 21
        prm = self.make_params(length, gridsize, payload, memory_limit_mb, levels, color, separability=separability)
        self.set_triangles(tris)
        self.partition(prm, want_counts=False)
        self.voxelize()
        nv, nn, nd = self.build()
        if fetch:
            nodes = self.fetch_nodes(0, nn)
            data = self.fetch_data(0, nd)
        else:
            nodes = np.empty(0, dtype=np.uint8)
            data = np.empty(0, dtype=np.uint8)
        st = self.stats()
        return Octree(header_bytes(gridsize, nn, nd), nodes, data, st["n_partitions"], nv, st)

    def run_host(self, params: Params, tris: np.ndarray, nodes_out: np.ndarray, data_out: np.ndarray) -> dict:
        """svo_run: one C-ABI call, host buffers in and out."""
        s = Stats()
        self._ck(self._lib.svo_run(self._h, C.byref(params), tris.ctypes.data, tris.shape[0],
                                   nodes_out.ctypes.data, nodes_out.size // NODE_BYTES,
                                   data_out.ctypes.data, data_out.size // DATA_BYTES, C.byref(s)))
        return s.as_dict()
