"""Multi-GPU driver of the sharded build (one context per GPU) and its host-side plan.

The data path has exactly one exchange: the table of top-of-shard subtree records
(`svo_shard_count` -> all-reduce(sum) -> `svo_shard_emit`, include/svo_b200.h).
With torch.distributed the all-reduce runs over NCCL / NVLink on the device
buffer the library filled; nothing else crosses GPUs.  Every rank ends up with a
contiguous range of the .octreenodes / .octreedata files; the ranges tile them.

`plan()` restates the library's shard geometry in Python so that the host logic
can be tested without a GPU (tests/test_sharded_cpu.py, gloo, world_size 2).
"""
from __future__ import annotations

import math
import os
import time
from dataclasses import dataclass

import numpy as np

from .api import SvoBuilder, SvoError, header_bytes, estimate_partitions, NODE_BYTES, DATA_BYTES


@dataclass
class ShardPlan:
    world: int
    rank: int
    depth: int            # D = log2(gridsize)
    k: int                # log8(P)
    dc: int               # chunk depth
    top_local_level: int  # J
    chunk_range: tuple    # [c0, c1) chunks of 8^(D-dc) voxels
    morton_range: tuple   # [start, end) Morton codes owned
    partition_range: tuple  # [p_first, p_last] logical partitions that intersect the slab
    table_entries: int    # WJ: global level-J words = table entries (x4 u64)
    own_entries: tuple    # [wj0, wj1) entries this rank writes


def plan(gridsize: int, n_partitions: int, world: int, rank: int) -> ShardPlan:
    """Mirror of setup_geometry() in csrc/svo_api.cu."""
    D = int(math.log2(gridsize))
    k = int(round(math.log(n_partitions, 8))) if n_partitions > 1 else 0
    dc = 0
    if world > 1:
        need = 0
        while 8 ** need < world:
            need += 1
        dc = max(k, need)
        if D - dc < 2:
            raise ValueError("gridsize too small for this many shards")
    J = max((D - dc) // 2 - 1, 0)
    nchunks = 8 ** dc
    c0, c1 = nchunks * rank // world, nchunks * (rank + 1) // world
    bits = 3 * (D - dc)
    start, end = c0 << bits, c1 << bits
    sh = 3 * (dc - k)
    shJ = 6 * (J + 1)
    WJ = (1 << (3 * D - shJ)) if 3 * D >= shJ else 1
    return ShardPlan(world, rank, D, k, dc, J, (c0, c1), (start, end), (c0 >> sh, (c1 - 1) >> sh), WJ,
                     (start >> shJ, max(end >> shJ, (start >> shJ) + 1)))


def merge_tables(tables: list[np.ndarray]) -> np.ndarray:
    """What the all-reduce computes: the entries of different ranks are disjoint, so sum == union."""
    out = np.zeros_like(tables[0])
    for t in tables:
        assert not ((out != 0) & (t != 0)).any(), "two ranks wrote the same table entry"
        out += t
    return out


@dataclass
class ShardResult:
    rank: int
    n_voxels: int
    n_nodes: int
    n_data: int
    node_range: tuple
    data_range: tuple
    nodes: np.ndarray      # uint8, this rank's node records
    data: np.ndarray       # uint8, this rank's data records
    stats: dict


def assemble(results: list[ShardResult], gridsize: int):
    """Concatenate the per-rank ranges into the complete file images (host side, tests / rank-0 writer)."""
    rs = sorted(results, key=lambda r: r.rank)
    nn, nd = rs[0].n_nodes, rs[0].n_data
    pos = 0
    for r in rs:
        assert r.node_range[0] == pos or r.node_range[0] == r.node_range[1], (r.rank, r.node_range, pos)
        pos = max(pos, r.node_range[1])
    assert pos == nn, (pos, nn)
    nodes = np.zeros(nn * NODE_BYTES, dtype=np.uint8)
    data = np.zeros(nd * DATA_BYTES, dtype=np.uint8)
    for r in rs:
        nodes[r.node_range[0] * NODE_BYTES: r.node_range[1] * NODE_BYTES] = r.nodes
        data[r.data_range[0] * DATA_BYTES: r.data_range[1] * DATA_BYTES] = r.data
    return header_bytes(gridsize, nn, nd), nodes, data


def slice_bounds(n_tris: int, world: int, rank: int) -> tuple[int, int]:
    """Rank `rank`'s contiguous slice [lo, hi) of the triangle file (file order, equal sizes up to rounding)."""
    per = (n_tris + world - 1) // world
    return min(rank * per, n_tris), min((rank + 1) * per, n_tris)


def run_single_process(tris, length: float, gridsize: int, world: int, memory_limit_mb: int = 2048,
                       color: str = "model", device: int = 0, fetch: bool = True, dispatch: bool = False,
                       slices: list | None = None, remote: bool = False, levels: bool = False) -> list[ShardResult]:
    """All `world` ranks as contexts of ONE process (sharing a GPU is fine): the table exchange is a
    host-side sum. Used by the single-GPU parity tests of the sharded path and by the CLI.
    dispatch=True: every rank starts with only its slice of the file and the triangle dispatch
    (svo_shard_dispatch_*) routes the records through the peers' inboxes -- same kernels as across GPUs,
    the peer pointers just happen to be on one device. `slices` overrides the equal split ([(lo, hi)] per rank).
    remote=True: remote staging (svo_shard_slice_*): every rank keeps only its slice, the voxelizers read the blocks
    they need from the owners' buffers."""
    import torch
    payload = tris.shape[1] == 21
    ctxs = [SvoBuilder(device) for _ in range(world)]
    try:
        prm = SvoBuilder.make_params(length, gridsize, payload, memory_limit_mb, levels, color)
        tables = []
        if dispatch:
            T = tris.shape[0]
            for r, sb in enumerate(ctxs):
                sb.shard_configure(r, world)
            ptrs = [sb.dispatch_create(T, tris.shape[1]) for sb in ctxs]
            for sb in ctxs:
                sb.dispatch_attach([p[0] for p in ptrs], [p[1] for p in ptrs])
            bounds = slices or [slice_bounds(T, world, r) for r in range(world)]
            local = [torch.from_numpy(np.ascontiguousarray(tris[lo:hi])).to("cuda:%d" % device) for lo, hi in bounds]
            for sb, l in zip(ctxs, local):          # each phase for all ranks before the next (one host thread)
                sb.dispatch_count(prm, l)
            for sb in ctxs:
                sb.dispatch_send()
            for sb in ctxs:
                sb.dispatch_finish()
        if remote:
            T = tris.shape[0]
            bounds = slices or [slice_bounds(T, world, r) for r in range(world)]
            cap = max(max(hi - lo for lo, hi in bounds), 1)
            for r, sb in enumerate(ctxs):
                sb.shard_configure(r, world)
            wins = [sb.slice_create(cap, tris.shape[1]) for sb in ctxs]
            for sb in ctxs:
                sb.slice_attach(wins)
            for sb, (lo, hi) in zip(ctxs, bounds):
                sb.slice_upload(tris[lo:hi])
            for sb in ctxs:                         # each phase for all ranks before the next (one host thread)
                sb.slice_publish(prm, T)
        for r, sb in enumerate(ctxs):
            if not dispatch and not remote:
                sb.shard_configure(r, world)
                sb.set_triangles(tris)
            sb.partition(prm, want_counts=not (dispatch or remote))
            sb.voxelize()
            tables.append(torch.zeros(sb.shard_table_size(), dtype=torch.int64, device="cuda:%d" % device))
        return shard_build_single_process(ctxs, tables, remote, fetch)
    finally:
        for sb in ctxs:
            sb.close()


def shard_build_single_process(ctxs, tables, remote: bool, fetch: bool = True) -> list[ShardResult]:
    """svo_shard_count -> exchange -> svo_shard_emit for all ranks of one process (contexts already voxelized).
    Repeats the three calls for ALL ranks when the library answers SVO_E_RETRY (include/svo_b200.h)."""
    import torch
    for attempt in range(4):
        for sb, t in zip(ctxs, tables):
            sb.shard_count(t.data_ptr())
            if not remote:
                sb.synchronize()
        if remote:
            # one host thread drives all ranks: issue every rank's exchange before waiting on any of them (the
            # device-side waits resolve once all ranks have pushed; nothing in between may synchronize the device)
            for sb, t in zip(ctxs, tables):
                sb.shard_exchange(t.data_ptr())
        else:
            merged = torch.stack(tables).sum(dim=0)
            # disjointness is part of the contract
            assert int((torch.stack([(t != 0).to(torch.int64) for t in tables]).sum(dim=0) > 1).sum()) == 0
        out, retry = [], 0
        for r, sb in enumerate(ctxs):
            try:
                nv, nn, nd = sb.shard_emit(tables[r].data_ptr() if remote else merged.data_ptr())
            except SvoError as e:
                if e.code != 5:
                    raise
                retry += 1
                continue
            nlo, nhi, dlo, dhi = sb.shard_ranges()
            nodes = sb.fetch_nodes(nlo, nhi - nlo) if fetch else np.empty(0, np.uint8)
            data = sb.fetch_data(dlo, dhi - dlo) if fetch else np.empty(0, np.uint8)
            out.append(ShardResult(r, nv, nn, nd, (nlo, nhi), (dlo, dhi), nodes, data, sb.stats()))
        if retry == 0:
            if remote:
                assert all(bool((t == tables[0]).all()) for t in tables), "the peer-memory exchange left different tables on the ranks"
            return out
        assert retry == len(ctxs), "SVO_E_RETRY must be answered by every rank in the same step (%d of %d)" % (retry, len(ctxs))
    raise RuntimeError("sharded build did not settle after 4 attempts")



class DistributedBuilder:
    """One rank of a torch.distributed job (backend nccl, one process per GPU)."""

    def __init__(self, dist, device: int):
        import torch
        self.dist = dist
        self.torch = torch
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()
        self.sb = SvoBuilder(device)
        self.sb.shard_configure(self.rank, self.world)
        self.device = device
        self.table = None
        self.stream = None

    def set_stream(self, stream):
        """Run the library's kernels AND the table all-reduce on this torch stream (one queue, no race)."""
        self.stream = stream
        self.sb.set_stream(stream.cuda_stream)

    def set_triangles(self, tris):
        self.sb.set_triangles(tris)
        self.local = None
        self.sliced = False

    def enable_dispatch(self, capacity_tris: int, fpt: int):
        """Allocates the inbox, exchanges CUDA IPC handles with the peers (once) and maps their inboxes:
        afterwards triangle records travel GPU -> GPU as NVLink stores issued by our own kernel. Raises on every
        rank if any rank fails."""
        from .api import ipc_export, ipc_open
        inbox = ctrl = mine = err = None
        try:
            inbox, ctrl = self.sb.dispatch_create(capacity_tris, fpt)
            mine = (ipc_export(inbox), ipc_export(ctrl))
        except Exception as e:      # noqa: BLE001
            err = e
        handles = [None] * self.world
        self.dist.all_gather_object(handles, mine)
        ib, cb = [], []
        self._opened = getattr(self, "_opened", [])
        if err is None and all(h is not None for h in handles):
            try:
                for r, (hi, hc) in enumerate(handles):
                    if r == self.rank:
                        ib.append(inbox); cb.append(ctrl)
                    else:
                        a, b = ipc_open(hi), ipc_open(hc)
                        self._opened += [a, b]
                        ib.append(a); cb.append(b)
                self.sb.dispatch_attach(ib, cb)
            except Exception as e:      # noqa: BLE001
                err = e
        elif err is None:
            err = RuntimeError("a peer rank could not create its inbox")
        if not self._agree(err is None):
            raise RuntimeError("triangle dispatch unavailable: %s" % (err if err is not None else "a peer rank failed"))
        self.dist.barrier()

    def _agree(self, ok: bool) -> bool:
        """True iff every rank reports ok (one tiny all-reduce; keeps the ranks in step when one of them fails)."""
        t = self.torch.tensor([1 if ok else 0], dtype=self.torch.int32, device="cuda:%d" % self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(int(t))

    def enable_slices(self, capacity_tris: int, fpt: int, n_total: int):
        """Remote staging: allocates this rank's window, exchanges CUDA IPC handles with the peers (once) and maps
        theirs. Afterwards the voxelizer stages triangle blocks straight from the owners' HBM. Raises on EVERY rank if
        any rank cannot allocate or map (so that the caller can fall back collectively)."""
        from .api import ipc_export, ipc_open
        mine, handle, err = None, None, None
        try:
            mine = self.sb.slice_create(capacity_tris, fpt)
            handle = ipc_export(mine)
        except Exception as e:      # noqa: BLE001
            err = e
        handles = [None] * self.world
        self.dist.all_gather_object(handles, handle)
        wins = []
        self._opened = getattr(self, "_opened", [])
        if err is None and all(h is not None for h in handles):
            try:
                for r, h in enumerate(handles):
                    if r == self.rank:
                        wins.append(mine)
                    else:
                        p = ipc_open(h)
                        self._opened.append(p)
                        wins.append(p)
                self.sb.slice_attach(wins)
            except Exception as e:      # noqa: BLE001
                err = e
        elif err is None:
            err = RuntimeError("a peer rank could not create its window")
        if not self._agree(err is None):
            raise RuntimeError("remote staging unavailable: %s" % (err if err is not None else "a peer rank failed"))
        self.n_total = n_total
        self.sliced = True
        self.dist.barrier()

    def upload_slice(self, tris):
        """This rank's slice (numpy host array or torch CUDA tensor) -> the mapped slice buffer, stream ordered."""
        self.sb.slice_upload(tris)

    def set_local_triangles(self, local):
        """This rank's slice of the triangle file (torch CUDA tensor, file order across ranks)."""
        self.local = local

    def step(self, prm):
        """[dispatch ->] partition -> voxelize -> local build -> NCCL all-reduce of the table -> merged emit."""
        sb = self.sb
        if getattr(self, "sliced", False):
            sb.slice_publish(prm, self.n_total)
        elif getattr(self, "local", None) is not None:
            sb.dispatch_count(prm, self.local)
            sb.dispatch_send()
            sb.dispatch_finish()
        sb.partition(prm, want_counts=False)
        sb.voxelize()
        n = sb.shard_table_size()
        if self.table is None or self.table.numel() != n:
            self.table = self.torch.zeros(n, dtype=self.torch.int64, device="cuda:%d" % self.device)
        self.retries = 0
        while True:
            sb.shard_count(self.table.data_ptr())
            if getattr(self, "sliced", False) and os.environ.get("SVO_TABLE_EXCHANGE", "peer") == "peer":
                sb.shard_exchange(self.table.data_ptr())     # our own exchange over the peer windows (no NCCL on the path)
            elif self.stream is not None:
                with self.torch.cuda.stream(self.stream):   # same queue as the kernels that filled / will read the table
                    self.dist.all_reduce(self.table)
            else:
                sb.synchronize()                            # the library runs on its own stream: order it against torch's
                self.dist.all_reduce(self.table)
                self.torch.cuda.current_stream().synchronize()
            try:
                return sb.shard_emit(self.table.data_ptr())      # sum over ranks == union of disjoint entries (NCCL over NVLink)
            except SvoError as e:
                # SVO_E_RETRY: some rank's speculative local build outgrew its lists; every rank sees it in the same step
                if e.code != 5 or self.retries >= 3:
                    raise
                self.retries += 1

    def close(self):
        self.sb.synchronize()
        if getattr(self, "_opened", None):
            from .api import ipc_close
            self.dist.barrier()                      # nobody unmaps while a peer may still be writing
            for p in self._opened:
                ipc_close(p)
            self._opened = []
        self.sb.close()


# ----------------------------------------------------------------------------
# numpy closed form of the file layout: used by the CPU tests of the
# exchange protocol and as an independent check of the node ordering.
# ----------------------------------------------------------------------------

def subtree_table_from_codes(codes: np.ndarray, gridsize: int, p: ShardPlan) -> np.ndarray:
    """Table entries {mask, S, leaves, 0} of the level-J subtrees inside p.morton_range, from sorted Morton codes."""
    D, J = p.depth, p.top_local_level
    shJ = 6 * (J + 1)
    t = np.zeros(p.table_entries * 4, dtype=np.int64)
    own = codes[(codes >= p.morton_range[0]) & (codes < p.morton_range[1])]
    if own.size == 0:
        return t
    root_depth = D - 2 * (J + 1)
    keys = own >> np.uint64(shJ)
    for key in np.unique(keys):
        sub = own[keys == key]
        S = 0
        for d in range(max(root_depth, 0) + 1, D + 1):
            S += np.unique(sub >> np.uint64(3 * (D - d))).size
        if root_depth < 0:       # virtual top word: depth -1 -> depth 0 (the root) is a child of the word
            S += 0
        bits = np.unique((sub >> np.uint64(shJ - 6)) & np.uint64(63))
        mask = 0
        for b in bits:
            mask |= 1 << int(b)
        e = int(key) * 4
        t[e] = np.array(mask, dtype=np.uint64).astype(np.int64)
        t[e + 1] = S
        t[e + 2] = sub.size
    return t


def node_count_from_codes(codes: np.ndarray, gridsize: int) -> int:
    """n_nodes of the reference's tree: distinct Morton prefixes over all depths (root included)."""
    D = int(math.log2(gridsize))
    if codes.size == 0:
        return 1
    return 1 + sum(np.unique(codes >> np.uint64(3 * (D - d))).size for d in range(1, D + 1))
