#!/usr/bin/env python
"""bench.py -- voxelize + SVO build throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one pass of the hot path (partition -> voxelize -> build, incl. the sparse clear that re-arms the
bit-grid) over the workload. N = 1 runs BASELINE.json configs[1]: `svo_builder_binary -s 1024` on a synthetic
2 M-triangle displaced sphere. N > 1 (torchrun, one rank per GPU) runs the sharded path: a (2*1024)^3 grid whose 8
logical partitions are 1024^3 each; N of the octants hold one displaced sphere each; every rank holds 1/N of the
triangle file, voxelizes (staging the records it needs from the owning GPU's HBM over NVLink) and builds the partitions
it owns, the subtree table is exchanged over peer memory and the shared upper levels are merged on the device (weak
scaling: per-GPU work fixed). After the timed loop every rank ALSO builds the whole mesh in a plain single-GPU context
and compares its range of the node file byte for byte on the device (`parity`), and -- unless SVO_BENCH_STRONG=0 --
the north-star strong-scaling record is taken: BASELINE.json configs[4] (8192^3, 100 M-triangle thin shell) on N GPUs
against one GPU, checked against the reference's golden file checksum (`strong`).

`value` is triangles/s with inputs resident in HBM; `e2e` is the same metric through the C ABI with HOST buffers
(H2D + D2H inside the timed region). The reference arm (--impl reference) times the unmodified reference CPU binary
from oracle/_ref (1 thread: the reference is single threaded) on the same workload (a bounded sample of it when one
run takes too long for K + W runs).
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "voxelize+SVO build throughput"
UNIT = "triangles/s"
GRID = 1024
SPHERE_N = 1000            # 1000 x 1000 quads -> 2,000,000 triangles
BENCH_GRID_SHARDED = 2048  # 8 logical partitions of 1024^3 (default -l 2048)
PROFILE_JSON = os.path.join(ROOT, "profiles", "r02_traffic_c2.json")     # ncu --set full capture of the same C2 step


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)", 1965.0


# ----------------------------------------------------------------------------
# workloads (shared by both arms: identical `config`)
# ----------------------------------------------------------------------------
def workload(n_gpus: int):
    """Returns (mesh, gridsize, workload string)."""
    from ooc_svo_builder_b200 import meshgen
    if n_gpus <= 1:
        return (meshgen.displaced_sphere(SPHERE_N, SPHERE_N, seed=1), GRID,
                "svo_builder_binary -s 1024, synthetic 2M-triangle displaced sphere (BASELINE.json configs[1])")
    # weak scaling: one displaced sphere per populated octant of a 2048^3 grid. File order: the i-th sphere of the file
    # lies in the octant that rank i+1 owns, so with every rank holding the i-th slice of the file ALL triangle records
    # cross NVLink (nothing is local by construction).
    octants = {2: [0, 4], 4: [0, 2, 4, 6], 8: list(range(8))}[n_gpus]
    base = meshgen.displaced_sphere(SPHERE_N, SPHERE_N, seed=1, length=1.0)
    parts = []
    for i in range(n_gpus):
        o = octants[(i + 1) % n_gpus]
        off = np.array([(o & 1), (o >> 1) & 1, (o >> 2) & 1], dtype=np.float32)
        parts.append((base.tris.reshape(-1, 3, 3) + off).reshape(-1, 9))
    tris = np.ascontiguousarray(np.concatenate(parts), dtype=np.float32)
    return (meshgen.Mesh(tris, 2.0), BENCH_GRID_SHARDED,
            "svo_builder_binary -s 2048 (8 logical partitions of 1024^3), one 2M-triangle displaced sphere in each of %d octants" % n_gpus)


def config_dict(wl: str, grid: int, n_triangles: int, n_voxels, n_nodes, partitions: int, n_gpus: int) -> dict:
    """The `config` object of the JSON line: the same dict from both arms (workload description only)."""
    return {"workload": wl, "gridsize": grid, "memory_limit_mb": 2048, "n_triangles": int(n_triangles),
            "n_voxels": None if n_voxels is None else int(n_voxels), "n_nodes": None if n_nodes is None else int(n_nodes),
            "partitions": int(partitions), "n_gpus": int(n_gpus)}


# ----------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ----------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, smmax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        allsm = []
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                allsm.append(float(f[1]))
                if t0 <= ts <= t1 + 0.2:
                    sm.append(float(f[1])); smmax.append(float(f[2]))
                    for i, n in enumerate(names):
                        if f[5 + i].lower().startswith("active"):
                            reasons.add(n)
            except ValueError:
                continue
        if not sm:
            sm = allsm
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smmax) if smmax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference CLI on the host CPU
# ----------------------------------------------------------------------------
def run_reference_cpu(mesh, gridsize: int, steps: int, warmup: int, budget_s: float = 200.0):
    """Times the reference's own CPU implementation of the path: one step = one full run of the reference CLI
    (partition + voxelize + build + its own file IO, page cache warm) on the workload -- or, when steps + warmup such
    runs would not fit `budget_s`, on a bounded SAMPLE of it (a prefix of the triangle file, same grid).
    Returns (triangles/s, voxels/s, info)."""
    from oracle import oracle as O
    from ooc_svo_builder_b200 import meshgen
    d = tempfile.mkdtemp(prefix="svo_bench_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        kind = "reference" if O.ref_available() else "port"
        sample = mesh
        note = "whole workload"

        def one(m):
            if kind == "reference":
                hdr = meshgen.write_tri(os.path.join(d, "m"), m)
                t = time.perf_counter()
                p = subprocess.run([O.ref_exe(False), "-f", hdr, "-s", str(gridsize)], capture_output=True, text=True)
                dt = time.perf_counter() - t
                nv = nn = None
                for line in p.stdout.splitlines():
                    if line.startswith("Total amount of voxels:"):
                        nv = int(line.split(":")[1])
                for fn in os.listdir(d):
                    if fn.endswith(".octree"):
                        for line in open(os.path.join(d, fn)):
                            if line.startswith("n_nodes"):
                                nn = int(line.split()[1])
                return dt, nv, nn
            t = time.perf_counter()
            r = O.build(m.tris, m.length, gridsize)
            return time.perf_counter() - t, r.n_voxels, r.n_nodes

        # probe run (counts as the first warm-up run): does the whole workload fit the budget?
        dt, nvox, nnodes = one(sample)
        whole_counts = (nvox, nnodes)
        runs_needed = steps + warmup
        if dt * runs_needed > budget_s and sample.n_triangles > 1000:
            keep = max(1000, int(sample.n_triangles * budget_s / (dt * runs_needed)))
            sample = meshgen.Mesh(np.ascontiguousarray(mesh.tris[:keep]), mesh.length)
            note = "bounded sample: the first %d of %d triangles of the file, same grid" % (keep, mesh.n_triangles)
            dt, nvox, nnodes = one(sample)
        times = []
        for i in range(1, runs_needed):
            dt, nvox, nnodes = one(sample)
            if i >= warmup:
                times.append(dt)
        if not times:
            times = [dt]
        mean = sum(times) / len(times)
        return sample.n_triangles / mean, (nvox or 0) / mean, {
            "kind": kind, "cores": 1, "host_cores": os.cpu_count(), "mean_s": mean, "best_s": min(times), "runs": len(times),
            "n_voxels": nvox, "n_nodes": nnodes, "n_triangles": sample.n_triangles, "whole": sample is mesh, "whole_counts": whole_counts,
            "sample": "%s (%d triangles, -s %d), %d timed run(s) of the %s after %d warm-up run(s), wall clock incl. its file IO, page cache warm"
                      % (note, sample.n_triangles, gridsize, len(times),
                         "unmodified reference CLI (oracle/_ref/svo_builder_binary)" if kind == "reference" else "C restatement (oracle/liboracle.so)", warmup)}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def reference_arm(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from ooc_svo_builder_b200 import estimate_partitions
    mesh, grid, wl = workload(args.gpus)
    tps, vps, info = run_reference_cpu(mesh, grid, max(1, args.steps), max(args.warmup, 3))
    # `config` describes the WHOLE workload; its counts come from the (untimed) probe run of the whole workload
    line = {
        "impl": "reference", "metric": METRIC, "value": tps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": info["mean_s"] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "voxels_per_s": vps,
        "config": config_dict(wl, grid, mesh.n_triangles, info["whole_counts"][0], info["whole_counts"][1],
                              estimate_partitions(grid, 2048), args.gpus),
        "cpu_baseline": {"value": tps, "unit": UNIT, "cores": 1, "kind": info["kind"], "sample": info["sample"], "host_cores": info["host_cores"]},
        "e2e": {"value": tps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------
# roofline bookkeeping
# ----------------------------------------------------------------------------
def load_profile():
    try:
        with open(PROFILE_JSON) as f:
            return json.load(f)
    except Exception:
        return {}


def leaf_emitter_bytes(n_brick_records, n_bricks):
    """Algorithmic bytes of k_emit_leaf: 24 B per record of the brick subtrees (leaves + depth D-1 nodes) and per brick
    record (the parent's children blocks, written by the same kernel), 24 B read per brick (key, word, file base)."""
    return 24 * (n_brick_records + n_bricks) + 24 * n_bricks


def frac_entry(nbytes: float, ms: float, peak: float) -> dict:
    ach = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
    return {"ms": ms, "algorithmic_bytes": int(nbytes), "achieved": ach, "unit": "GB/s", "frac": ach / peak}


# ----------------------------------------------------------------------------
# our arm, one GPU
# ----------------------------------------------------------------------------
def ours_single(args, peak, peak_src, sm_max_mhz):
    import torch
    from ooc_svo_builder_b200 import SvoBuilder, PinnedBuffer, meshgen

    local = int(os.environ.get("LOCAL_RANK", "0"))
    mesh, grid, wl = workload(1)
    T = mesh.n_triangles
    sb = SvoBuilder(local)
    stream = torch.cuda.Stream()
    sb.set_stream(stream.cuda_stream)
    prm = sb.make_params(mesh.length, grid, False)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2
    warm = max(args.warmup, 3)

    with torch.cuda.stream(stream):
        d_tris = torch.from_numpy(mesh.tris).cuda()
        torch.cuda.synchronize()
        sb.set_triangles(d_tris)

        def step():
            sb.partition(prm, want_counts=False)
            sb.voxelize()
            return sb.build()

        for _ in range(warm):
            flush.zero_()
            nv, nn, nd = step()
        torch.cuda.synchronize()
        sampler = ClockSampler(local)
        sampler.start()
        time.sleep(0.3)
        t_wall0 = time.time()
        evs, per_stage = [], []
        for _ in range(args.steps):
            flush.zero_()                                  # L2 flush between timed iterations (outside the event pair)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            nv, nn, nd = step()
            e1.record(stream)
            evs.append((e0, e1))
            per_stage.append(sb.stats())
        torch.cuda.synchronize()
        t_wall1 = time.time()
        step_ms = [a.elapsed_time(b) for a, b in evs]
        clocks = sampler.stop(t_wall0, t_wall1)
    ms_per_step = sum(step_ms) / len(step_ms)
    value = T / (ms_per_step * 1e-3)

    def avg(k):
        return sum(s[k] for s in per_stage) / len(per_stage)

    st = per_stage[-1]
    launches = st["kernel_launches"]
    # ---- roofline: every kernel against ITS OWN algorithmic bytes and ITS OWN time (CUDA events inside the library,
    # on its launch stream), the octree-build stage against the stage's bytes and the stage's time ----
    prof = load_profile()
    n_bricks, n_t1, n_brec = st["n_bricks"], st["n_tiles1"], st["n_brick_records"]
    vox_bytes = T * 36 + 8 * n_bricks + 8 * n_t1          # records read + one 64-bit word per touched brick / level-1 tile
    leaf_bytes = leaf_emitter_bytes(n_brec, n_bricks)
    kernels = {
        "k_vox_warp": frac_entry(vox_bytes, avg("ms_vox_small"), peak),
        "k_emit_leaf": frac_entry(leaf_bytes, avg("ms_emit_leaf"), peak),
    }
    dom = max(kernels, key=lambda k: kernels[k]["ms"])
    issue_peak = 148 * 4 * sm_max_mhz * 1e6 / 1e9          # G warp-instructions/s: 148 SMs x 4 schedulers x clock
    issue = None
    if "k_vox_warp" in prof and "warp_instructions" in prof["k_vox_warp"]:
        wi = prof["k_vox_warp"]["warp_instructions"]
        ach = wi / (kernels["k_vox_warp"]["ms"] * 1e-3) / 1e9
        issue = {"bound": "issue", "kernel": "k_vox_warp", "warp_instructions_per_launch": wi, "warp_instructions_per_triangle": wi / T,
                 "achieved": ach, "peak": issue_peak, "unit": "G warp-instructions/s", "frac": ach / issue_peak,
                 "source": "instruction count from the committed ncu capture of the same step (%s), time measured in this run" % os.path.relpath(PROFILE_JSON, ROOT)}
    traffic = None
    if dom in prof:
        traffic = prof[dom].get("dram_bytes_read", 0) + prof[dom].get("dram_bytes_write", 0)
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved"], "peak": peak, "unit": "GB/s", "frac": kernels[dom]["frac"],
        "traffic": traffic, "traffic_source": "committed ncu --set full capture of the same step (%s); not measured in this run" % os.path.relpath(PROFILE_JSON, ROOT),
        "peak_source": peak_src, "algorithmic_bytes": kernels[dom]["algorithmic_bytes"], "kernel_ms": kernels[dom]["ms"],
        "note": "each kernel: its own algorithmic bytes / its own CUDA-event time. k_vox_warp: T*36 B records + 8 B per touched brick and level-1 "
                "tile; it is instruction-issue bound (see `issue`), not HBM bound. k_emit_leaf: 24 B per record it writes (leaves, depth D-1 nodes and the "
                "bricks' own records) + 24 B read per brick (key, word, file base). octree_build_stage: SURVEY.md 8d bytes (8*N + 24*N_nodes) / the whole build stage (ms_build).",
        "issue": issue, "kernels": kernels,
        "octree_build_stage": frac_entry(8 * nv + 24 * nn, avg("ms_build"), peak),
    }

    # ---- e2e: one C-ABI call (svo_run) per step with HOST (pinned) buffers: H2D of the step's triangles and D2H of
    # the step's node + data files inside the timed region. Two contexts on two host threads keep two steps in
    # flight, so the upload of step i+1 overlaps compute + download of step i (PCIe is full duplex); the
    # single-step latency (one context, nothing overlapped) is reported next to the pipelined throughput.
    n_pipe = 2
    ctxs = [sb, SvoBuilder(local)]
    bufs = []
    for _ in range(n_pipe):
        h_tris = PinnedBuffer(mesh.tris.nbytes)
        h_tris.array[:] = mesh.tris.view(np.uint8).reshape(-1)
        bufs.append((h_tris, PinnedBuffer(nn * 24), PinnedBuffer(nd * 32)))
    sb.set_stream(None)                                       # each context on its own stream
    views = [b[0].array.view(np.float32).reshape(T, 9) for b in bufs]
    lat = []
    for i in range(warm + args.steps):                        # latency: one step at a time
        sb.synchronize()
        t = time.perf_counter()
        sb.run_host(prm, views[0], bufs[0][1].array, bufs[0][2].array)
        if i >= warm:
            lat.append(time.perf_counter() - t)
    lat_s = sum(lat) / len(lat)
    for k in range(n_pipe):                                   # warm both contexts
        for _ in range(3):
            ctxs[k].run_host(prm, views[k], bufs[k][1].array, bufs[k][2].array)
    per_thread = (args.steps + n_pipe - 1) // n_pipe

    def worker(k):
        for _ in range(per_thread):
            ctxs[k].run_host(prm, views[k], bufs[k][1].array, bufs[k][2].array)

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(n_pipe)]
    t = time.perf_counter()
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    pipe_s = (time.perf_counter() - t) / (per_thread * n_pipe)
    assert bufs[1][1].array.tobytes() == bufs[0][1].array.tobytes()      # both contexts produced the same file image
    ctxs[1].close()
    e2e = {"value": T / pipe_s, "unit": UNIT, "h2d_bytes_per_step": int(mesh.tris.nbytes), "d2h_bytes_per_step": int(nn * 24 + nd * 32),
           "ms_per_step": pipe_s * 1e3, "latency_ms": lat_s * 1e3, "latency_value": T / lat_s, "in_flight": n_pipe,
           "api": "svo_run (C ABI): pinned host triangles in, node + data file images out, wall clock; %d contexts / host threads keep %d "
                  "steps in flight (upload of one overlaps compute + download of the other); latency_* = one step alone" % (n_pipe, n_pipe)}

    # ---- CPU baseline: the reference itself on this box's host cores (bounded: one timed run, ~5 s) ----
    cpu_tps, cpu_vps, info = run_reference_cpu(mesh, grid, 1, 1)
    cpu_baseline = {"value": cpu_tps, "unit": UNIT, "cores": 1, "kind": info["kind"], "sample": info["sample"],
                    "host_cores": info["host_cores"], "voxels_per_s": cpu_vps, "seconds": info["mean_s"]}

    # ---- e2e at the PROCESS boundary (the reference's only real interface): our CLI against the reference CLI on the
    # same files (in /dev/shm), wall clock of the whole process incl. CUDA start-up and file IO ----
    e2e_cli = None
    try:
        d = tempfile.mkdtemp(prefix="svo_bench_cli_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        hdr = meshgen.write_tri(os.path.join(d, "m"), mesh)
        exe = os.path.join(ROOT, "ooc_svo_builder_b200", "bin", "svo_builder_binary")
        ts = []
        for _ in range(3):
            t = time.perf_counter()
            subprocess.run([exe, "-f", hdr, "-s", str(grid)], capture_output=True, text=True, check=True)
            ts.append(time.perf_counter() - t)
        ours_s = min(ts)
        e2e_cli = {"ours_s": ours_s, "reference_s": info["mean_s"], "ratio": info["mean_s"] / ours_s, "value": T / ours_s, "unit": UNIT,
                   "what": "wall clock of the whole process on the same .tri/.tridata files in /dev/shm: ooc_svo_builder_b200/bin/svo_builder_binary "
                           "(best of 3, CUDA context creation + file IO included) vs oracle/_ref/svo_builder_binary"}
        shutil.rmtree(d, ignore_errors=True)
    except Exception as e:      # noqa: BLE001
        e2e_cli = {"error": str(e)[:200]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "voxels_per_s": nv / (ms_per_step * 1e-3),
        "config": config_dict(wl, grid, T, nv, nn, st["n_partitions"], 1),
        "timing": {"l2": "flushed between timed iterations (256 MB write)", "clock": "CUDA events on the launch stream, per step",
                   "speculative_builds": int(sum(s["speculative"] for s in per_stage))},
        "stage_ms": {k: avg(k) for k in ("ms_partition", "ms_voxelize", "ms_vox_small", "ms_compact", "ms_build", "ms_emit", "ms_emit_leaf", "ms_clear")},
        "pairs": {k: st[k] for k in ("n_pairs", "n_small", "n_medium", "n_large")},
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "e2e_cli": e2e_cli,
        "gpu_launches": int(launches) * args.steps, "gpu_launches_per_step": int(launches),
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    sb.close()


# ----------------------------------------------------------------------------
# our arm, N GPUs (torchrun): weak scaling of the sharded path + byte parity + the strong-scaling record
# ----------------------------------------------------------------------------
def device_range_equal(torch, sb_sharded, sb_single, nlo, nhi):
    """Compares records [nlo, nhi) of the sharded context's node buffer with the same records of a single-GPU build,
    on the device. Returns (equal, bytes compared)."""
    n = nhi - nlo
    if n == 0:
        return True, 0
    ok = True
    chunk = 1 << 26                                     # records per comparison (1.6 GB)
    for lo in range(nlo, nhi, chunk):
        k = min(chunk, nhi - lo)
        a = torch.empty(k * 24, dtype=torch.uint8, device="cuda")
        b = torch.empty(k * 24, dtype=torch.uint8, device="cuda")
        sb_sharded.fetch_nodes(lo, k, a)
        sb_single.fetch_nodes(lo, k, b)
        ok = ok and bool(torch.equal(a, b))
        del a, b
    return ok, n * 24


def ours_sharded(args, rank, world, local, dist, peak, peak_src):
    import hashlib
    import torch
    from ooc_svo_builder_b200 import SvoBuilder, PinnedBuffer
    from ooc_svo_builder_b200.sharded import DistributedBuilder, slice_bounds

    mesh, G, wl = workload(world)
    tris, length = mesh.tris, mesh.length
    T = tris.shape[0]
    lo_t, hi_t = slice_bounds(T, world, rank)
    db = DistributedBuilder(dist, local)
    stream = torch.cuda.Stream()
    db.set_stream(stream)
    prm = SvoBuilder.make_params(length, G, False)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    warm = max(args.warmup, 3)
    # Input path (SVO_BENCH_INPUT): "remote" (default) = remote staging of triangle slices over NVLink peer memory;
    # "dispatch" = copying all-to-all of triangle records into peer inboxes; "replicated" = every rank holds the whole
    # mesh. Both peer-memory modes need CUDA IPC; if the box cannot map peer memory, fall back to "replicated"
    # (decided collectively, reported in `input`).
    mode = os.environ.get("SVO_BENCH_INPUT", "remote")
    per = (T + world - 1) // world
    ok = torch.ones(1, dtype=torch.int32, device="cuda")
    if mode != "replicated":
        try:
            if mode == "remote":
                db.enable_slices(per, 9, T)
            else:
                db.enable_dispatch(T, 9)
        except Exception as e:      # noqa: BLE001
            print("rank %d: peer memory unavailable (%s)" % (rank, e), flush=True)
            ok.zero_()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if not bool(int(ok)):
        mode = "replicated"
        db.sliced = False
    with torch.cuda.stream(stream):
        if mode == "remote":
            db.upload_slice(torch.from_numpy(tris[lo_t:hi_t]).cuda())
            torch.cuda.synchronize()
        elif mode == "dispatch":
            d_local = torch.from_numpy(tris[lo_t:hi_t]).cuda()
            torch.cuda.synchronize()
            db.set_local_triangles(d_local)
        else:
            d_tris = torch.from_numpy(tris).cuda()
            torch.cuda.synchronize()
            db.set_triangles(d_tris)
        for _ in range(warm):
            flush.zero_()
            nv, nn, nd = db.step(prm)
        torch.cuda.synchronize()
        dist.barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
            time.sleep(0.3)
        t0 = time.time()
        evs, spec_steps, retries = [], 0, 0
        for _ in range(args.steps):
            flush.zero_()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            nv, nn, nd = db.step(prm)
            e1.record(stream)
            evs.append((e0, e1))
            spec_steps += db.sb.stats()["speculative"]
            retries += db.retries
        torch.cuda.synchronize()
        dist.barrier()
        t1 = time.time()
        ms = torch.tensor([a.elapsed_time(b) for a, b in evs], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)                     # per step: the slowest rank
    ms_per_step = float(ms.mean())
    st = db.sb.stats()
    launches = torch.tensor([st["kernel_launches"]], dtype=torch.int64, device="cuda")
    dist.all_reduce(launches)
    stage_keys = ("ms_dispatch", "ms_peer_wait", "ms_partition", "ms_voxelize", "ms_vox_small", "ms_compact", "ms_build", "ms_emit", "ms_emit_leaf", "ms_clear")
    stage = torch.tensor([st[k] for k in stage_keys], dtype=torch.float64, device="cuda")
    dist.all_reduce(stage, op=dist.ReduceOp.MAX)
    stage_max = {k: float(v) for k, v in zip(stage_keys, stage)}
    nlo, nhi, dlo, dhi = db.sb.shard_ranges()

    # ---- byte parity of the path just measured: this rank's [node_lo, node_hi) against a plain single-GPU build of the
    # whole mesh in a second context on the same device, compared on the device ----
    single = SvoBuilder(local)
    d_all = torch.from_numpy(tris).cuda()
    single.set_triangles(d_all)
    single.partition(prm, want_counts=False)
    single.voxelize()
    s_nv, s_nn, s_nd = single.build()
    eq, nbytes = device_range_equal(torch, db.sb, single, nlo, nhi)
    eq = eq and (s_nv, s_nn, s_nd) == (nv, nn, nd)
    flag = torch.tensor([1 if eq else 0], dtype=torch.int64, device="cuda")
    dist.all_reduce(flag)
    ranges = [None] * world
    dist.all_gather_object(ranges, (nlo, nhi))
    nodes_sha = None
    if rank == 0:
        h = hashlib.sha256()
        for lo in range(0, s_nn, 1 << 24):
            h.update(single.fetch_nodes(lo, min(1 << 24, s_nn - lo)).tobytes())
        nodes_sha = h.hexdigest()
    tiled = sorted(r for r in ranges if r[1] > r[0])
    tiles_ok = bool(tiled) and tiled[0][0] == 0 and tiled[-1][1] == nn and all(a[1] == b[0] for a, b in zip(tiled, tiled[1:]))
    parity = {"checked": True, "ranks_equal": int(flag), "ranges_tile_the_file": tiles_ok, "nodes_sha256": nodes_sha,
              "what": "every rank: records [node_lo, node_hi) of the sharded build (the multi-process NVLink / IPC path timed above) == the same records of "
                      "a single-GPU svo_build of the whole mesh, byte compare on the device; nodes_sha256 = the single-GPU node file (rank 0)"}
    single.close()
    del d_all
    torch.cuda.empty_cache()

    # ---- e2e: HOST triangles in, this rank's node / data range out to pinned host memory. Every rank uploads only its
    # 1/N slice of the triangle file over PCIe; the records reach the ranks that voxelize them over NVLink, the sharded
    # step runs and each rank fetches its own range of the output files. Same method as N = 1: TWO steps in flight -- two
    # independent sets of contexts / peer windows / streams per rank, each driven by its own host thread -- and no
    # per-step barrier: the ranks are coupled by the exchange inside a step only. (Input modes other than "remote" use
    # NCCL inside the step and keep one step in flight.)
    h_slice = torch.empty((per, 9), dtype=torch.float32).pin_memory()
    h_slice[: hi_t - lo_t].copy_(torch.from_numpy(tris[lo_t:hi_t]))
    use_dispatch = mode != "replicated"
    n_lanes = 2 if mode == "remote" else 1
    lanes = [(db, stream)]
    if n_lanes == 2:
        db2 = DistributedBuilder(dist, local)
        stream2 = torch.cuda.Stream()
        db2.set_stream(stream2)
        db2.enable_slices(per, 9, T)
        lanes.append((db2, stream2))
    d_in = torch.empty((per if use_dispatch else world * per, 9), dtype=torch.float32, device="cuda")
    outs = [(PinnedBuffer(max(nhi - nlo, 1) * 24 + 24 * 4096), PinnedBuffer(64)) for _ in lanes]

    def e2e_step(k):
        dbk, _ = lanes[k]
        h_nodes, h_data = outs[k]
        if mode == "remote":
            dbk.upload_slice(h_slice.numpy()[: hi_t - lo_t])                                       # PCIe: 1/N of the mesh, NVLink: inside step()
        elif mode == "dispatch":
            d_in.copy_(h_slice, non_blocking=True)
            dbk.set_local_triangles(d_in[: hi_t - lo_t])
        else:
            d_in[rank * per:(rank + 1) * per].copy_(h_slice, non_blocking=True)
            dist.all_gather_into_tensor(d_in, d_in[rank * per:(rank + 1) * per])
            dbk.set_triangles(d_in[:T])
        dbk.step(prm)
        a, b, c_, d = dbk.sb.shard_ranges()
        dbk.sb.fetch_nodes(a, b - a, h_nodes.array[: (b - a) * 24])
        if d > c_:
            dbk.sb.fetch_data(c_, d - c_, h_data.array[: (d - c_) * 32])

    errors = []

    def lane_worker(k, n):
        try:
            torch.cuda.set_device(local)
            with torch.cuda.stream(lanes[k][1]):
                for _ in range(n):
                    e2e_step(k)
        except Exception as e:      # noqa: BLE001
            errors.append("%s: %s" % (type(e).__name__, e))

    for k in range(n_lanes):                                  # warm every lane (first step of a context is a sized build)
        lane_worker(k, 3)
    torch.cuda.synchronize()
    dist.barrier()
    per_lane = (args.steps + n_lanes - 1) // n_lanes
    ths = [threading.Thread(target=lane_worker, args=(k, per_lane)) for k in range(n_lanes)]
    t = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device="cuda")
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_s = float(dt) / (per_lane * n_lanes)
    if errors:
        raise SystemExit("e2e lane failed on rank %d: %s" % (rank, errors[0]))
    assert outs[0][0].array[: (nhi - nlo) * 24].tobytes() == outs[-1][0].array[: (nhi - nlo) * 24].tobytes()      # both lanes: the same file range
    if n_lanes == 2:
        lanes[1][0].close()

    strong = None
    if os.environ.get("SVO_BENCH_STRONG", "1") != "0":
        try:
            strong = strong_scaling_record(torch, dist, db, rank, world, local, stream, peak)
        except Exception as e:      # noqa: BLE001
            strong = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
        gathered = [None] * world
        dist.all_gather_object(gathered, strong if (strong and "error" in strong) else None)
        errs = [g for g in gathered if g]
        if errs and rank == 0:
            strong = errs[0]

    if rank == 0:
        clocks = sampler.stop(t0, t1)
        value = T / (ms_per_step * 1e-3)
        lm = stage_max["ms_emit_leaf"]
        leaf_bytes = leaf_emitter_bytes(st["n_brick_records"], st["n_bricks"])         # rank 0's own bricks
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "voxels_per_s": nv / (ms_per_step * 1e-3),
            "config": config_dict(wl, G, T, nv, nn, 8, world),
            "input": {"triangle_input": mode,
                      "how": {"remote": "each rank starts with 1/N of the triangle file in HBM; the voxelizer kernel stages the triangle blocks it needs straight from "
                                        "the owner's HBM with cp.async.bulk over NVLink (no copy); file ordered so that every record crosses NVLink",
                              "dispatch": "our own all-to-all kernel stores records into peer HBM over NVLink; file ordered so that every record crosses NVLink",
                              "replicated": "every rank holds the whole mesh"}[mode],
                      "table_exchange": "peer-memory stores + epoch flags (our own kernels)" if (mode == "remote" and os.environ.get("SVO_TABLE_EXCHANGE", "peer") == "peer") else "NCCL all-reduce"},
            "timing": {"l2": "flushed between timed iterations (256 MB write)", "clock": "CUDA events on the launch stream per step, max over ranks, barrier before every step",
                       "speculative_builds_rank0": int(spec_steps), "retries_rank0": int(retries)},
            "parity": parity,
            "roofline": {"bound": "hbm", "kernel": "k_emit_leaf (rank 0's bytes / slowest rank's time)", "achieved": leaf_bytes / max(lm, 1e-9) / 1e6, "peak": peak, "unit": "GB/s",
                         "frac": leaf_bytes / max(lm, 1e-9) / 1e6 / peak, "traffic": None, "peak_source": peak_src, "algorithmic_bytes": int(leaf_bytes), "kernel_ms": lm,
                         "note": "k_emit_leaf: 24 B per record it writes (brick subtrees + the bricks' own records) + 24 B read per brick, its own time. octree_build_stage: rank 0's share of 8*N + 24*N_nodes "
                                 "over the slowest rank's ms_build (includes the wait for the table exchange).",
                         "octree_build_stage": frac_entry((8 * nv + 24 * nn) / world, stage_max["ms_build"], peak)},
            "e2e": {"value": T / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(tris.nbytes),
                    "d2h_bytes_per_step": int(nn * 24 + nd * 32), "ms_per_step": e2e_s * 1e3,
                    "in_flight": n_lanes,
                    "api": "per rank: pinned H2D of 1/N of the .tridata + %s + sharded step + svo_fetch_* of its file range to pinned host memory; wall clock "
                           "over %d steps, %d in flight (independent context sets on their own host threads), max over ranks, no per-step barrier" % (
                               {"remote": "remote staging over NVLink inside the voxelizer", "dispatch": "triangle dispatch over NVLink",
                                "replicated": "NCCL all-gather over NVLink"}[mode], per_lane * n_lanes, n_lanes)},
            "gpu_launches": int(launches) * args.steps, "clocks": clocks,
            "stage_ms_max_over_ranks": stage_max,
            "pairs_rank0": {k: st[k] for k in ("n_pairs", "n_small", "n_medium", "n_large")},
            "strong": strong,
        }
        print(json.dumps(line), flush=True)
    db.close()
    dist.barrier()
    dist.destroy_process_group()
    if int(flag) != world or not tiles_ok:
        raise SystemExit("PARITY FAILURE: the sharded build differs from the single-GPU build on %d of %d ranks" % (world - int(flag), world))


def strong_scaling_record(torch, dist, db_unused, rank, world, local, stream, peak):
    """North-star record: BASELINE.json configs[4] (8192^3, 100 M-triangle thin shell) built by all N GPUs (every rank
    holds 1/N of the file) against the same build on ONE GPU (rank 0), both checked against the reference's golden file
    checksum (tests/golden/golden_scale.json, position-aware additive checksum: tests/filesum.py)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from filesum import filesum_torch, add as fs_add
    from ooc_svo_builder_b200 import SvoBuilder, meshgen
    from ooc_svo_builder_b200.sharded import DistributedBuilder, slice_bounds
    name = os.environ.get("SVO_BENCH_STRONG_CONFIG", "c5")
    cfg, g = {"c4": ("c4_sphere_4096", 4096), "c5": ("c5_shell_8192", 8192)}[name]
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_scale.json"))).get(name)
    mesh = meshgen.make(cfg)
    T = mesh.n_triangles
    prm = SvoBuilder.make_params(mesh.length, g, False)
    db = DistributedBuilder(dist, local)
    db.set_stream(stream)
    lo, hi = slice_bounds(T, world, rank)
    out = {"config": name, "gridsize": g, "n_triangles": T, "world": world}
    with torch.cuda.stream(stream):
        db.enable_slices((T + world - 1) // world, 9, T)
        db.upload_slice(mesh.tris[lo:hi])
        torch.cuda.synchronize()
        ms = []
        for i in range(6):
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            nv, nn, nd = db.step(prm)
            e1.record(stream)
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        t = torch.tensor(ms[2:], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_n = float(t.mean())
    st = db.sb.stats()
    nlo, nhi, _, _ = db.sb.shard_ranges()
    # golden check of the N-GPU result: every rank sums its own range on the device
    mine = (0, 0, 0)
    for a in range(nlo, nhi, 1 << 26):
        k = min(1 << 26, nhi - a)
        buf = torch.empty(k * 3, dtype=torch.int64, device="cuda")
        db.sb.fetch_nodes(a, k, buf)
        mine = fs_add(mine, filesum_torch(buf, a * 3))
        del buf
    sums = [None] * world
    dist.all_gather_object(sums, (mine, (nlo, nhi), st["ms_voxelize"], st["ms_build"], st["ms_emit_leaf"], st["ms_dispatch"], st["n_brick_records"], st["n_bricks"]))
    total = (0, 0, 0)
    for s in sums:
        total = fs_add(total, s[0])
    out.update({"n_voxels": nv, "n_nodes": nn, "ms_n_gpus": ms_n,
                "per_rank": [{"node_range": list(s[1]), "ms_voxelize": s[2], "ms_build": s[3], "ms_emit_leaf": s[4], "ms_filter": s[5],
                              "build_stage_hbm_frac": (8 * nv + 24 * nn) / world / max(s[3], 1e-9) / 1e6 / peak,
                              "emit_leaf_hbm_frac": (24 * s[6] + 16 * s[7]) / max(s[4], 1e-9) / 1e6 / peak} for s in sums]})
    if gold:
        out["golden"] = {"n_voxels_ok": nv == gold["n_voxels"], "n_nodes_ok": nn == gold["n_nodes"],
                         "nodes_filesum_ok": [int(x) for x in total] == gold["nodes_filesum"], "mesh": "regenerated (deterministic generator, tools/meshcheck.py)"}
    db.close()
    del db
    torch.cuda.empty_cache()
    # the same build on ONE GPU (rank 0); the others wait
    if rank == 0:
        sb = SvoBuilder(local)
        sb.set_stream(stream.cuda_stream)
        with torch.cuda.stream(stream):
            sb.set_triangles(mesh.tris)
            ms1 = []
            for i in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                sb.partition(prm, want_counts=False)
                sb.voxelize()
                nv1, nn1, nd1 = sb.build()
                e1.record(stream)
                torch.cuda.synchronize()
                ms1.append(e0.elapsed_time(e1))
        st1 = sb.stats()
        one = (0, 0, 0)
        for a in range(0, nn1, 1 << 26):
            k = min(1 << 26, nn1 - a)
            buf = torch.empty(k * 3, dtype=torch.int64, device="cuda")
            sb.fetch_nodes(a, k, buf)
            one = fs_add(one, filesum_torch(buf, a * 3))
            del buf
        ms_1 = sum(ms1[1:]) / len(ms1[1:])
        out.update({"ms_1_gpu": ms_1, "speedup": ms_1 / ms_n,
                    "one_gpu": {"ms_voxelize": st1["ms_voxelize"], "ms_build": st1["ms_build"], "ms_emit_leaf": st1["ms_emit_leaf"],
                                "build_stage_hbm_frac": (8 * nv1 + 24 * nn1) / max(st1["ms_build"], 1e-9) / 1e6 / peak,
                                "emit_leaf_hbm_frac": leaf_emitter_bytes(st1["n_brick_records"], st1["n_bricks"]) / max(st1["ms_emit_leaf"], 1e-9) / 1e6 / peak,
                                "nodes_filesum_ok": (None if not gold else [int(x) for x in one] == gold["nodes_filesum"])}})
        sb.close()
    dist.barrier()
    return out


def ours(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    peak, peak_src, sm_max = load_peaks()
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        return ours_sharded(args, rank, world, local, dist, peak, peak_src)
    return ours_single(args, peak, peak_src, sm_max)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
