#!/usr/bin/env python
"""Strong scaling of a BASELINE.json config over the GPUs of one box (torchrun, one rank per GPU, NCCL):
every rank holds 1/N of the triangle file (the others arrive by the NVLink triangle dispatch; SVO_BENCH_INPUT=remote|dispatch|replicated), voxelizes + builds its partitions, the subtree table is all-reduced, every rank emits
its range of the node file. Prints the max-over-ranks step time and checks the global counts and that the
per-rank ranges tile the file.

    python -m torch.distributed.run --nproc-per-node 8 tools/scale_sharded.py c5
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from ooc_svo_builder_b200 import SvoBuilder, meshgen  # noqa: E402
from ooc_svo_builder_b200.sharded import DistributedBuilder, slice_bounds  # noqa: E402

CFG = {"c2": ("c2_displaced_sphere_1024", 1024), "c3": ("c3_terrain_2048_payload", 2048), "c4": ("c4_sphere_4096", 4096), "c5": ("c5_shell_8192", 8192)}


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    out = {}
    for n in sys.argv[1:]:
        cfg, g = CFG[n]
        mesh = meshgen.make(cfg)
        db = DistributedBuilder(dist, local)
        stream = torch.cuda.Stream()
        db.set_stream(stream)
        prm = SvoBuilder.make_params(mesh.length, g, mesh.payload)
        mode = os.environ.get("SVO_BENCH_INPUT", "remote")
        use_dispatch = mode
        with torch.cuda.stream(stream):
            lo, hi = slice_bounds(mesh.n_triangles, world, rank)
            if mode == "remote":
                db.enable_slices((mesh.n_triangles + world - 1) // world, mesh.tris.shape[1], mesh.n_triangles)
                db.upload_slice(mesh.tris[lo:hi])
                torch.cuda.synchronize()
                d = None
            elif mode == "dispatch":
                d = torch.from_numpy(mesh.tris[lo:hi]).cuda()
                torch.cuda.synchronize()
                db.enable_dispatch(mesh.n_triangles, mesh.tris.shape[1])
                db.set_local_triangles(d)
            else:
                d = torch.from_numpy(mesh.tris).cuda()
                torch.cuda.synchronize()
                db.set_triangles(d)
            ms = []
            for i in range(5):
                dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                nv, nn, nd = db.step(prm)
                e1.record(stream)
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            t = torch.tensor(ms[2:], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        st = db.sb.stats()
        nlo, nhi, dlo, dhi = db.sb.shard_ranges()
        rng = [None] * world
        dist.all_gather_object(rng, (nlo, nhi, st["ms_voxelize"], st["ms_build"], st["ms_emit_leaf"], st["ms_dispatch"]))
        if rank == 0:
            pos = 0
            for lo, hi, *_ in rng:
                assert lo == pos or lo == hi, (rng, pos)
                pos = max(pos, hi)
            assert pos == nn
            step = float(t.mean())
            out[n] = {"world": world, "n_triangles": mesh.n_triangles, "gridsize": g, "n_voxels": nv, "n_nodes": nn, "ms_per_step_max_over_ranks": step,
                      "triangles_per_s": mesh.n_triangles / (step * 1e-3), "voxels_per_s": nv / (step * 1e-3),
                      "per_rank": [{"node_range": [r[0], r[1]], "ms_voxelize": r[2], "ms_build": r[3], "ms_emit_leaf": r[4], "ms_dispatch": r[5]} for r in rng],
                      "triangle_dispatch": use_dispatch}
            print(n, json.dumps(out[n]), flush=True)
        db.close()
        del mesh, d
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "scale_sharded_%d.json" % world), "w") as f:
            json.dump(out, f, indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
