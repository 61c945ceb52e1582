"""Worker of tests/test_gpu_multi.py: one rank of a torch.distributed job (one process per GPU, NCCL for the plumbing,
CUDA-IPC peer windows for the data path). Builds the cases with DistributedBuilder -- remote staging of triangle slices
over NVLink, peer-memory table exchange -- and compares the assembled per-rank file ranges with the CPU oracle on
rank 0. Writes a JSON verdict; exits non-zero on any mismatch."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from ooc_svo_builder_b200 import SvoBuilder, meshgen as mg  # noqa: E402
from ooc_svo_builder_b200.sharded import DistributedBuilder, slice_bounds  # noqa: E402


def cases():
    ico = mg.icosphere(5)
    return [
        # name, mesh, gridsize, memory_limit_mb, color
        ("binary_sphere_p8", mg.displaced_sphere(300, 300, seed=5), 512, 100, "model"),
        ("binary_soup_p8", mg.random_soup(4000, seed=13, large_frac=0.02), 256, 2, "model"),
        ("binary_sphere_p8_again", mg.displaced_sphere(300, 300, seed=5), 512, 100, "model"),      # speculative local builds
        ("binary_single_partition", ico, 128, 2048, "model"),
        ("payload_terrain_p8", mg.terrain(150, seed=2), 256, 3, "linear"),
        ("payload_ico_p1", mg.Mesh(mg.with_payload(ico.tris), ico.length), 256, 2048, "model"),
        ("binary_sphere_big_then_retry", mg.displaced_sphere(500, 500, seed=6), 1024, 2048, "model"),  # outgrows the lists: SVO_E_RETRY on all ranks
        ("empty", mg.empty_mesh(), 256, 3, "model"),
    ]


def main():
    out_path = sys.argv[1]
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    results = {}
    ok_all = True
    builders = {}
    all_cases = cases()
    cap = {}
    for _, mesh, _, _, _ in all_cases:      # one window per record size, large enough for every case's slices
        fpt = mesh.tris.shape[1]
        cap[fpt] = max(cap.get(fpt, 1), (mesh.n_triangles + world - 1) // world)
    for name, mesh, g, lim, color in all_cases:
        fpt = mesh.tris.shape[1]
        T = mesh.n_triangles
        if fpt not in builders:
            db = DistributedBuilder(dist, local)
            stream = torch.cuda.Stream()
            db.set_stream(stream)
            db.enable_slices(cap[fpt], fpt, T)
            builders[fpt] = (db, stream)
        db, stream = builders[fpt]
        db.n_total = T
        prm = SvoBuilder.make_params(mesh.length, g, fpt == 21, lim, False, color)
        lo, hi = slice_bounds(T, world, rank)
        with torch.cuda.stream(stream):
            db.upload_slice(np.ascontiguousarray(mesh.tris[lo:hi]))
            nv, nn, nd = db.step(prm)
            nlo, nhi, dlo, dhi = db.sb.shard_ranges()
            nodes = db.sb.fetch_nodes(nlo, nhi - nlo)
            data = db.sb.fetch_data(dlo, dhi - dlo)
        st = db.sb.stats()
        parts = [None] * world
        dist.all_gather_object(parts, (rank, nv, nn, nd, (nlo, nhi), (dlo, dhi), nodes.tobytes(), data.tobytes(), st["speculative"], db.retries))
        if rank == 0:
            from oracle import oracle as O
            want = O.build(mesh.tris, mesh.length, g, memory_limit_mb=lim, color=color)
            node_img = bytearray(nn * 24)
            data_img = bytearray(nd * 32)
            pos = 0
            tiles = True
            for r, _nv, _nn, _nd, nr, dr, nb, dbb, _s, _rt in sorted(parts):
                tiles = tiles and (nr[0] == pos or nr[0] == nr[1])
                pos = max(pos, nr[1])
                node_img[nr[0] * 24: nr[1] * 24] = nb
                data_img[dr[0] * 32: dr[1] * 32] = dbb
            same = (bytes(node_img) == want.nodes and bytes(data_img) == want.data and nv == want.n_voxels and
                    nn == want.n_nodes and nd == want.n_data and tiles and pos == nn and
                    all(p[1:4] == (nv, nn, nd) for p in parts))
            results[name] = {"ok": bool(same), "n_voxels": nv, "n_nodes": nn, "n_data": nd, "speculative": [p[8] for p in sorted(parts)],
                             "retries": [p[9] for p in sorted(parts)]}
            ok_all = ok_all and same
    for db, _ in builders.values():
        db.close()
    flag = torch.tensor([1 if ok_all else 0], device="cuda")
    dist.broadcast(flag, 0)
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump({"world": world, "ok": bool(ok_all), "cases": results}, f, indent=1)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
