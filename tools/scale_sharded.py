#!/usr/bin/env python
"""Strong scaling of a BASELINE.json config over the GPUs of one box (torchrun, one rank per GPU): every rank holds 1/N
of the triangle file, voxelizes + builds its partitions, the subtree table is exchanged over peer memory, every rank
emits its range of the node file. The N-GPU result and the 1-GPU result are both checked against the reference's golden
file checksum (tests/golden/golden_scale.json). The same record bench.py --gpus N attaches as `strong`.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/scale_sharded.py c4 c5
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import bench  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    peak, _, _ = bench.load_peaks()
    out = {}
    stream = torch.cuda.Stream()
    for n in sys.argv[1:]:
        os.environ["SVO_BENCH_STRONG_CONFIG"] = n
        rec = bench.strong_scaling_record(torch, dist, None, rank, world, local, stream, peak)
        if rank == 0:
            out[n] = rec
            print(n, json.dumps(rec), flush=True)
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "scale_sharded_%d.json" % world), "w") as f:
            json.dump(out, f, indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
