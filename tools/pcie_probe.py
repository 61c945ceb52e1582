#!/usr/bin/env python
"""Host <-> device copy bandwidth of every GPU of the box, alone and all at once (torchrun, one rank per GPU): the
host-side ceiling of the end-to-end numbers at N > 1 (every rank moves 72 MB in + 143 MB out per step).

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_probe.py
"""
import json
import os
import time

import torch
import torch.distributed as dist


def bw(fn, nbytes, reps=10):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t) / 1e9


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    n = 256 << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
    s2 = torch.cuda.Stream()
    res = {}

    def both():
        d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)

    for label, fn, nb in (("h2d", lambda: d.copy_(h, non_blocking=True), n), ("d2h", lambda: h.copy_(d, non_blocking=True), n), ("both", both, 2 * n)):
        # alone: one rank at a time
        alone = torch.zeros(world, dtype=torch.float64, device="cuda")
        for r in range(world):
            dist.barrier()
            if r == rank:
                alone[r] = bw(fn, nb)
            dist.barrier()
        dist.all_reduce(alone)
        dist.barrier()
        together = torch.zeros(world, dtype=torch.float64, device="cuda")
        together[rank] = bw(fn, nb)
        dist.all_reduce(together)
        res[label] = {"alone_GBs": [round(float(x), 1) for x in alone], "all_at_once_GBs": [round(float(x), 1) for x in together],
                      "all_at_once_sum_GBs": round(float(together.sum()), 1)}
    if rank == 0:
        print(json.dumps({"world": world, "bytes": n, "pinned": True, **res}), flush=True)
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/pcie_probe_%d.json" % world, "w") as f:
            json.dump({"world": world, **res}, f, indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
