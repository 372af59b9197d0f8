#!/usr/bin/env python
"""Host-to-device copy bandwidth of every GPU of the box, one at a time and all at once (torchrun, one rank per GPU):
what bounds the end-to-end rate of bench.py at N = 8.  Prints one JSON line on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/copy_fabric.py
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    import rkmh_b200
    cpus = rkmh_b200.bind_host_thread_to_gpu(local) if world > 1 else None
    nbytes = 1 << 30
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    back = torch.empty(nbytes // 8, dtype=torch.uint8).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def h2d(reps=4, with_d2h=False):
        s2 = torch.cuda.Stream()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            dev.copy_(host, non_blocking=True)
            if with_d2h:
                with torch.cuda.stream(s2):
                    back.copy_(dev[: nbytes // 8], non_blocking=True)
        torch.cuda.synchronize()
        return reps * nbytes / (time.perf_counter() - t0) / 1e9

    h2d(1)
    alone = [0.0] * world
    for r in range(world):                      # one GPU at a time
        barrier()
        if r == rank:
            alone[r] = h2d()
        barrier()
    barrier()
    together = h2d()                            # every GPU at once
    barrier()
    together_duplex = h2d(with_d2h=True)
    barrier()
    vals = torch.tensor([alone[rank], together, together_duplex], dtype=torch.float64, device="cuda")
    if world > 1:
        allv = [torch.empty_like(vals) for _ in range(world)]
        dist.all_gather(allv, vals)
    else:
        allv = [vals]
    if rank == 0:
        m = torch.stack(allv).cpu().numpy()
        print(json.dumps({"gpus": world, "h2d_gbs_alone": [round(float(x), 1) for x in m[:, 0]],
                          "h2d_gbs_all_at_once": [round(float(x), 1) for x in m[:, 1]],
                          "h2d_gbs_all_at_once_with_d2h": [round(float(x), 1) for x in m[:, 2]],
                          "sum_all_at_once": round(float(m[:, 1].sum()), 1), "host_cpus": os.cpu_count(),
                          "numa_cpus_rank0": len(cpus) if cpus else None}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
