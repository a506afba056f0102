"""Per-rank breakdown of the sharded MD step (configs[4]) under torchrun: host time spent in each call
and device time between events, to find what a real N-rank step adds over one rank's share.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/md_probe.py [steps]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import bench
from uf3_b200 import geometry, synthetic
from uf3_b200.distributed import ShardedEvaluator, atom_range

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
if os.environ.get("UF3B_BLOCKING_SYNC") == "1":
    from uf3_b200 import _native
    _native.check(_native.lib().uf3b_set_blocking_sync(1))
basis, coeff = bench.w_model23()
pos, numbers, cell, pbc = synthetic.bcc_w((25, 40, 50), a=3.206, sigma=0.15, seed=0)
n = len(pos)
images = geometry.image_table(cell, pbc, basis.r_cut)
ev = ShardedEvaluator(basis, coeff, device=local)
x = torch.from_numpy(pos).to(dev)
z = torch.from_numpy(numbers).to(dev)
stream = torch.cuda.current_stream().cuda_stream
out = torch.empty(3 * n + 1, dtype=torch.float64, device=dev)
centres = atom_range(n, rank, world)
eng = ev.engine


def step(t):
    t0 = time.perf_counter()
    eng.build_neighbors_device(x.data_ptr(), z.data_ptr(), n, images, stream, centres=centres)
    t1 = time.perf_counter()
    eng.energy_forces_device(out[3 * n:].data_ptr(), out.data_ptr(), stream)
    t2 = time.perf_counter()
    if world > 1:
        dist.all_reduce(out)
    t3 = time.perf_counter()
    if t is not None:
        t += np.array([t1 - t0, t2 - t1, t3 - t2])


for mode in ("full", "no_allreduce"):
    for _ in range(5):
        step(None)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    host = np.zeros(3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    if mode == "full":
        for _ in range(steps):
            step(host)
    else:
        save, world = world, 1
        for _ in range(steps):
            step(host)
        world = save
    e1.record()
    torch.cuda.synchronize()
    w1 = time.perf_counter()
    res = {"rank": rank, "mode": mode, "gpu_ms_per_step": e0.elapsed_time(e1) / steps, "wall_ms_per_step": (w1 - w0) * 1e3 / steps,
           "host_ms": dict(zip(("build", "energy_forces", "all_reduce"), (host * 1e3 / steps).round(4).tolist()))}
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, res)
        if rank == 0:
            for g in gathered:
                print(json.dumps(g))
    else:
        print(json.dumps(res))
if world > 1:
    # the collective alone
    for _ in range(5):
        dist.all_reduce(out)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        dist.all_reduce(out)
    e1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print(json.dumps({"all_reduce_alone_ms": e0.elapsed_time(e1) / 50, "bytes": out.numel() * 8}))
    dist.destroy_process_group()
