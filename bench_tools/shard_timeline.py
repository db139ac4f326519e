#!/usr/bin/env python
"""Per-phase timeline of the REAL multi-GPU sharded step (one process per GPU, NVLink transport):

    torchrun --nproc-per-node N bench_tools/shard_timeline.py [--batch 65536] [--steps 30]

Every rank enables the library's step timeline (a CUDA event after each group of launches: the events themselves
keep the launches of a step from overlapping, so the phases add up to a little more than an untimed step) and
rank 0 prints the median microseconds per phase of every rank, plus the untimed step for comparison."""
import argparse, ctypes, json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from arvae_b200 import _lib, synth
from arvae_b200 import distributed as adist

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--steps", type=int, default=30)
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
lib = _lib.load()
c = synth.make_case("c4_mnist_b65536", args.batch)
dims = tuple(c["reg_dims"])
n = c["B"] // world
z = c["z"][rank * n:(rank + 1) * n].to(dev)
lab = c["labels"][rank * n:(rank + 1) * n].to(dev)
comm = adist.ShardComm(n, len(dims))
n_all = [n] * world
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
buf = ctypes.create_string_buffer(4096)
tl, plain = {}, []
for it in range(args.steps + 5):
    for timed in (True, False):
        flush.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if timed:
            lib.arvae_timeline_enable(1)
        e0.record()
        comm.h.step(z, lab, dims, dims, n_all, c["gamma"], c["delta"], True)
        e1.record()
        torch.cuda.synchronize()
        if timed:
            lib.arvae_timeline_report(buf, 4096)
            lib.arvae_timeline_enable(0)
            if it >= 5:
                for item in buf.value.decode().split(";"):
                    if item:
                        k, v = item.rsplit(":", 1)
                        tl.setdefault(k, []).append(float(v) * 1e3)
        elif it >= 5:
            plain.append(e0.elapsed_time(e1) * 1e3)
mine = {k: round(statistics.median(v), 1) for k, v in tl.items()}
mine["untimed_step"] = round(statistics.median(plain), 1)
allr = [None] * world
dist.all_gather_object(allr, mine)
if rank == 0:
    print(json.dumps({"world": world, "B": c["B"], "per_rank_us": allr}))
comm.close()
dist.destroy_process_group()
