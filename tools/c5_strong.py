#!/usr/bin/env python
"""BASELINE config 5 only (25-point constant fp32, 1024^3 x 200, z-slabs over the ranks of this torchrun launch):
prints bench.py's strong_scaling_c5 block.  Measurement tool."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import girih_b200 as G  # noqa: E402

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for copy in (True, False):
    r = bench.strong_scaling_c5(G, dist, world, rank, local, copy)
    if rank == 0:
        print(json.dumps(r), flush=True)
dist.barrier()
dist.destroy_process_group()
