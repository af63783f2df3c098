"""Training step at N ranks, both exchange schedules of FusedTrainStep (coarse all-reduce overlapped with the fine pass's
backward / one all-reduce after both passes), next to the autograd route:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_scale_probe.py

One JSON line per schedule from rank 0 (bench.py's training_probe, max over ranks of CUDA-event times)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench                                # noqa: E402
from anerf_b200 import parallel             # noqa: E402

rank, world, local = parallel.init_distributed()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
for overlap in ("1", "0"):
    os.environ["ANERF_TRAIN_OVERLAP"] = overlap
    t = bench.training_probe(dev, rank, world, steps=10)
    if rank == 0:
        print(json.dumps({"world": world, "overlap_exchange": overlap == "1", "fused_step_ms": t["ms_per_step"],
                          "autograd_route_ms": t["autograd_route"]["ms_per_step"], "allreduce_bytes": t["allreduce_bytes"]}), flush=True)
if world > 1:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
