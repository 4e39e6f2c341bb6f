"""Strong-scaling point of config 4 (16384^2, N = 100) under torchrun for several slab halo depths:
    torchrun --nproc-per-node N scripts/strong_probe.py 48 104
With `halo` rows of halo a slab runs floor(halo / 8) * 8 sweeps between two pressure exchanges; halo >= N + 4 makes the
whole solve one group: one exchange of the scaled divergence and the mask, no pressure exchange at all (cold start)."""
import argparse
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from natrix_b200 import slabs, workloads as W  # noqa: E402

rank, world, local = (int(os.environ.get(k, "0")) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
args = argparse.Namespace(steps=10)
for halo in sys.argv[1:]:
    os.environ["NATRIX_SLAB_HALO"] = halo
    r = slabs.strong_scaling_point(args, W.cfg4_workload(16384), local, 8, "Mcell-steps/s")
    if rank == 0:
        print(json.dumps({"halo": int(halo), "n_gpus": world, "ms_per_step": r["ms_per_step"], "value": r["value"],
                          "one_gpu_ms": r["one_gpu"]["ms_per_step"], "efficiency": r["efficiency"],
                          "host_enqueue_ms": r["host_enqueue_ms_per_step"]}), flush=True)
dist.barrier()
dist.destroy_process_group()
