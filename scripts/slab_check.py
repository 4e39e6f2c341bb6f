"""Multi-GPU parity: the row-slab run over NCCL must be bit-identical to the single-GPU run.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 scripts/slab_check.py [width height steps iterations]

Every rank computes the single-GPU reference of the whole grid on its own device, then compares
its slab's rows of every field - and of a dye field cut into the same slabs (2x the grid's resolution
under pipeline 1, 1.5x under pipeline 0) - after each step."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from natrix_b200 import slab_parity  # noqa: E402

width, height, steps, iters = (int(a) for a in (sys.argv[1:5] + ["1024", "2048", "3", "37"][len(sys.argv) - 1:]))
warm = len(sys.argv) > 5 and sys.argv[5] == "warm"        # NATRIX_OPT_WARM_START on both sides
rank, world, local = (int(os.environ.get(k, "0")) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
import json  # noqa: E402

res = slab_parity.check(width, height, steps, iters, local=local, warm=warm, verbose=True)
if rank == 0:
    print("SLAB_CHECK", "PASS" if res["bit_identical"] else "FAIL", f"world={world} grid={width}x{height}", json.dumps(res), flush=True)
# a non-default speed set before the first step (the velocity halo depends on dt * speed), both drivers
res2 = slab_parity.check(width, height, 2, iters, local=local, pipelines=(1,), schedules=(True,), speed=1000.0)
# a fast flow: |v| up to 1.7 and a strong confinement force, so back-traces reach beyond the rows a |v| <= 1 flow
# needs; the library exchanges every allocated velocity halo row, the result must still be bit-identical
res3 = slab_parity.check(width, height, 2, iters, local=local, pipelines=(1,), schedules=(True,), drivers=("native",),
                         v0_scale=1.7, vorticity=7.5)
if rank == 0:
    print("SLAB_CHECK_SPEED", "PASS" if res2["bit_identical"] else "FAIL", json.dumps(res2), flush=True)
    print("SLAB_CHECK_FAST", "PASS" if res3["bit_identical"] else "FAIL", json.dumps(res3), flush=True)
res["bit_identical"] = res["bit_identical"] and res2["bit_identical"] and res3["bit_identical"]
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if res["bit_identical"] else 1)
